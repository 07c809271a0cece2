// Fused segmentation loss for one deep-supervision scale:  w_ce * CE + w_dice * SoftDice + w_ti * (B)TI
// (reference loss/compound_bti_loss.py:33-61; Dice / CE themselves are upstream nnU-Net: MemoryEfficientSoftDiceLoss
// and RobustCrossEntropyLoss with default arguments — SURVEY.md §8f rank 1).  The reference evaluates softmax three
// times per scale (Dice, CE, BTI:132) plus ~20 element-wise passes; here the logits are read twice in total:
//   dsloss_stats : logits + target -> argmax label (u8), per-voxel cross entropy (fp64, for the masked BTI sum),
//                  per-(batch, class) sums  P = sum p_c,  I = sum p_c*y_c,  G = sum y_c,  and sum CE   (fp64 partials)
//   [bti_critical_map + bti_masked_sum from csrc/bti.cu on the labels / CE]
//   dsloss_bwd   : logits + target + critical map + per-class coefficients -> d logits
// The tiny per-class algebra between the two passes (dice value, its derivative coefficients) runs on the host side
// wrapper with torch ops on [B, C]-sized tensors, which also lets it all-reduce the sums for DDP batch-dice.
#include "common.cuh"

namespace nextou {

constexpr int DSL_MAXC = 32;
constexpr int DSL_THREADS = 256;
enum { DT_F32 = 0, DT_BF16 = 1, DT_I64 = 2, DT_U8 = 3, DT_I32 = 4 };

__device__ __forceinline__ int dsl_target(const void* t, int code, long long i) {
  switch (code) {
    case DT_F32: return (int)reinterpret_cast<const float*>(t)[i];
    case DT_BF16: return (int)__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(t)[i]);
    case DT_I64: return (int)reinterpret_cast<const long long*>(t)[i];
    case DT_U8: return (int)reinterpret_cast<const uint8_t*>(t)[i];
    default: return reinterpret_cast<const int*>(t)[i];
  }
}

// grid (nblk, B).  partial[b][blk][3*NC + 1] doubles: P_c, I_c, G_c (c = 0..NC-1), sum CE
template <typename T, int NC>
__global__ void __launch_bounds__(DSL_THREADS)
    dsloss_stats_kernel(const T* __restrict__ logits, long long sb, long long sc, long long sv, long long V,
                        const void* __restrict__ target, int tcode, uint8_t* __restrict__ labels,
                        double* __restrict__ ce_out, double* __restrict__ partial) {
  const int b = blockIdx.y;
  float P[NC], I[NC], G[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) P[c] = I[c] = G[c] = 0.f;
  double ce_sum = 0.0;
  for (long long v = (long long)blockIdx.x * DSL_THREADS + threadIdx.x; v < V; v += (long long)gridDim.x * DSL_THREADS) {
    const T* p = logits + (long long)b * sb + v * sv;
    float x[NC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      x[c] = to_f(p[(long long)c * sc]);
      mx = fmaxf(mx, x[c]);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) { x[c] = expf(x[c] - mx); s += x[c]; }
    // label = argmax(softmax(x)) like the reference (bti_loss.py:132-134): fp32 softmax ties (logits within ~3e-8 of the
    // maximum) resolve to the lowest index — same expression as csrc/bti.cu::bti_argmax_ce_kernel
    float pbest = -1.f;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float pc = __fdiv_rn(x[c], s);
      if (pc > pbest) { pbest = pc; arg = c; }
    }
    const float inv = 1.f / s;
    const int t = dsl_target(target, tcode, (long long)b * V + v);
    float pt = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float pc = x[c] * inv;
      P[c] += pc;
      if (c == t) { I[c] += pc; G[c] += 1.f; pt = pc; }
    }
    const double ce = (t >= 0 && t < NC) ? -(double)__logf(fmaxf(pt, 1e-38f)) : 0.0;
    labels[(long long)b * V + v] = (uint8_t)arg;
    ce_out[(long long)b * V + v] = ce;
    ce_sum += ce;
  }
  // block reduction: warp shuffles then shared memory, fixed order
  __shared__ double sh[DSL_THREADS / 32][3 * NC + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float a = P[c], i2 = I[c], g = G[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      i2 += __shfl_xor_sync(0xffffffffu, i2, o);
      g += __shfl_xor_sync(0xffffffffu, g, o);
    }
    if (lane == 0) { sh[warp][c] = a; sh[warp][NC + c] = i2; sh[warp][2 * NC + c] = g; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ce_sum += __shfl_xor_sync(0xffffffffu, ce_sum, o);
  if (lane == 0) sh[warp][3 * NC] = ce_sum;
  __syncthreads();
  if (threadIdx.x < 3 * NC + 1) {
    double t = 0.0;
    for (int w = 0; w < DSL_THREADS / 32; ++w) t += sh[w][threadIdx.x];
    partial[((long long)b * gridDim.x + blockIdx.x) * (3 * NC + 1) + threadIdx.x] = t;
  }
}

// sums[b][3*NC+1] = fixed-order sum over the CTA partials
__global__ void dsloss_reduce_kernel(const double* __restrict__ partial, int nblk, int width, double* __restrict__ sums) {
  const int b = blockIdx.x, j = threadIdx.x;
  if (j >= width) return;
  double t = 0.0;
  for (int i = 0; i < nblk; ++i) t += partial[((long long)b * nblk + i) * width + j];
  sums[(long long)b * width + j] = t;
}

// d logit_c = dice: p_c (g_c - sum_k p_k g_k), g_c = a[b][c] * y_c + bcoef[b][c];   ce / ti: (s_ce + s_ti * crit) (p_c - y_c)
template <typename T, int NC>
__global__ void __launch_bounds__(DSL_THREADS)
    dsloss_bwd_kernel(const T* __restrict__ logits, long long sb, long long sc, long long sv, long long V,
                      const void* __restrict__ target, int tcode, const uint8_t* __restrict__ crit,
                      const float* __restrict__ coef_a, const float* __restrict__ coef_b, const float* __restrict__ scal,
                      T* __restrict__ dlogits, long long dsb, long long dsc, long long dsv) {
  const int b = blockIdx.y;
  float a[NC], bc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) { a[c] = coef_a[b * NC + c]; bc[c] = coef_b[b * NC + c]; }
  const float s_ce = scal[0], s_ti = scal[1];
  for (long long v = (long long)blockIdx.x * DSL_THREADS + threadIdx.x; v < V; v += (long long)gridDim.x * DSL_THREADS) {
    const T* p = logits + (long long)b * sb + v * sv;
    float x[NC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) { x[c] = to_f(p[(long long)c * sc]); mx = fmaxf(mx, x[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) { x[c] = __expf(x[c] - mx); s += x[c]; }
    const float inv = 1.f / s;
    const int t = dsl_target(target, tcode, (long long)b * V + v);
    const bool tv = t >= 0 && t < NC;
    const float k = tv ? s_ce + (crit != nullptr && crit[(long long)b * V + v] ? s_ti : 0.f) : 0.f;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      x[c] *= inv;
      dot += x[c] * (bc[c] + (c == t ? a[c] : 0.f));
    }
    T* dp = dlogits + (long long)b * dsb + v * dsv;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float g = bc[c] + (c == t ? a[c] : 0.f);
      dp[(long long)c * dsc] = from_f<T>(x[c] * (g - dot) + k * (x[c] - (c == t ? 1.f : 0.f)));
    }
  }
}

// Per-class algebra between the two passes, one tiny launch (it replaces ~25 ATen kernels on [B, NC] doubles per scale):
//   total = w_ce * mean CE  -  w_dice * mean_{b, c >= first} (2 I + smooth) / clip(G + P + smooth, 1e-8)  +  w_ti * ti
//   coef[0][b][c] = d total / d I_bc,  coef[1][b][c] = d total / d P_bc   (zero for c < first; times grad_world under DDP)
// (MemoryEfficientSoftDiceLoss.forward; batch_dice: the sums of the batch items are pooled first.)  sums: [B][3 NC + 1];
// pooled != NULL: [3 NC] already summed over the batch (and all-reduced over the ranks by the caller).
__global__ void dsloss_finish_kernel(const double* __restrict__ sums, const double* __restrict__ pooled, int B, int NC, double V,
                                     double w_ce, double w_dice, double w_ti, const double* __restrict__ ti, int batch_dice,
                                     int first, double smooth, double grad_world, double* __restrict__ total,
                                     double* __restrict__ coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int width = 3 * NC + 1;
  const int rows = batch_dice ? 1 : B;
  const double n_terms = (double)rows * (NC - first);
  double ce = 0.0;
  for (int b = 0; b < B; ++b) ce += sums[(long long)b * width + 3 * NC];
  double dice = 0.0;
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < NC; ++c) {
      double P = 0.0, I = 0.0, G = 0.0;
      if (batch_dice) {
        if (pooled != nullptr) { P = pooled[c]; I = pooled[NC + c]; G = pooled[2 * NC + c]; }
        else for (int b = 0; b < B; ++b) { P += sums[(long long)b * width + c]; I += sums[(long long)b * width + NC + c]; G += sums[(long long)b * width + 2 * NC + c]; }
      } else {
        P = sums[(long long)r * width + c]; I = sums[(long long)r * width + NC + c]; G = sums[(long long)r * width + 2 * NC + c];
      }
      const double raw = G + P + smooth;
      const double den = raw < 1e-8 ? 1e-8 : raw;
      const double num = 2.0 * I + smooth;
      double dI = 0.0, dP = 0.0;
      if (c >= first) {
        dice += num / den;
        dI = -2.0 * w_dice * grad_world / n_terms / den;
        dP = raw >= 1e-8 ? w_dice * grad_world / n_terms * num / (den * den) : 0.0;
      }
      for (int b = (batch_dice ? 0 : r); b < (batch_dice ? B : r + 1); ++b) {
        coef[(long long)b * NC + c] = dI;
        coef[(long long)(B + b) * NC + c] = dP;
      }
    }
  double t = ce * (w_ce / ((double)B * V)) - dice * (w_dice / n_terms);
  if (ti != nullptr) t += w_ti * ti[0];
  total[0] = t;
}

// backward scalars: coef32 = coef * gout, scal = (gout * w_ce / (B V), gout * w_ti / B)
__global__ void dsloss_scale_kernel(const double* __restrict__ coef, const double* __restrict__ gout, int n, double s_ce, double s_ti,
                                    float* __restrict__ coef32, float* __restrict__ scal) {
  const double g = gout[0];
  for (int i = threadIdx.x; i < n; i += blockDim.x) coef32[i] = (float)(coef[i] * g);
  if (threadIdx.x == 0) {
    scal[0] = (float)(g * s_ce);
    scal[1] = (float)(g * s_ti);
  }
}

#define DSL_DISPATCH_NC(NCV, ...)                     \
  switch (NCV) {                                      \
    case 2: { constexpr int NC = 2; __VA_ARGS__ } break;   \
    case 3: { constexpr int NC = 3; __VA_ARGS__ } break;   \
    case 4: { constexpr int NC = 4; __VA_ARGS__ } break;   \
    case 5: { constexpr int NC = 5; __VA_ARGS__ } break;   \
    case 6: { constexpr int NC = 6; __VA_ARGS__ } break;   \
    case 7: { constexpr int NC = 7; __VA_ARGS__ } break;   \
    case 8: { constexpr int NC = 8; __VA_ARGS__ } break;   \
    case 14: { constexpr int NC = 14; __VA_ARGS__ } break; \
    case 16: { constexpr int NC = 16; __VA_ARGS__ } break; \
    case 19: { constexpr int NC = 19; __VA_ARGS__ } break; \
    default:                                          \
      set_error("dsloss: %d classes not instantiated (2-8, 14, 16, 19)", NCV); \
      return NEXTOU_ERR_UNSUPPORTED;                  \
  }

static int dsl_blocks(long long V, int B) {
  long long nblk = (V + DSL_THREADS * 4 - 1) / (DSL_THREADS * 4);
  const long long cap = (4LL * num_sms() + B - 1) / B;
  if (nblk > cap) nblk = cap;
  return (int)(nblk < 1 ? 1 : nblk);
}

}  // namespace nextou

using namespace nextou;

extern "C" int nextou_dsloss_plan(long long V, int B, int* nblk_out) {
  *nblk_out = dsl_blocks(V, B);
  return 0;
}

// labels [B][V] u8, ce [B][V] f64, partial [B][nblk][3*NC+1] f64, sums [B][3*NC+1] f64 (P_c | I_c | G_c | sum CE)
extern "C" int nextou_dsloss_stats(const void* logits, int dtype, long long stride_b, long long stride_c, long long stride_v,
                                   int B, int NC, long long V, const void* target, int target_code, uint8_t* labels,
                                   double* ce, double* partial, double* sums, void* stream) {
  NEXTOU_REQUIRE(logits && target && labels && ce && partial && sums, "dsloss_stats: null pointer");
  NEXTOU_REQUIRE(B > 0 && B <= 65535 && V > 0 && NC >= 2 && NC <= DSL_MAXC, "dsloss_stats: bad shape");
  NEXTOU_REQUIRE(target_code >= 0 && target_code <= 4, "dsloss_stats: bad target dtype code");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = dsl_blocks(V, B);
  dim3 grid(nblk, B);
  DISPATCH_T(dtype, DSL_DISPATCH_NC(NC, dsloss_stats_kernel<T, NC><<<grid, DSL_THREADS, 0, st>>>(
                                            (const T*)logits, stride_b, stride_c, stride_v, V, target, target_code, labels, ce,
                                            partial);))
  int rc = check_launch("dsloss_stats_kernel");
  if (rc) return rc;
  dsloss_reduce_kernel<<<B, 128, 0, st>>>(partial, nblk, 3 * NC + 1, sums);
  return check_launch("dsloss_reduce_kernel");
}

// coef_a / coef_b: [B][NC] fp32 dice coefficients (already scaled by the dice weight and upstream gradient);
// scal[0] = CE scale (w_ce * gout / (B*V)), scal[1] = TI scale (w_ti * gout / B); crit may be NULL.
extern "C" int nextou_dsloss_bwd(const void* logits, int dtype, long long stride_b, long long stride_c, long long stride_v,
                                 int B, int NC, long long V, const void* target, int target_code, const uint8_t* crit,
                                 const float* coef_a, const float* coef_b, const float* scal, void* dlogits,
                                 long long dstride_b, long long dstride_c, long long dstride_v, void* stream) {
  NEXTOU_REQUIRE(logits && target && coef_a && coef_b && scal && dlogits, "dsloss_bwd: null pointer");
  NEXTOU_REQUIRE(B > 0 && B <= 65535 && V > 0 && NC >= 2 && NC <= DSL_MAXC, "dsloss_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(dsl_blocks(V, B), B);
  DISPATCH_T(dtype, DSL_DISPATCH_NC(NC, dsloss_bwd_kernel<T, NC><<<grid, DSL_THREADS, 0, st>>>(
                                            (const T*)logits, stride_b, stride_c, stride_v, V, target, target_code, crit, coef_a,
                                            coef_b, scal, (T*)dlogits, dstride_b, dstride_c, dstride_v);))
  return check_launch("dsloss_bwd_kernel");
}


// total (fp64 scalar) and the Dice derivative coefficients coef[2][B][NC] (fp64) from the sums of nextou_dsloss_stats; `pooled`
// (may be NULL): the [3 NC] batch-pooled P | I | G sums when the caller has already reduced them (DDP batch dice); ti: the
// (B)TI term (device scalar) or NULL.
extern "C" int nextou_dsloss_finish(const double* sums, const double* pooled, int B, int NC, long long V, double w_ce, double w_dice,
                                    double w_ti, const double* ti, int batch_dice, int do_bg, double smooth, double grad_world,
                                    double* total, double* coef, void* stream) {
  NEXTOU_REQUIRE(sums && total && coef && B > 0 && NC >= 2 && NC <= DSL_MAXC && V > 0, "dsloss_finish: bad arguments");
  dsloss_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, pooled, B, NC, (double)V, w_ce, w_dice, w_ti, ti, batch_dice, do_bg ? 0 : 1,
                                                          smooth, grad_world, total, coef);
  return check_launch("dsloss_finish_kernel");
}

// coef32[2*B*NC] = coef * gout[0]; scal[0] = gout * w_ce / (B V), scal[1] = gout * w_ti / B   (inputs of nextou_dsloss_bwd)
extern "C" int nextou_dsloss_scale(const double* coef, const double* gout, int B, int NC, long long V, double w_ce, double w_ti,
                                   float* coef32, float* scal, void* stream) {
  NEXTOU_REQUIRE(coef && gout && coef32 && scal && B > 0 && NC > 0 && V > 0, "dsloss_scale: bad arguments");
  dsloss_scale_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(coef, gout, 2 * B * NC, w_ce / ((double)B * (double)V), w_ti / (double)B, coef32,
                                                          scal);
  return check_launch("dsloss_scale_kernel");
}
