// Binary Topological Interaction loss (and its scalar-label twin TI).
// Replaces BTI_Loss.forward / binary_topological_interaction_module (reference loss/bti_loss.py:76-145;
// loss/ti_loss.py:76-145 is the same with singleton class sets).
//
// The reference builds, per interaction, two fp64 masks with torch.isin and runs two fp64 conv3d with a
// ones(3,3,3) kernel, thresholded at >= 1: that is a binary dilation.  Here every voxel's argmax class c
// becomes the bit 1<<c, the 3^d (or cross) neighbourhood is OR-reduced ONCE into a "classes present"
// word, and all interactions are evaluated from that word with two ANDs each — the result is the same
// boolean map, bit for bit.  The masked cross-entropy (bti_loss.py:141-143) is evaluated per voxel in
// fp64 like the reference (x.double()), summed deterministically (fixed-order two-stage reduction).
//
// Kernels (all HBM-bound, one thread per voxel, coalesced along the fastest spatial axis):
//   bti_argmax_ce   : one read of the logits -> uint8 argmax label + fp32->fp64 CE per voxel
//   bti_critical    : labels -> uint8 critical map (27 byte-loads per voxel, L1/L2 resident)
//   bti_masked_sum  : sum_v ce[v]*crit[v] per batch item (fp64), then mean over the batch
//   bti_ce_bwd      : dlogits = crit * gout/B * (softmax - onehot)
#include "common.cuh"

namespace nextou {

constexpr int BTI_MAX_INTER = 128;   // interactions per LAUNCH (by-value table); longer lists run in chunks that OR into the map
struct BtiTable {
  uint32_t a[BTI_MAX_INTER];
  uint32_t c[BTI_MAX_INTER];
  int n;
};

// target dtype codes
enum { TGT_F32 = 0, TGT_BF16 = 1, TGT_I64 = 2, TGT_U8 = 3, TGT_I32 = 4 };

__device__ __forceinline__ int load_target(const void* t, int code, long long i) {
  switch (code) {
    case TGT_F32: return (int)reinterpret_cast<const float*>(t)[i];
    case TGT_BF16: return (int)__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(t)[i]);
    case TGT_I64: return (int)reinterpret_cast<const long long*>(t)[i];
    case TGT_U8: return (int)reinterpret_cast<const uint8_t*>(t)[i];
    default: return reinterpret_cast<const int*>(t)[i];
  }
}

template <typename T>
__global__ void bti_argmax_ce_kernel(const T* __restrict__ logits, long long sb, long long sc, long long sv, int NC,
                                     long long V, const void* __restrict__ target, int tcode,
                                     uint8_t* __restrict__ labels, double* __restrict__ ce) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (v >= V) return;
  const T* p = logits + (long long)b * sb + v * sv;
  float best = to_f(p[0]);
  for (int c = 1; c < NC; ++c) best = fmaxf(best, to_f(p[(long long)c * sc]));
  // The reference takes argmax(softmax(x)) (bti_loss.py:132-134), not argmax(x): fp32 softmax is not injective — logits
  // closer than ~3e-8 to the maximum get the same exp(x - max) = 1.0f, and torch.argmax then returns the LOWEST such
  // index.  Evaluate the same fp32 expression (exp(x - max) / sum in class order, first maximum wins).
  float e[32];
  float s = 0.f;
  for (int c = 0; c < NC; ++c) {
    e[c] = expf(to_f(p[(long long)c * sc]) - best);
    s += e[c];
  }
  float pbest = -1.f;
  int bi = 0;
  for (int c = 0; c < NC; ++c) {
    const float pc = __fdiv_rn(e[c], s);
    if (pc > pbest) { pbest = pc; bi = c; }
  }
  labels[(long long)b * V + v] = (uint8_t)bi;
  if (ce) {
    const int t = load_target(target, tcode, (long long)b * V + v);
    double s = 0.0;
    for (int c = 0; c < NC; ++c) s += exp((double)to_f(p[(long long)c * sc]) - (double)best);
    // out-of-range targets contribute nothing (CrossEntropyLoss ignore_index semantics)
    ce[(long long)b * V + v] = (t >= 0 && t < NC) ? (log(s) + (double)best) - (double)to_f(p[(long long)t * sc]) : 0.0;
  }
}

__global__ void bti_critical_kernel(const uint8_t* __restrict__ labels, int D, int H, int W, int rd, int r, int cross,
                                    BtiTable tab, int accumulate, uint8_t* __restrict__ crit) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int bz = blockIdx.z;  // b * D + z
  if (x >= W || y >= H) return;
  const int z = bz % D;
  const uint8_t* lab = labels + (long long)(bz - z) * H * W;  // start of this batch item
  const uint32_t me = 1u << lab[((long long)z * H + y) * W + x];
  uint32_t nb = 0;
  for (int dz = -rd; dz <= rd; ++dz) {
    const int zz = z + dz;
    if (zz < 0 || zz >= D) continue;
    for (int dy = -r; dy <= r; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -r; dx <= r; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        if (cross && (abs(dz) + abs(dy) + abs(dx) > 1)) continue;
        nb |= 1u << lab[((long long)zz * H + yy) * W + xx];
      }
    }
  }
  bool c = false;
  for (int t = 0; t < tab.n; ++t) {
    const bool inA = (me & tab.a[t]) != 0, inC = (me & tab.c[t]) != 0;
    const bool nearA = (nb & tab.a[t]) != 0, nearC = (nb & tab.c[t]) != 0;
    c = c || (nearC && inA) || (nearA && inC);
  }
  uint8_t* out = crit + (long long)bz * H * W + (long long)y * W + x;
  if (accumulate) {           // later chunk of a long interaction list: OR into the map of the earlier chunks
    if (c) *out = 1;
  } else {
    *out = c ? 1 : 0;
  }
}

constexpr int SUM_THREADS = 256;
// stage 1: grid (nblk, B): fixed-order per-thread strided sum, then fixed tree in shared memory
__global__ void __launch_bounds__(SUM_THREADS)
    bti_masked_sum_kernel(const double* __restrict__ ce, const uint8_t* __restrict__ crit, long long V,
                          double* __restrict__ partial) {
  __shared__ double sh[SUM_THREADS];
  const int b = blockIdx.y;
  const double* c = ce + (long long)b * V;
  const uint8_t* m = crit + (long long)b * V;
  double s = 0.0;
  for (long long v = (long long)blockIdx.x * SUM_THREADS + threadIdx.x; v < V; v += (long long)gridDim.x * SUM_THREADS)
    if (m[v]) s += c[v];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = SUM_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(long long)b * gridDim.x + blockIdx.x] = sh[0];
}
// stage 2: one thread: per-batch sums in block order, then the batch mean (bti_loss.py:143)
__global__ void bti_finish_kernel(const double* __restrict__ partial, int nblk, int B, double* __restrict__ out) {
  double tot = 0.0;
  for (int b = 0; b < B; ++b) {
    double s = 0.0;
    for (int i = 0; i < nblk; ++i) s += partial[(long long)b * nblk + i];
    tot += s;
  }
  out[0] = tot / (double)B;
}

template <typename T>
__global__ void bti_ce_bwd_kernel(const T* __restrict__ logits, long long sb, long long sc, long long sv, int NC,
                                  long long V, const void* __restrict__ target, int tcode,
                                  const uint8_t* __restrict__ crit, const double* __restrict__ gout, int B,
                                  T* __restrict__ dlogits, long long dsb, long long dsc, long long dsv) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (v >= V) return;
  T* dp = dlogits + (long long)b * dsb + v * dsv;
  const int t = load_target(target, tcode, (long long)b * V + v);
  if (!crit[(long long)b * V + v] || t < 0 || t >= NC) {
    for (int c = 0; c < NC; ++c) dp[(long long)c * dsc] = from_f<T>(0.f);
    return;
  }
  const T* p = logits + (long long)b * sb + v * sv;
  double mx = (double)to_f(p[0]);
  for (int c = 1; c < NC; ++c) mx = fmax(mx, (double)to_f(p[(long long)c * sc]));
  double s = 0.0;
  for (int c = 0; c < NC; ++c) s += exp((double)to_f(p[(long long)c * sc]) - mx);
  const double g = gout[0] / (double)B;
  for (int c = 0; c < NC; ++c) {
    const double sm = exp((double)to_f(p[(long long)c * sc]) - mx) / s;
    dp[(long long)c * dsc] = from_f<T>((float)(g * (sm - (c == t ? 1.0 : 0.0))));
  }
}

}  // namespace nextou

using namespace nextou;

extern "C" int nextou_bti_argmax_ce(const void* logits, int dtype, long long stride_b, long long stride_c,
                                    long long stride_v, int B, int NC, long long V, const void* target, int target_code,
                                    uint8_t* labels, double* ce, void* stream) {
  NEXTOU_REQUIRE(logits && labels, "bti_argmax_ce: null pointer");
  NEXTOU_REQUIRE(B > 0 && B <= 65535 && NC > 0 && NC <= 32 && V > 0, "bti_argmax_ce: bad shape B=%d classes=%d V=%lld (<= 32 classes)", B, NC, V);
  NEXTOU_REQUIRE(ce == nullptr || (target != nullptr && target_code >= 0 && target_code <= 4), "bti_argmax_ce: ce needs a target");
  dim3 grid((unsigned)((V + 255) / 256), B);
  DISPATCH_T(dtype, bti_argmax_ce_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
                        (const T*)logits, stride_b, stride_c, stride_v, NC, V, target, target_code, labels, ce);)
  return check_launch("bti_argmax_ce_kernel");
}

extern "C" int nextou_bti_critical_map(const uint8_t* labels, int B, int D, int H, int W, int dim,
                                       const uint32_t* mask_a_host, const uint32_t* mask_c_host,
                                       const uint8_t* inclusion_host, int n_inter, int connectivity, int min_thick,
                                       uint8_t* crit, void* stream) {
  NEXTOU_REQUIRE(labels && crit, "bti_critical_map: null pointer");
  NEXTOU_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && (long long)B * D <= 65535, "bti_critical_map: bad shape");
  NEXTOU_REQUIRE(n_inter >= 0 && (n_inter == 0 || (mask_a_host && mask_c_host && inclusion_host)), "bti_critical_map: bad interaction table");
  NEXTOU_REQUIRE(dim == 2 || dim == 3, "bti_critical_map: dim=%d", dim);
  NEXTOU_REQUIRE(dim == 3 || D == 1, "bti_critical_map: dim=2 needs D=1");
  const bool box = (dim == 3 && connectivity == 26) || (dim == 2 && connectivity == 8);
  const bool cross = (dim == 3 && connectivity == 6) || (dim == 2 && connectivity == 4);
  NEXTOU_REQUIRE(box || cross, "bti_critical_map: connectivity %d invalid for dim %d (bti_loss.py:57-71)", connectivity, dim);
  NEXTOU_REQUIRE(min_thick >= 1, "bti_critical_map: min_thick=%d", min_thick);
  const int r = box ? min_thick : 1;
  dim3 block(32, 8);
  dim3 grid((W + 31) / 32, (H + 7) / 8, B * D);
  // The reference has no limit on the number of interactions (nnUNetTrainer_NexToU_TI builds all C(n, 2) label pairs: 78 for
  // the 13 Synapse organs).  Duplicate (A, C) pairs cannot change the OR; the rest goes out in chunks of BTI_MAX_INTER.
  BtiTable tab;
  tab.n = 0;
  int launched = 0;
  for (int t = 0; t <= n_inter; ++t) {
    if (t < n_inter) {
      const uint32_t a = mask_a_host[t];
      const uint32_t c = inclusion_host[t] ? ~(mask_c_host[t] | mask_a_host[t]) : mask_c_host[t];  // bti_loss.py:91-95
      bool dup = false;
      for (int j = 0; j < tab.n && !dup; ++j) dup = tab.a[j] == a && tab.c[j] == c;
      if (!dup) { tab.a[tab.n] = a; tab.c[tab.n] = c; ++tab.n; }
    }
    if (tab.n == BTI_MAX_INTER || (t == n_inter && (tab.n > 0 || launched == 0))) {
      bti_critical_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(labels, D, H, W, dim == 3 ? r : 0, r, cross ? 1 : 0,
                                                                    tab, launched > 0 ? 1 : 0, crit);
      int rc = check_launch("bti_critical_kernel");
      if (rc) return rc;
      ++launched;
      tab.n = 0;
    }
  }
  return NEXTOU_OK;
}

extern "C" size_t nextou_bti_masked_sum_workspace_bytes(int B) { return sizeof(double) * (size_t)B * 296; }

extern "C" int nextou_bti_masked_sum(const double* ce, const uint8_t* crit, int B, long long V, double* workspace,
                                     double* out, void* stream) {
  NEXTOU_REQUIRE(ce && crit && workspace && out && B > 0 && B <= 65535 && V > 0, "bti_masked_sum: bad args");
  long long nblk = (V + SUM_THREADS * 8 - 1) / (SUM_THREADS * 8);
  if (nblk > 296) nblk = 296;
  if (nblk < 1) nblk = 1;
  dim3 grid((unsigned)nblk, B);
  bti_masked_sum_kernel<<<grid, SUM_THREADS, 0, (cudaStream_t)stream>>>(ce, crit, V, workspace);
  int rc = check_launch("bti_masked_sum_kernel");
  if (rc) return rc;
  bti_finish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(workspace, (int)nblk, B, out);
  return check_launch("bti_finish_kernel");
}

extern "C" int nextou_bti_ce_bwd(const void* logits, int dtype, long long stride_b, long long stride_c,
                                 long long stride_v, int B, int NC, long long V, const void* target, int target_code,
                                 const uint8_t* crit, const double* grad_out, void* dlogits, long long dstride_b,
                                 long long dstride_c, long long dstride_v, void* stream) {
  NEXTOU_REQUIRE(logits && target && crit && grad_out && dlogits, "bti_ce_bwd: null pointer");
  NEXTOU_REQUIRE(B > 0 && B <= 65535 && NC > 0 && NC <= 32 && V > 0, "bti_ce_bwd: bad shape");
  NEXTOU_REQUIRE(target_code >= 0 && target_code <= 4, "bti_ce_bwd: bad target dtype code %d", target_code);
  dim3 grid((unsigned)((V + 255) / 256), B);
  DISPATCH_T(dtype, bti_ce_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
                        (const T*)logits, stride_b, stride_c, stride_v, NC, V, target, target_code, crit, grad_out, B,
                        (T*)dlogits, dstride_b, dstride_c, dstride_v);)
  return check_launch("bti_ce_bwd_kernel");
}
