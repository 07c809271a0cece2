// Batch / instance normalisation (+ LeakyReLU) on token-major activations, forward and backward.
// Replaces nn.BatchNorm{2,3}d (train-mode batch statistics, TR:54-55; 78 layers in 3d_fullres_nextou),
// nn.InstanceNorm{2,3}d(affine) of the Pool graph convs (ED:22, TN:42-46) and the LeakyReLU(0.01) that follows
// most of them (TR:57, TN:16-17).
//
// All kernels are HBM-bound streaming passes over a dense [rows][C] matrix (C = physical row pitch; zero padded
// channels are simply extra channels).  The matrix is walked as a FLAT array in "sweeps" of R rows, R chosen on the
// host so that C*R is a multiple of VEC and a CTA of C*R/VEC threads covers one sweep with one VEC-wide load per
// thread: every thread then sees the same VEC channels in every sweep (its flat offset advances by a multiple of
// C), keeps their statistics / scale+shift in registers, and all loads are full-width and perfectly coalesced no
// matter how odd C is (33, 66, 132, 324 ...).  Reductions are deterministic: per-thread partials -> shared memory
// -> fixed-order per-channel sums -> per-CTA partial rows in global memory -> fixed-order finalize.
#include "common.cuh"

namespace nextou {

// NV elements per thread per sweep (8-byte bf16 / 16-byte fp32 loads).  Measured on B200 (tools/norm_ab.py): 16-byte bf16
// loads (NV = 8) are no faster in the forward kernels and 40 % slower in the backward ones (108 registers: one CTA per SM).
template <typename T, int NV> struct VecIO;
template <> struct VecIO<float, 4> {
  typedef float4 raw_t;
  static __device__ __forceinline__ raw_t load_raw(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void unpack(const raw_t& t, float (&v)[4]) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
__device__ __forceinline__ void bf16x2_to_f32(unsigned u, float& a, float& b) {
  a = __uint_as_float(u << 16);          // bf16 -> fp32 is a 16-bit shift
  b = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ unsigned f32x2_to_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<unsigned*>(&t);
}
template <> struct VecIO<__nv_bfloat16, 4> {
  typedef uint2 raw_t;
  static __device__ __forceinline__ raw_t load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
  static __device__ __forceinline__ void unpack(const raw_t& u, float (&v)[4]) {
    bf16x2_to_f32(u.x, v[0], v[1]);
    bf16x2_to_f32(u.y, v[2], v[3]);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 u;
    u.x = f32x2_to_bf16x2(v[0], v[1]);
    u.y = f32x2_to_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
  }
};
// guarded load / store for the ragged tail of an instance (total not a multiple of the sweep)
template <typename T, int NV>
__device__ __forceinline__ void load_guard(const T* base, long long off, long long total, float (&v)[NV]) {
  if (off + NV <= total) {
    VecIO<T, NV>::unpack(VecIO<T, NV>::load_raw(base + off), v);
  } else {
#pragma unroll
    for (int e = 0; e < NV; ++e) v[e] = (off + e < total) ? to_f(base[off + e]) : 0.f;
  }
}
template <typename T, int NV>
__device__ __forceinline__ void store_guard(T* base, long long off, long long total, const float (&v)[NV]) {
  if (off + NV <= total) {
    VecIO<T, NV>::store(base + off, v);
  } else {
#pragma unroll
    for (int e = 0; e < NV; ++e)
      if (off + e < total) base[off + e] = from_f<T>(v[e]);
  }
}

// The sweep loop shared by every streaming kernel.  A CTA owns the sweeps s = blockIdx.x, blockIdx.x + gridDim.x, ...; sweeps
// that lie completely inside the instance need no bounds check, so U of them are fetched back to back (U * NIN independent
// vector loads in flight per thread: the HBM pipe needs ~50 bytes in flight per thread at this occupancy) before any
// arithmetic; only the instance's last, ragged sweep takes the guarded path.  f(off, v0, v1) consumes one sweep of this
// thread (v1 is unused when NIN == 1) and may store its result.
template <typename T, int NV, int U, int NIN, typename F>
__device__ __forceinline__ void for_each_sweep(const T* __restrict__ in0, const T* __restrict__ in1, long long total, long long S,
                                               F&& f) {
  typedef typename VecIO<T, NV>::raw_t raw_t;
  const long long nfull = total / S, stride = gridDim.x, toff = (long long)NV * threadIdx.x;
  long long s = blockIdx.x;
  for (; s + (U - 1) * stride < nfull; s += U * stride) {
    raw_t r0[U], r1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      r0[u] = VecIO<T, NV>::load_raw(in0 + (s + u * stride) * S + toff);
      if (NIN > 1) r1[u] = VecIO<T, NV>::load_raw(in1 + (s + u * stride) * S + toff);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float v0[NV], v1[NV];
      VecIO<T, NV>::unpack(r0[u], v0);
      if (NIN > 1) VecIO<T, NV>::unpack(r1[u], v1);
      f((s + u * stride) * S + toff, v0, v1);
    }
  }
  for (; s < nfull; s += stride) {
    float v0[NV], v1[NV];
    VecIO<T, NV>::unpack(VecIO<T, NV>::load_raw(in0 + s * S + toff), v0);
    if (NIN > 1) VecIO<T, NV>::unpack(VecIO<T, NV>::load_raw(in1 + s * S + toff), v1);
    f(s * S + toff, v0, v1);
  }
  if (s == nfull && nfull * S + toff < total) {   // the ragged last sweep (owned by exactly one CTA)
    float v0[NV], v1[NV];
    load_guard<T, NV>(in0, s * S + toff, total, v0);
    if (NIN > 1) load_guard<T, NV>(in1, s * S + toff, total, v1);
    f(s * S + toff, v0, v1);
  }
}

__device__ __forceinline__ float lrelu_grad(float pre, float slope) { return pre > 0.f ? 1.f : slope; }

// ---- two-value per-channel reduction shared by the statistics and the backward-reduce kernels ---------------
// acc[e][0..1] are this thread's partial sums for channel (NV*tid + e) % C.  Writes partial[blk][2][C].
template <int NV>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[NV][2], int C, int R, float* smem,
                                                     float* __restrict__ out) {
  const int tid = threadIdx.x;
  const int S = C * R;  // == NV * blockDim.x
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    smem[NV * tid + e] = acc[e][0];
    smem[S + NV * tid + e] = acc[e][1];
  }
  __syncthreads();
  for (int c = tid; c < 2 * C; c += blockDim.x) {
    const int which = c / C, ch = c % C;
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += smem[which * S + ch + r * C];
    out[c] = s;
  }
}

// one-value variant: acc[e] is this thread's partial for channel (NV*tid + e) % C.  Writes partial[blk][C].
template <int NV>
__device__ __forceinline__ void block_channel_reduce1(float (&acc)[NV], int C, int R, float* smem, float* __restrict__ out) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int e = 0; e < NV; ++e) smem[NV * tid + e] = acc[e];
  __syncthreads();
  for (int ch = tid; ch < C; ch += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += smem[ch + r * C];
    out[ch] = s;
  }
}

// grid (nblk, instances); block C*R/NV threads; dynamic smem 2*C*R floats
template <typename T, int NV>
__global__ void norm_stats_kernel(const T* __restrict__ x, int C, int R, long long rows, float* __restrict__ partial) {
  extern __shared__ float smem[];
  const long long total = rows * C;
  const T* base = x + (long long)blockIdx.y * total;
  float acc[NV][2] = {};
  for_each_sweep<T, NV, 6, 1>(base, base, total, (long long)C * R, [&](long long, const float (&v)[NV], const float (&)[NV]) {
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      acc[e][0] += v[e];
      acc[e][1] = fmaf(v[e], v[e], acc[e][1]);
    }
  });
  block_channel_reduce<NV>(acc, C, R, smem, partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C);
}

// Fixed-order fp64 column sums of the CTA partial rows.  The finalize kernels are pure latency chains (a few hundred rows of a
// few hundred columns from L2), so the rows are spread as widely as possible: a block of 1024 threads owns FIN_COLS = 8 columns
// x 128 row slices (slice s adds rows s, s + 128, ...: at most 5 loads per thread for 592 partial rows, all in flight at once);
// the slices are combined by warp shuffles (lanes of equal column), shared memory across the 32 warps, and shuffles again.
// `cols` = row length of `partial`.  Returns the sum of column `col` in lanes 0..7 of warp 0 (col = blockIdx.x * 8 + lane).
constexpr int FIN_COLS = 8;
constexpr int FIN_THREADS = 1024;
__device__ __forceinline__ double sliced_column_sum(const float* __restrict__ rows_base, int nblk, int cols, int col0, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = col0 + (threadIdx.x & (FIN_COLS - 1));
  const int slice = threadIdx.x / FIN_COLS;                 // 0 .. 127
  double s = 0.0;
  if (c < cols) {
    float v[5];
    int n = 0;
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int b = slice + u * (FIN_THREADS / FIN_COLS);
      v[u] = b < nblk ? rows_base[(long long)b * cols + c] : 0.f;
      n = u;
    }
    (void)n;
#pragma unroll
    for (int u = 0; u < 5; ++u) s += (double)v[u];
    for (int b = slice + 5 * (FIN_THREADS / FIN_COLS); b < nblk; b += FIN_THREADS / FIN_COLS) s += (double)rows_base[(long long)b * cols + c];
  }
  s += __shfl_xor_sync(0xffffffffu, s, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  if (lane < FIN_COLS) sh[warp * FIN_COLS + lane] = s;
  __syncthreads();
  double tot = 0.0;
  if (warp == 0) {
    const int part = lane / FIN_COLS, cc = lane & (FIN_COLS - 1);     // 4 parts x 8 warps each
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += sh[(part * 8 + w) * FIN_COLS + cc];
    tot += __shfl_xor_sync(0xffffffffu, tot, 8);
    tot += __shfl_xor_sync(0xffffffffu, tot, 16);
  }
  __syncthreads();
  return tot;  // meaningful in lanes 0..7 of warp 0
}

// grid (ceil(C/8), instances): mean / invstd per (instance, channel), running stats
__global__ void __launch_bounds__(FIN_THREADS)
    norm_finalize_kernel(const float* __restrict__ partial, int nblk, int C, long long rows, int instances, float eps,
                         float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                         float* __restrict__ running_var, float momentum, long long* __restrict__ num_batches_tracked,
                         int c_valid) {
  __shared__ double sh[32 * FIN_COLS];
  const int inst = blockIdx.y;
  if (num_batches_tracked != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *num_batches_tracked += 1;
  const int col0 = blockIdx.x * FIN_COLS;
  const int c = col0 + (threadIdx.x & (FIN_COLS - 1));
  const float* base = partial + (long long)inst * nblk * 2 * C;
  // columns [0, C) hold sum x, [C, 2C) sum x^2: the second call reads the same channels of the second half
  const double s1 = sliced_column_sum(base, nblk, 2 * C, col0, sh);
  const double s2 = sliced_column_sum(base + C, nblk, 2 * C, col0, sh);
  if (threadIdx.x >= FIN_COLS || c >= C) return;
  const double n = (double)rows;
  const double m = s1 / n;
  double var = s2 / n - m * m;  // biased variance (what normalisation uses)
  if (var < 0.0) var = 0.0;
  mean[inst * C + c] = (float)m;
  invstd[inst * C + c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr && inst == 0 && c < c_valid) {  // batch norm: a single instance spans the whole batch
    const double unbiased = rows > 1 ? var * n / (n - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
  }
}

// y = lrelu((x - mean) * invstd * gamma + beta); slope == 1 -> no activation
template <typename T, int NV>
__global__ void norm_apply_kernel(const T* __restrict__ x, int C, int R, long long rows, const float* __restrict__ mean,
                                  const float* __restrict__ invstd, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float slope, T* __restrict__ y, int c_valid,
                                  const T* __restrict__ residual) {
  const long long total = rows * C;
  const T* base = x + (long long)blockIdx.y * total;
  const T* rbase = residual ? residual + (long long)blockIdx.y * total : nullptr;
  T* obase = y + (long long)blockIdx.y * total;
  float sc[NV], sh[NV];
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    const int ch = (NV * threadIdx.x + e) % C;
    const float g = (gamma && ch < c_valid) ? gamma[ch] : 1.f, b = (beta && ch < c_valid) ? beta[ch] : 0.f;
    sc[e] = invstd[blockIdx.y * C + ch] * g;
    sh[e] = b - mean[blockIdx.y * C + ch] * sc[e];
  }
  auto act = [&](float (&v)[NV]) {
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      const float t = fmaf(v[e], sc[e], sh[e]);
      v[e] = t > 0.f ? t : t * slope;
    }
  };
  if (rbase == nullptr) {
    for_each_sweep<T, NV, 4, 1>(base, base, total, (long long)C * R, [&](long long off, const float (&x0)[NV], const float (&)[NV]) {
      float v[NV];
#pragma unroll
      for (int e = 0; e < NV; ++e) v[e] = x0[e];
      act(v);
      store_guard<T, NV>(obase, off, total, v);
    });
  } else {   // fused residual add: y = T(act(norm(x))) + shortcut, rounded like the two separate ops
    for_each_sweep<T, NV, 2, 2>(base, rbase, total, (long long)C * R, [&](long long off, const float (&x0)[NV], const float (&r)[NV]) {
      float v[NV];
#pragma unroll
      for (int e = 0; e < NV; ++e) v[e] = x0[e];
      act(v);
#pragma unroll
      for (int e = 0; e < NV; ++e) v[e] = to_f(from_f<T>(v[e])) + r[e];
      store_guard<T, NV>(obase, off, total, v);
    });
  }
}

// backward pass 1: per channel  s1 = sum dy',  s2 = sum dy' * xhat,  dy' = dy * lrelu'(pre-activation)
template <typename T, int NV>
__global__ void norm_bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy, int C, int R, long long rows,
                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, float slope,
                                       float* __restrict__ partial, int c_valid) {
  extern __shared__ float smem[];
  const long long total = rows * C;
  const T* xb = x + (long long)blockIdx.y * total;
  const T* db = dy + (long long)blockIdx.y * total;
  float mu[NV], is[NV], g[NV], b[NV];
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    const int ch = (NV * threadIdx.x + e) % C;
    mu[e] = mean[blockIdx.y * C + ch];
    is[e] = invstd[blockIdx.y * C + ch];
    g[e] = (gamma && ch < c_valid) ? gamma[ch] : 1.f;
    b[e] = (beta && ch < c_valid) ? beta[ch] : 0.f;
  }
  float acc[NV][2] = {};
  for_each_sweep<T, NV, 3, 2>(xb, db, total, (long long)C * R, [&](long long, const float (&v)[NV], const float (&d)[NV]) {
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      const float xh = (v[e] - mu[e]) * is[e];
      const float dd = d[e] * lrelu_grad(fmaf(xh, g[e], b[e]), slope);
      acc[e][0] += dd;
      acc[e][1] = fmaf(dd, xh, acc[e][1]);
    }
  });
  block_channel_reduce<NV>(acc, C, R, smem, partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C);
}

// sums[inst][cols] (fp32) from the CTA partial rows [inst][nblk][cols], fixed order, fp64 accumulation;
// grid (ceil(cols/8), instances)
__global__ void __launch_bounds__(FIN_THREADS)
    norm_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int cols, int instances,
                             float* __restrict__ sums) {
  __shared__ double sh[32 * FIN_COLS];
  const int inst = blockIdx.y;
  const int col0 = blockIdx.x * FIN_COLS;
  const int c = col0 + (threadIdx.x & (FIN_COLS - 1));
  const double s = sliced_column_sum(partial + (long long)inst * nblk * cols, nblk, cols, col0, sh);
  if (threadIdx.x < FIN_COLS && c < cols) sums[(long long)inst * cols + c] = (float)s;
}

// backward pass 2: dx = gamma * invstd * (dy' - s1/n - xhat * s2/n)
template <typename T, int NV>
__global__ void norm_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, int C, int R, long long rows,
                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float slope,
                                      const float* __restrict__ sums, T* __restrict__ dx,
                                      float* __restrict__ dx_partial, int c_valid, float inv_n) {
  extern __shared__ float smem[];
  const long long total = rows * C;
  const T* xb = x + (long long)blockIdx.y * total;
  const T* db = dy + (long long)blockIdx.y * total;
  T* ob = dx + (long long)blockIdx.y * total;
  float mu[NV], is[NV], g[NV], b[NV], m1[NV], m2[NV];
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    const int ch = (NV * threadIdx.x + e) % C;
    mu[e] = mean[blockIdx.y * C + ch];
    is[e] = invstd[blockIdx.y * C + ch];
    g[e] = (gamma && ch < c_valid) ? gamma[ch] : 1.f;
    b[e] = (beta && ch < c_valid) ? beta[ch] : 0.f;
    m1[e] = sums[(long long)blockIdx.y * 2 * C + ch] * inv_n;
    m2[e] = sums[(long long)blockIdx.y * 2 * C + C + ch] * inv_n;
  }
  float csum[NV] = {};
  for_each_sweep<T, NV, 3, 2>(xb, db, total, (long long)C * R, [&](long long off, const float (&v)[NV], const float (&d)[NV]) {
    float o[NV];
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      const float xh = (v[e] - mu[e]) * is[e];
      const float dd = d[e] * lrelu_grad(fmaf(xh, g[e], b[e]), slope);
      o[e] = g[e] * is[e] * (dd - m1[e] - xh * m2[e]);
      csum[e] += to_f(from_f<T>(o[e]));   // column sums of dx as stored (= bias gradient of the producing conv / linear)
    }
    store_guard<T, NV>(ob, off, total, o);
  });
  if (dx_partial != nullptr)
    block_channel_reduce1<NV>(csum, C, R, smem, dx_partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * C);
}

// eval-mode affine: y = lrelu(x * scale[c] + shift[c])  (running statistics folded by the caller)
template <typename T, int NV>
__global__ void affine_act_kernel(const T* __restrict__ x, int C, int R, long long rows, const float* __restrict__ scale,
                                  const float* __restrict__ shift, float slope, T* __restrict__ y) {
  const long long total = rows * C;
  float sc[NV], sh[NV];
#pragma unroll
  for (int e = 0; e < NV; ++e) {
    const int ch = (NV * threadIdx.x + e) % C;
    sc[e] = scale[ch];
    sh[e] = shift[ch];
  }
  for_each_sweep<T, NV, 4, 1>(x, x, total, (long long)C * R, [&](long long off, const float (&x0)[NV], const float (&)[NV]) {
    float v[NV];
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      const float t = fmaf(x0[e], sc[e], sh[e]);
      v[e] = t > 0.f ? t : t * slope;
    }
    store_guard<T, NV>(y, off, total, v);
  });
}

struct SweepPlan {
  int R, threads, nblk;
  size_t smem;
};

// R rows per sweep: C*R % NV == 0, C*R/NV <= 1024 threads, aim for ~512 threads
static int plan_sweep(int C, long long rows, int instances, SweepPlan& p) {
  if (C <= 0 || C > 4096 || rows <= 0 || instances <= 0 || instances > 65535) {
    set_error("norm: unsupported shape C=%d rows=%lld instances=%d (C <= 4096)", C, rows, instances);
    return NEXTOU_ERR_INVALID;
  }
  if (instances > 1 && (rows * C) % 4) {
    set_error("norm: rows*C=%lld must be a multiple of 4 when instances > 1 (vector alignment)", rows * C);
    return NEXTOU_ERR_INVALID;
  }
  constexpr int NVp = 4;
  int r0 = 1;
  while ((C * r0) % NVp) ++r0;
  int R = r0;
  while ((long long)C * (R + r0) / NVp <= 512) R += r0;
  if ((long long)C * R / NVp > 1024) {
    set_error("norm: C=%d needs more than 1024 threads per sweep", C);
    return NEXTOU_ERR_INVALID;
  }
  p.R = R;
  p.threads = C * R / NVp;
  const long long sweeps = (rows + R - 1) / R;
  long long nblk = (4LL * num_sms() + instances - 1) / instances;   // one wave at 4 resident CTAs / SM (grid-stride sweeps)
  // small (L2-resident) tensors: fewer, fatter CTAs -> fewer partial rows for the latency-bound finalize (>= 16 sweeps each)
  const long long fat = (sweeps + 15) / 16;
  const long long floor_blk = ((long long)num_sms() + instances - 1) / instances;   // measured: 1 CTA / SM beats 2 for <= 6 MB tensors
  if (nblk > fat) nblk = fat > floor_blk ? fat : floor_blk;
  if (nblk > sweeps) nblk = sweeps;
  if (nblk < 1) nblk = 1;
  p.nblk = (int)nblk;
  p.smem = sizeof(float) * 2 * (size_t)C * R;
  return 0;
}

}  // namespace nextou

using namespace nextou;

#define DISPATCH_TV(dtype, ...)                                       \
  if ((dtype) == NEXTOU_F32) {                                        \
    using T = float; constexpr int NV = 4;                            \
    __VA_ARGS__                                                       \
  } else if ((dtype) == NEXTOU_BF16) {                                \
    using T = __nv_bfloat16; constexpr int NV = 4;                    \
    __VA_ARGS__                                                       \
  } else {                                                            \
    ::nextou::set_error("bad dtype %d", (dtype));                     \
    return NEXTOU_ERR_INVALID;                                        \
  }

extern "C" int nextou_norm_plan(int C, long long rows, int instances, int* nblk_out) {
  SweepPlan p;
  int rc = plan_sweep(C, rows, instances, p);
  if (rc) return rc;
  *nblk_out = p.nblk;
  return 0;
}

extern "C" int nextou_norm_stats(const void* x, int dtype, int C, long long rows, int instances, float eps,
                                 float* partial, float* mean, float* invstd, float* running_mean, float* running_var,
                                 float momentum, void* stream) {
  return nextou_norm_stats_tracked(x, dtype, C, C, rows, instances, eps, partial, mean, invstd, running_mean, running_var,
                                   momentum, nullptr, stream);
}

extern "C" int nextou_norm_stats_tracked(const void* x, int dtype, int C, int c_valid, long long rows, int instances,
                                         float eps, float* partial, float* mean, float* invstd, float* running_mean,
                                         float* running_var, float momentum, long long* num_batches_tracked, void* stream) {
  NEXTOU_REQUIRE(x && partial && mean && invstd, "norm_stats: null pointer");
  SweepPlan p;
  int rc = plan_sweep(C, rows, instances, p);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(p.nblk, instances);
  DISPATCH_TV(dtype, {
    rc = ensure_smem(norm_stats_kernel<T, NV>, p.smem);
    if (rc) return rc;
    norm_stats_kernel<T, NV><<<grid, p.threads, p.smem, st>>>((const T*)x, C, p.R, rows, partial);
  })
  rc = check_launch("norm_stats_kernel");
  if (rc) return rc;
  norm_finalize_kernel<<<dim3((C + FIN_COLS - 1) / FIN_COLS, instances), FIN_THREADS, 0, st>>>(
      partial, p.nblk, C, rows, instances, eps, mean, invstd, running_mean, running_var, momentum, num_batches_tracked,
      c_valid);
  return check_launch("norm_finalize_kernel");
}

extern "C" int nextou_norm_apply(const void* x, int dtype, int C, long long rows, int instances, const float* mean,
                                 const float* invstd, const float* gamma, const float* beta, float slope, void* y,
                                 void* stream) {
  return nextou_norm_apply_cv(x, dtype, C, C, rows, instances, mean, invstd, gamma, beta, slope, y, stream);
}

extern "C" int nextou_norm_apply_cv(const void* x, int dtype, int C, int c_valid, long long rows, int instances,
                                    const float* mean, const float* invstd, const float* gamma, const float* beta,
                                    float slope, void* y, void* stream) {
  return nextou_norm_apply_res(x, dtype, C, c_valid, rows, instances, mean, invstd, gamma, beta, slope, nullptr, y, stream);
}

extern "C" int nextou_norm_apply_res(const void* x, int dtype, int C, int c_valid, long long rows, int instances,
                                     const float* mean, const float* invstd, const float* gamma, const float* beta,
                                     float slope, const void* residual, void* y, void* stream) {
  NEXTOU_REQUIRE(x && y && mean && invstd, "norm_apply: null pointer");
  SweepPlan p;
  int rc = plan_sweep(C, rows, instances, p);
  if (rc) return rc;
  dim3 grid(p.nblk, instances);
  DISPATCH_TV(dtype, norm_apply_kernel<T, NV><<<grid, p.threads, 0, (cudaStream_t)stream>>>(
                               (const T*)x, C, p.R, rows, mean, invstd, gamma, beta, slope, (T*)y, c_valid, (const T*)residual);)
  return check_launch("norm_apply_kernel");
}

extern "C" int nextou_norm_bwd(const void* x, const void* dy, int dtype, int C, long long rows, int instances,
                               const float* mean, const float* invstd, const float* gamma, const float* beta,
                               float slope, float* partial, float* sums, void* dx, void* stream) {
  return nextou_norm_bwd_colsum(x, dy, dtype, C, C, rows, instances, mean, invstd, gamma, beta, slope, partial, sums, dx,
                                nullptr, stream);
}

extern "C" int nextou_norm_bwd_colsum(const void* x, const void* dy, int dtype, int C, int c_valid, long long rows,
                                      int instances, const float* mean, const float* invstd, const float* gamma,
                                      const float* beta, float slope, float* partial, float* sums, void* dx,
                                      float* dx_colsum, void* stream) {
  int rc = nextou_norm_bwd_reduce(x, dy, dtype, C, c_valid, rows, instances, mean, invstd, gamma, beta, slope, partial, sums, stream);
  if (rc) return rc;
  return nextou_norm_bwd_apply(x, dy, dtype, C, c_valid, rows, instances, rows, mean, invstd, gamma, beta, slope, sums, partial,
                               dx, dx_colsum, stream);
}

// backward step 1: sums[inst][0][C] = sum dy', sums[inst][1][C] = sum dy' * xhat over THIS rank's rows
extern "C" int nextou_norm_bwd_reduce(const void* x, const void* dy, int dtype, int C, int c_valid, long long rows,
                                      int instances, const float* mean, const float* invstd, const float* gamma,
                                      const float* beta, float slope, float* partial, float* sums, void* stream) {
  NEXTOU_REQUIRE(x && dy && mean && invstd && partial && sums, "norm_bwd_reduce: null pointer");
  SweepPlan p;
  int rc = plan_sweep(C, rows, instances, p);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(p.nblk, instances);
  DISPATCH_TV(dtype, {
    rc = ensure_smem(norm_bwd_reduce_kernel<T, NV>, p.smem);
    if (rc) return rc;
    norm_bwd_reduce_kernel<T, NV><<<grid, p.threads, p.smem, st>>>((const T*)x, (const T*)dy, C, p.R, rows, mean, invstd,
                                                                  gamma, beta, slope, partial, c_valid);
  })
  rc = check_launch("norm_bwd_reduce_kernel");
  if (rc) return rc;
  norm_bwd_finalize_kernel<<<dim3((2 * C + FIN_COLS - 1) / FIN_COLS, instances), FIN_THREADS, 0, st>>>(partial, p.nblk, 2 * C,
                                                                                              instances, sums);
  return check_launch("norm_bwd_finalize_kernel");
}

// backward step 2: dx = gamma*invstd*(dy' - sums0/n_total - xhat*sums1/n_total); n_total = rows normalised together
// (== rows locally; the global row count when `sums` were all-reduced over the ranks of a SyncBatchNorm)
extern "C" int nextou_norm_bwd_apply(const void* x, const void* dy, int dtype, int C, int c_valid, long long rows,
                                     int instances, long long n_total, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, float slope, const float* sums, float* partial,
                                     void* dx, float* dx_colsum, void* stream) {
  NEXTOU_REQUIRE(x && dy && dx && mean && invstd && sums && n_total > 0, "norm_bwd_apply: null pointer");
  NEXTOU_REQUIRE(dx_colsum == nullptr || partial != nullptr, "norm_bwd_apply: dx_colsum needs the partial workspace");
  SweepPlan p;
  int rc = plan_sweep(C, rows, instances, p);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(p.nblk, instances);
  float* dx_partial = dx_colsum ? partial : nullptr;
  const size_t smem2 = dx_colsum ? p.smem / 2 : 0;
  DISPATCH_TV(dtype, {
    rc = ensure_smem(norm_bwd_apply_kernel<T, NV>, smem2);
    if (rc) return rc;
    norm_bwd_apply_kernel<T, NV><<<grid, p.threads, smem2, st>>>((const T*)x, (const T*)dy, C, p.R, rows, mean, invstd, gamma,
                                                                beta, slope, sums, (T*)dx, dx_partial, c_valid,
                                                                1.f / (float)n_total);
  })
  rc = check_launch("norm_bwd_apply_kernel");
  if (rc || !dx_colsum) return rc;
  norm_bwd_finalize_kernel<<<dim3((C + FIN_COLS - 1) / FIN_COLS, instances), FIN_THREADS, 0, st>>>(partial, p.nblk, C, instances,
                                                                                          dx_colsum);
  return check_launch("norm_bwd_finalize_kernel");
}

extern "C" int nextou_affine_act(const void* x, int dtype, int C, long long rows, const float* scale,
                                 const float* shift, float slope, void* y, void* stream) {
  NEXTOU_REQUIRE(x && y && scale && shift, "affine_act: null pointer");
  SweepPlan p;
  int rc = plan_sweep(C, rows, 1, p);
  if (rc) return rc;
  DISPATCH_TV(dtype, affine_act_kernel<T, NV><<<p.nblk, p.threads, 0, (cudaStream_t)stream>>>((const T*)x, C, p.R, rows,
                                                                                               scale, shift, slope, (T*)y);)
  return check_launch("affine_act_kernel");
}

// Per-channel column sums of a [rows][C] matrix: sums[0][C] = sum_r x, sums[1][C] = sum_r x^2 (fp32).  Used for the
// bias gradients of the GEMM / convolution layers (d bias = column sum of dY).
extern "C" int nextou_colsum(const void* x, int dtype, int C, long long rows, float* partial, float* sums,
                             void* stream) {
  NEXTOU_REQUIRE(x && partial && sums, "colsum: null pointer");
  SweepPlan p;
  int rc = plan_sweep(C, rows, 1, p);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_TV(dtype, {
    rc = ensure_smem(norm_stats_kernel<T, NV>, p.smem);
    if (rc) return rc;
    norm_stats_kernel<T, NV><<<dim3(p.nblk, 1), p.threads, p.smem, st>>>((const T*)x, C, p.R, rows, partial);
  })
  rc = check_launch("norm_stats_kernel");
  if (rc) return rc;
  norm_bwd_finalize_kernel<<<dim3((2 * C + FIN_COLS - 1) / FIN_COLS, 1), FIN_THREADS, 0, st>>>(partial, p.nblk, 2 * C, 1, sums);
  return check_launch("norm_bwd_finalize_kernel");
}


// The first half of nextou_norm_stats / nextou_norm_bwd_reduce alone: this rank's per-CTA partial rows [nblk][2C]
// (nblk = nextou_norm_plan), for the cross-rank finalize of SyncBatchNorm (csrc/syncnorm.cu).
extern "C" int nextou_norm_partial_stats(const void* x, int dtype, int C, long long rows, float* partial, int* nblk_out,
                                         void* stream) {
  NEXTOU_REQUIRE(x && partial && nblk_out, "norm_partial_stats: null pointer");
  SweepPlan p;
  int rc = plan_sweep(C, rows, 1, p);
  if (rc) return rc;
  DISPATCH_TV(dtype, {
    rc = ensure_smem(norm_stats_kernel<T, NV>, p.smem);
    if (rc) return rc;
    norm_stats_kernel<T, NV><<<dim3(p.nblk, 1), p.threads, p.smem, (cudaStream_t)stream>>>((const T*)x, C, p.R, rows, partial);
  })
  *nblk_out = p.nblk;
  return check_launch("norm_stats_kernel");
}

extern "C" int nextou_norm_bwd_partial(const void* x, const void* dy, int dtype, int C, int c_valid, long long rows,
                                       const float* mean, const float* invstd, const float* gamma, const float* beta, float slope,
                                       float* partial, int* nblk_out, void* stream) {
  NEXTOU_REQUIRE(x && dy && mean && invstd && partial && nblk_out, "norm_bwd_partial: null pointer");
  SweepPlan p;
  int rc = plan_sweep(C, rows, 1, p);
  if (rc) return rc;
  DISPATCH_TV(dtype, {
    rc = ensure_smem(norm_bwd_reduce_kernel<T, NV>, p.smem);
    if (rc) return rc;
    norm_bwd_reduce_kernel<T, NV><<<dim3(p.nblk, 1), p.threads, p.smem, (cudaStream_t)stream>>>(
        (const T*)x, (const T*)dy, C, p.R, rows, mean, invstd, gamma, beta, slope, partial, c_valid);
  })
  *nblk_out = p.nblk;
  return check_launch("norm_bwd_reduce_kernel");
}
