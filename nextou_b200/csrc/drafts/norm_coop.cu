// DRAFT (not built, not run on a GPU yet — see drafts/README.md).
//
// Single-launch train-mode normalisation (+ LeakyReLU, + residual) for L2-RESIDENT activations: the three kernels of
// nextou_norm_stats_tracked + nextou_norm_apply_res (statistics -> finalize -> apply; 4-9 µs each on the 77 small tensors of
// 3d_fullres_nextou, launch / latency bound) become the three phases of ONE cooperative kernel separated by grid barriers.
// Same arithmetic and the same fixed summation orders as csrc/norm.cu (per-thread partials -> shared memory -> per-channel
// sums -> per-CTA partial rows -> fp64 column sums in a fixed order), so results are bit-identical to the 3-kernel path.
//
// Launch with cudaLaunchCooperativeKernel, grid = min(plan.nblk, co-resident CTAs) x 1, block = C*R/NV threads, dynamic
// shared memory 2*C*R floats (as norm_stats_kernel).  instances == 1 only (batch norm; instance norm at batch 1).
#include "../common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace nextou {

constexpr int NVC = 4;

template <typename T> struct VecIOc;
template <> struct VecIOc<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[NVC]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[NVC]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct VecIOc<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[NVC]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[NVC]) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<unsigned*>(&lo);
    u.y = *reinterpret_cast<unsigned*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// total = rows * C must be a multiple of NVC (the host falls back to the 3-kernel path otherwise)
template <typename T>
__global__ void norm_fwd_coop_kernel(const T* __restrict__ x, int C, int R, long long rows, float eps,
                                     float* __restrict__ partial, float* __restrict__ mean, float* __restrict__ invstd,
                                     float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                     long long* __restrict__ num_batches_tracked, int c_valid,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, float slope,
                                     const T* __restrict__ residual, T* __restrict__ y) {
  extern __shared__ float smem[];
  cg::grid_group grid = cg::this_grid();
  const long long total = rows * C;
  const long long S = (long long)C * R;
  const int tid = threadIdx.x;

  // ---- phase 1: per-CTA partial sums (identical to norm_stats_kernel) ----
  float acc[NVC][2] = {};
  for (long long off = (long long)blockIdx.x * S + NVC * tid; off < total; off += (long long)gridDim.x * S) {
    float v[NVC];
    VecIOc<T>::load(x + off, v);
#pragma unroll
    for (int e = 0; e < NVC; ++e) {
      acc[e][0] += v[e];
      acc[e][1] = fmaf(v[e], v[e], acc[e][1]);
    }
  }
#pragma unroll
  for (int e = 0; e < NVC; ++e) {
    smem[NVC * tid + e] = acc[e][0];
    smem[S + NVC * tid + e] = acc[e][1];
  }
  __syncthreads();
  float* prow = partial + (long long)blockIdx.x * 2 * C;
  for (int c = tid; c < 2 * C; c += blockDim.x) {
    const int which = c / C, ch = c % C;
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += smem[which * S + ch + r * C];
    prow[c] = s;
  }
  grid.sync();

  // ---- phase 2: one warp per channel: fixed-order fp64 sums over the CTA rows (lane l adds rows l, l+32, ...; then a
  //      xor butterfly), mean / invstd / running statistics ----
  {
    const int lane = tid & 31;
    const int warps_per_cta = blockDim.x >> 5;                      // full warps only
    const int gw = blockIdx.x * warps_per_cta + (tid >> 5);
    if ((tid >> 5) < warps_per_cta) {
      for (int c = gw; c < C; c += gridDim.x * warps_per_cta) {
        double s1 = 0.0, s2 = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) {
          s1 += (double)partial[(long long)b * 2 * C + c];
          s2 += (double)partial[(long long)b * 2 * C + C + c];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
          const double n = (double)rows;
          const double m = s1 / n;
          double var = s2 / n - m * m;
          if (var < 0.0) var = 0.0;
          mean[c] = (float)m;
          invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
          if (running_mean != nullptr && c < c_valid) {
            const double unbiased = rows > 1 ? var * n / (n - 1.0) : var;
            running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * m);
            running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
          }
        }
      }
    }
    if (num_batches_tracked != nullptr && blockIdx.x == 0 && tid == 0) *num_batches_tracked += 1;
  }
  grid.sync();

  // ---- phase 3: apply (identical to norm_apply_kernel) ----
  float sc[NVC], sh[NVC];
#pragma unroll
  for (int e = 0; e < NVC; ++e) {
    const int ch = (NVC * tid + e) % C;
    const float g = (gamma && ch < c_valid) ? gamma[ch] : 1.f, b = (beta && ch < c_valid) ? beta[ch] : 0.f;
    sc[e] = invstd[ch] * g;
    sh[e] = b - mean[ch] * sc[e];
  }
  for (long long off = (long long)blockIdx.x * S + NVC * tid; off < total; off += (long long)gridDim.x * S) {
    float v[NVC];
    VecIOc<T>::load(x + off, v);
#pragma unroll
    for (int e = 0; e < NVC; ++e) {
      const float t = fmaf(v[e], sc[e], sh[e]);
      v[e] = t > 0.f ? t : t * slope;
    }
    if (residual != nullptr) {
      float r[NVC];
      VecIOc<T>::load(residual + off, r);
#pragma unroll
      for (int e = 0; e < NVC; ++e) v[e] = to_f(from_f<T>(v[e])) + r[e];
    }
    VecIOc<T>::store(y + off, v);
  }
}

}  // namespace nextou

// NOTE for bring-up: the 3-kernel path sums the CTA rows as 32 slices x 8-way unrolled loads (sliced_column_sum); the
// butterfly above is a DIFFERENT fixed order, so mean / invstd may differ from it in the last fp32 bit.  Either port
// sliced_column_sum's order here or accept (and document) the new order for the cooperative path.
using namespace nextou;

extern "C" int nextou_norm_fwd_coop(const void* x, int dtype, int C, int c_valid, long long rows, float eps, float* partial,
                                    float* mean, float* invstd, float* running_mean, float* running_var, float momentum,
                                    long long* num_batches_tracked, const float* gamma, const float* beta, float slope,
                                    const void* residual, void* y, int R, int threads, int nblk, void* stream) {
  NEXTOU_REQUIRE(x && y && partial && mean && invstd, "norm_fwd_coop: null pointer");
  NEXTOU_REQUIRE((rows * C) % NVC == 0 && threads == C * R / NVC && threads <= 1024, "norm_fwd_coop: bad plan");
  const size_t smem = sizeof(float) * 2 * (size_t)C * R;
  void* fn = dtype == NEXTOU_BF16 ? (void*)norm_fwd_coop_kernel<__nv_bfloat16> : (void*)norm_fwd_coop_kernel<float>;
  if (smem > 48 * 1024) NEXTOU_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  NEXTOU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
  NEXTOU_REQUIRE(per_sm >= 1, "norm_fwd_coop: kernel does not fit an SM");
  int grid = per_sm * num_sms();
  if (grid > nblk) grid = nblk;
  void* args[] = {&x, &C, &R, &rows, &eps, &partial, &mean, &invstd, &running_mean, &running_var, &momentum,
                  &num_batches_tracked, &c_valid, &gamma, &beta, &slope, &residual, &y};
  NEXTOU_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, smem, (cudaStream_t)stream));
  return check_launch("norm_fwd_coop_kernel");
}
