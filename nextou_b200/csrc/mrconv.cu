// Max-relative graph convolution, message-passing half.
// Replaces MRConv.forward lines 401-409 of the reference (network_architecture/NexToU_Encoder_Decoder.py):
//   x_i = batched_index_select(x, edge_index[1]); x_j = batched_index_select(y|x, edge_index[0])
//   m = max_j (x_j - x_i);  out = interleave(x, m) on channels        (torch_nn.py:94-115 is the gather)
// The reference materialises two (B', C, N, k) tensors (318 MB each at Swin s2); here one warp owns a
// query token, keeps its k neighbour row ids in registers, streams the neighbour rows (L2-resident:
// the candidate set of a graph is <= 1344 rows) and writes the interleaved [x0, m0, x1, m1, ...] row
// once, plus the arg-max neighbour slot (uint8) for the backward scatter.
// Swin windows are addressed through row maps (torch.roll + window_partition folded into the index,
// ED:634-660, 784), so no partition / reverse / roll copy exists.
#include "common.cuh"

namespace nextou {

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<__nv_bfloat16> { using type = __nv_bfloat162; };

__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld2(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<unsigned*>(&lo);
  u.y = *reinterpret_cast<unsigned*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}

constexpr int MR_WARPS = 8;

template <typename T>
__global__ void __launch_bounds__(MR_WARPS * 32)
    mrconv_gather_fwd_kernel(const T* __restrict__ x, long long ldx, const T* __restrict__ y, long long ldy, int C,
                             const int32_t* __restrict__ idx, int k, const int32_t* __restrict__ qmap,
                             const int32_t* __restrict__ ymap, long long R, int N, int M, T* __restrict__ out,
                             long long ldo, uint8_t* __restrict__ arg) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * MR_WARPS + (threadIdx.x >> 5);
  if (r >= R) return;
  const long long g = r / N;
  const long long qrow = qmap ? qmap[r] : r;
  long long nbrow = 0;
  if (lane < k) {
    const long long j = g * M + idx[r * k + lane];
    nbrow = ymap ? ymap[j] : j;
  }
  const T* xr = x + qrow * ldx;
  T* orow = out + qrow * ldo;
  uint8_t* arow = arg + qrow * C;
  // every lane takes part in the shuffles: iterate to a warp-uniform bound, clamp the channel of idle lanes
  const int iters = (C + 63) / 64;
  for (int it = 0; it < iters; ++it) {
    const int c = it * 64 + lane * 2;
    const bool act = c < C;
    const int cc = act ? c : 0;
    const float2 xv = ld2(xr + cc);
    float b0 = -INFINITY, b1 = -INFINITY;
    int a0 = 0, a1 = 0;
    for (int j = 0; j < k; ++j) {
      const long long row = __shfl_sync(0xffffffffu, nbrow, j);
      const float2 yv = ld2(y + row * ldy + cc);
      const float d0 = yv.x - xv.x, d1 = yv.y - xv.y;
      if (d0 > b0) { b0 = d0; a0 = j; }
      if (d1 > b1) { b1 = d1; a1 = j; }
    }
    if (act) {
      st4(orow + 2 * c, xv.x, b0, xv.y, b1);
      *reinterpret_cast<uchar2*>(arow + c) = make_uchar2((unsigned char)a0, (unsigned char)a1);
    }
  }
}

// d(out) -> d(x) (query rows) and d(y) (candidate rows), fp32 accumulation buffers (pre-zeroed by the caller;
// dy may alias dx for the self graph).  Backward of gather = scatter-add to the arg-max neighbour
// (ATen does this with index_put_ atomics as well).
template <typename T>
__global__ void __launch_bounds__(MR_WARPS * 32)
    mrconv_gather_bwd_kernel(const T* __restrict__ dout, long long ldo, int C, const int32_t* __restrict__ idx, int k,
                             const uint8_t* __restrict__ arg, const int32_t* __restrict__ qmap,
                             const int32_t* __restrict__ ymap, long long R, int N, int M, float* __restrict__ dx,
                             long long lddx, float* __restrict__ dy, long long lddy) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * MR_WARPS + (threadIdx.x >> 5);
  if (r >= R) return;
  const long long g = r / N;
  const long long qrow = qmap ? qmap[r] : r;
  long long nbrow = 0;
  if (lane < k) {
    const long long j = g * M + idx[r * k + lane];
    nbrow = ymap ? ymap[j] : j;
  }
  const T* drow = dout + qrow * ldo;
  const uint8_t* arow = arg + qrow * C;
  // every lane must take part in the shuffles: iterate to a warp-uniform bound
  const int iters = (C + 63) / 64;
  for (int it = 0; it < iters; ++it) {
    const int c = it * 64 + lane * 2;
    const bool act = c < C;
    float2 dxm0 = make_float2(0.f, 0.f), dxm1 = make_float2(0.f, 0.f);
    int a0 = 0, a1 = 0;
    if (act) {
      dxm0 = ld2(drow + 2 * c);      // (d x_c, d m_c)
      dxm1 = ld2(drow + 2 * c + 2);  // (d x_{c+1}, d m_{c+1})
      const uchar2 a = *reinterpret_cast<const uchar2*>(arow + c);
      a0 = a.x;
      a1 = a.y;
    }
    const long long r0 = __shfl_sync(0xffffffffu, nbrow, a0);
    const long long r1 = __shfl_sync(0xffffffffu, nbrow, a1);
    if (act) {
      atomicAdd(dx + qrow * lddx + c, dxm0.x - dxm0.y);
      atomicAdd(dx + qrow * lddx + c + 1, dxm1.x - dxm1.y);
      atomicAdd(dy + r0 * lddy + c, dxm0.y);
      atomicAdd(dy + r1 * lddy + c + 1, dxm1.y);
    }
  }
}

}  // namespace nextou

using namespace nextou;

extern "C" int nextou_mrconv_gather_fwd(const void* x, long long ldx, const void* y, long long ldy, int dtype, int C,
                                        const int32_t* idx, int k, const int32_t* q_row_map,
                                        const int32_t* y_row_map, long long R, int N, int M, void* out,
                                        long long ldo, uint8_t* arg, void* stream) {
  NEXTOU_REQUIRE(x && y && idx && out && arg, "mrconv_gather_fwd: null pointer");
  NEXTOU_REQUIRE(C > 0 && C % 2 == 0 && ldx % 2 == 0 && ldy % 2 == 0 && ldo % 4 == 0,
                 "mrconv_gather_fwd: C=%d, ldx, ldy must be even and ldo %% 4 == 0", C);
  NEXTOU_REQUIRE(k >= 1 && k <= 32, "mrconv_gather_fwd: k=%d outside [1,32]", k);
  NEXTOU_REQUIRE(R > 0 && N > 0 && M > 0 && R % N == 0, "mrconv_gather_fwd: bad R=%lld N=%d M=%d", R, N, M);
  const unsigned blocks = (unsigned)((R + MR_WARPS - 1) / MR_WARPS);
  DISPATCH_T(dtype, mrconv_gather_fwd_kernel<T><<<blocks, MR_WARPS * 32, 0, (cudaStream_t)stream>>>(
                        (const T*)x, ldx, (const T*)y, ldy, C, idx, k, q_row_map, y_row_map, R, N, M, (T*)out, ldo,
                        arg);)
  return check_launch("mrconv_gather_fwd_kernel");
}

extern "C" int nextou_mrconv_gather_bwd(const void* dout, long long ldo, int dtype, int C, const int32_t* idx, int k,
                                        const uint8_t* arg, const int32_t* q_row_map, const int32_t* y_row_map,
                                        long long R, int N, int M, float* dx, long long lddx, float* dy,
                                        long long lddy, void* stream) {
  NEXTOU_REQUIRE(dout && idx && arg && dx && dy, "mrconv_gather_bwd: null pointer");
  NEXTOU_REQUIRE(C > 0 && C % 2 == 0 && ldo % 2 == 0, "mrconv_gather_bwd: C=%d and ldo must be even", C);
  NEXTOU_REQUIRE(k >= 1 && k <= 32, "mrconv_gather_bwd: k=%d outside [1,32]", k);
  NEXTOU_REQUIRE(R > 0 && N > 0 && M > 0 && R % N == 0, "mrconv_gather_bwd: bad R=%lld N=%d M=%d", R, N, M);
  const unsigned blocks = (unsigned)((R + MR_WARPS - 1) / MR_WARPS);
  DISPATCH_T(dtype, mrconv_gather_bwd_kernel<T><<<blocks, MR_WARPS * 32, 0, (cudaStream_t)stream>>>(
                        (const T*)dout, ldo, C, idx, k, arg, q_row_map, y_row_map, R, N, M, dx, lddx, dy, lddy);)
  return check_launch("mrconv_gather_bwd_kernel");
}
