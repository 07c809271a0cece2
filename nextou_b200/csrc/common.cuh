// Shared helpers for libnextou_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/nextou_b200.h"

namespace nextou {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return NEXTOU_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return NEXTOU_OK;
}

#define NEXTOU_REQUIRE(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      ::nextou::set_error(__VA_ARGS__);      \
      return NEXTOU_ERR_INVALID;             \
    }                                        \
  } while (0)

#define NEXTOU_CUDA(call)                                                        \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      ::nextou::set_error("%s failed: %s", #call, cudaGetErrorString(e__));      \
      return NEXTOU_ERR_CUDA;                                                    \
    }                                                                            \
  } while (0)

// opt in to > 48 KB dynamic shared memory once per kernel
template <typename K>
inline int ensure_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e));
      return NEXTOU_ERR_CUDA;
    }
  }
  return NEXTOU_OK;
}

__device__ __forceinline__ float ld_as_f32(const void* p, int dtype, long long i) {
  if (dtype == NEXTOU_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return reinterpret_cast<const float*>(p)[i];
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }


// Weight packing, one output element: master weight w[R][Cc/groups][taps] (nn.Conv / nn.ConvTranspose layout) ->
//   A[r][t][c]   (row pitch taps*lda_c) = w[r][c][t]     forward operand        ([Cout][taps][Cin pad])
//   Bt[c][t'][r] (row pitch taps*ldb_c) = w[r][c][t]     data-gradient operand  ([Cin][taps][Cout pad]), t' = flipped t
// i indexes the A elements first, then the Bt elements.  Zero where c / r are channel padding, outside the diagonal blocks of
// a grouped layer, and inside the channel gap [gap_lo, gap_hi) that the decoder's concatenation layout [up | gap | skip]
// inserts into the INPUT channel axis (physical channel c -> logical channel c - (gap_hi - gap_lo) behind the gap).
template <typename T>
__device__ __forceinline__ void pack_element(const NextouPackJob& j, const T* __restrict__ w, long long i) {
  const int gapw = j.gap_hi - j.gap_lo;
  const int Cp = j.Cc + gapw;
  const long long na = (long long)j.R * j.taps * j.lda_c;
  const int cpg = j.Cc / j.groups, rpg = j.R / j.groups;
  int r, c, t;
  if (i < na) {
    c = (int)(i % j.lda_c);
    t = (int)((i / j.lda_c) % j.taps);
    r = (int)(i / ((long long)j.lda_c * j.taps));
  } else {
    const long long q = i - na;
    r = (int)(q % j.ldb_c);
    const int tf = (int)((q / j.ldb_c) % j.taps);
    t = j.flip_b ? j.taps - 1 - tf : tf;
    c = (int)(q / ((long long)j.ldb_c * j.taps));
  }
  float v = 0.f;
  if (r < j.R && c < Cp && !(c >= j.gap_lo && c < j.gap_hi)) {
    const int cl = c < j.gap_lo ? c : c - gapw;
    if (j.groups == 1 || cl / cpg == r / rpg) v = to_f(w[((long long)r * cpg + (j.groups == 1 ? cl : cl % cpg)) * j.taps + t]);
  }
  if (i < na) reinterpret_cast<__nv_bfloat16*>(j.A)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(j.Bt)[i - na] = __float2bfloat16_rn(v);
}

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace nextou

#define DISPATCH_T(dtype, ...)                                   \
  if ((dtype) == NEXTOU_F32) {                                   \
    using T = float;                                             \
    __VA_ARGS__                                                  \
  } else if ((dtype) == NEXTOU_BF16) {                           \
    using T = __nv_bfloat16;                                     \
    __VA_ARGS__                                                  \
  } else {                                                       \
    ::nextou::set_error("bad dtype %d", (dtype));                          \
    return NEXTOU_ERR_INVALID;                                   \
  }


