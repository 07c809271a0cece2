// Raw PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarrier, TMA, tensor memory, UMMA descriptors.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace nextou {

// ------------------------------------------------------------------------------------------------------
// raw PTX wrappers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug must trap (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  const long long t_start = clock64();
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if ((spin & 1023u) == 1023u && clock64() - t_start > 4000000000LL) __trap();  // ~2 s: never hang the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// one lane of a fully converged warp (warp-uniform control flow around it keeps descriptors in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100 version field = 1):
// rows are 128 B (64 bf16), 8-row groups are 1024 B apart (SBO); LBO is unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;            // leading byte offset (16 B units) — ignored
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D = fp32, A = B = bf16, both K-major, dense
__device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}


// ---- epilogue store: 16 consecutive columns of one output row ---------------------------------------
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <typename OutT>
__device__ __forceinline__ void store_chunk16(OutT* row_ptr, int col0, const float (&v)[16], long long ldc) {
  // col0 is a multiple of 16 and the row pitch a multiple of 8 elements.  Full chunks whose address is 32-byte aligned
  // go out as 256-bit stores (one full sector per lane instead of two half-sector stores); else 16-byte halves.
  OutT* dst = row_ptr + col0;
  if (col0 + 16 <= ldc && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
    if constexpr (sizeof(OutT) == 2) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      st_global_256(dst, w);
    } else {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = __float_as_uint(v[j]);
      st_global_256(dst, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = __float_as_uint(v[8 + j]);
      st_global_256(dst + 8, w);
    }
    return;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = col0 + h * 8;
    if (c + 8 <= ldc) {
      if constexpr (sizeof(OutT) == 2) {
        uint4 u;
        u.x = pack_bf16x2(v[h * 8 + 0], v[h * 8 + 1]);
        u.y = pack_bf16x2(v[h * 8 + 2], v[h * 8 + 3]);
        u.z = pack_bf16x2(v[h * 8 + 4], v[h * 8 + 5]);
        u.w = pack_bf16x2(v[h * 8 + 6], v[h * 8 + 7]);
        *reinterpret_cast<uint4*>(row_ptr + c) = u;
      } else {
        *reinterpret_cast<float4*>(row_ptr + c) = make_float4(v[h * 8 + 0], v[h * 8 + 1], v[h * 8 + 2], v[h * 8 + 3]);
        *reinterpret_cast<float4*>(row_ptr + c + 4) = make_float4(v[h * 8 + 4], v[h * 8 + 5], v[h * 8 + 6], v[h * 8 + 7]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c + j < ldc) row_ptr[c + j] = from_f<OutT>(v[h * 8 + j]);
    }
  }
}


// ---- streamlined bf16 epilogue ---------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st_global_128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// The CTA's bias slice in shared memory: sbias[i] = bias[n0 + i] for n0 + i < N, else 0 (so padded columns come out as
// exact zeros: their accumulators are zero because the weight rows beyond N are zero-filled by TMA).
// sbias is float[512]: [0, 256) the additive term, [256, 512) the per-column scale (1 unless an inference-time norm is folded in)
__device__ __forceinline__ void stage_bias(float* sbias, const float* bias, int n0, int N, int block_n,
                                           const float* scale = nullptr) {
  for (int i = threadIdx.x; i < block_n; i += blockDim.x) {
    sbias[i] = (bias != nullptr && n0 + i < N) ? bias[n0 + i] : 0.f;
    sbias[256 + i] = (scale != nullptr && n0 + i < N) ? scale[n0 + i] : 1.f;
  }
}

// bf16 epilogue of ONE accumulator row (this thread's TMEM lane): columns [0, block_n) of the CTA's N tile are read
// 32 at a time, biased, rounded and stored with 128/256-bit stores.  All lanes must call it (the TMEM loads are
// warp-collective); dst == nullptr skips the stores (row outside the volume).  ncols = number of columns that exist in
// the output row from n0 on (min(block_n, ldc - n0), a multiple of 8); row32 = every row start is 32-byte aligned.
// out = lrelu(acc * scale[col] + shift[col], slope): scale = 1, slope = 1 is the plain "+ bias" epilogue of training; an
// eval-mode BatchNorm (+ LeakyReLU) behind the layer is folded in through scale / shift / slope (inference path)
__device__ __forceinline__ float affine_act(uint32_t acc, float scale, float shift, float slope) {
  const float t = fmaf(__uint_as_float(acc), scale, shift);
  return t > 0.f ? t : t * slope;
}
template <int CHUNK>
__device__ __forceinline__ void epilogue_store_chunk(const uint32_t* raw, const float* sb, __nv_bfloat16* dst, int c, int ncols,
                                                     bool row32, float slope, bool plain = false) {
  uint32_t w[CHUNK / 2];
  if (plain) {     // training forward / data gradient: "+ bias" only (warp-uniform branch)
#pragma unroll
    for (int g = 0; g < CHUNK / 4; ++g) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + c + 4 * g);
      w[2 * g] = pack_bf16x2(__uint_as_float(raw[4 * g]) + b4.x, __uint_as_float(raw[4 * g + 1]) + b4.y);
      w[2 * g + 1] = pack_bf16x2(__uint_as_float(raw[4 * g + 2]) + b4.z, __uint_as_float(raw[4 * g + 3]) + b4.w);
    }
  } else {
#pragma unroll
    for (int g = 0; g < CHUNK / 4; ++g) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + c + 4 * g);
      const float4 s4 = *reinterpret_cast<const float4*>(sb + 256 + c + 4 * g);
      w[2 * g] = pack_bf16x2(affine_act(raw[4 * g], s4.x, b4.x, slope), affine_act(raw[4 * g + 1], s4.y, b4.y, slope));
      w[2 * g + 1] = pack_bf16x2(affine_act(raw[4 * g + 2], s4.z, b4.z, slope), affine_act(raw[4 * g + 3], s4.w, b4.w, slope));
    }
  }
  if (dst == nullptr) return;
  if (c + CHUNK <= ncols) {
    if (row32) {
#pragma unroll
      for (int g = 0; g < CHUNK / 16; ++g) {
        uint32_t t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = w[8 * g + j];
        st_global_256(dst + c + 16 * g, t);
      }
    } else {
#pragma unroll
      for (int g = 0; g < CHUNK / 8; ++g) st_global_128(dst + c + 8 * g, w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
    }
  } else {
#pragma unroll
    for (int g = 0; g < CHUNK / 8; ++g)
      if (c + 8 * g + 8 <= ncols) st_global_128(dst + c + 8 * g, w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
  }
}

// `part` of `nparts` warps that share this TMEM lane quadrant: the 32-column chunks (and the 16-column tail) of the row are dealt
// round-robin to them; plain = "+ bias" only (scale == 1, slope == 1)
__device__ __forceinline__ void epilogue_row_bf16(uint32_t tmem_row, int block_n, const float* sbias, __nv_bfloat16* dst,
                                                  int ncols, bool row32, float slope, int part = 0, int nparts = 1,
                                                  bool plain = false) {
  int c = 0, ci = 0;
  for (; c + 32 <= block_n; c += 32, ++ci) {
    if (ci % nparts != part) continue;
    uint32_t raw[32];
    tmem_ld32(tmem_row + (uint32_t)c, raw);
    tmem_ld_wait();
    epilogue_store_chunk<32>(raw, sbias, dst, c, ncols, row32, slope, plain);
  }
  if (c < block_n && ci % nparts == part) {   // block_n is a multiple of 16
    uint32_t raw[16];
    tmem_ld16(tmem_row + (uint32_t)c, raw);
    tmem_ld_wait();
    epilogue_store_chunk<16>(raw, sbias, dst, c, ncols, row32, slope, plain);
  }
}


// ---- split-K reduction of weight-gradient tiles: 16 consecutive fp32 columns of one row, vector reds when aligned ----
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// adds raw[0..16) to dst[0..16), limited to the first `avail` columns; vec = dst is 16-byte aligned
__device__ __forceinline__ void red_add_row16(float* dst, const uint32_t (&raw)[16], int avail, bool vec) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (vec && 4 * g + 4 <= avail) {
      red_add_v4(dst + 4 * g, __uint_as_float(raw[4 * g]), __uint_as_float(raw[4 * g + 1]), __uint_as_float(raw[4 * g + 2]),
                 __uint_as_float(raw[4 * g + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * g + j < avail) atomicAdd(dst + 4 * g + j, __uint_as_float(raw[4 * g + j]));
    }
  }
}


// ---- host side: tensor maps -----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

inline int encode_bf16_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                           const cuuint64_t* strides_bytes, const cuuint32_t* box, const char* what,
                           const cuuint32_t* elem_strides = nullptr) {
  // cuTensorMapEncodeTiled is a DRIVER call: it needs the primary context to be current in the calling thread.  PyTorch's
  // autograd worker threads only get one through their first runtime call, so bind it here once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);
    ctx_bound = true;
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return NEXTOU_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides)   // strided traversal: the box spans box[i] source elements, ceil(box[i] / stride[i]) land in smem
    for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return NEXTOU_ERR_CUDA;
  }
  return 0;
}

inline int pick_block_n(int N) {
  const int tiles = (N + 255) / 256;
  int bn = (N + tiles - 1) / tiles;
  bn = (bn + 15) / 16 * 16;
  return bn < 16 ? 16 : bn;
}
inline int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}


}  // namespace nextou
