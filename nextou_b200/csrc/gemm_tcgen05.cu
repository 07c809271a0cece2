// tcgen05 / TMEM / TMA GEMM engine for the dense contractions of the NexToU hot path (sm_100a only):
//   * pointwise (1x1) convolutions  out[T][Cout] = tok[T][Cin] * W[Cout][Cin]^T + b     (ED:305, 373-381, 710-720, 833-842)
//   * spatial convolutions as implicit GEMM over NDHWC activations (StackedConvBlocks, ED:125-141, 281-300):
//       M = output voxels (a td x th x tw brick of <= 128 voxels per CTA), N = Cout, K = taps x Cin.
//     The im2col gather never exists: for every tap a 5-D TMA box {64 ch, tw, th, td, 1} is fetched at the
//     tap-shifted coordinate, out-of-range coordinates are zero-filled by the TMA unit (= the conv's zero padding),
//     and the box lands in shared memory as 128-byte rows (one voxel x 64 channels) in the SWIZZLE_128B K-major
//     layout that tcgen05.mma consumes directly.
// Structure (one CTA = one 128 x BLOCK_N output tile, 192 threads):
//   warp 0   : TMA producer (one elected lane), kStages-deep mbarrier ring
//   warp 1   : MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16, bf16 x bf16 -> fp32 in TMEM
//   warps 2-5: epilogue: tcgen05.ld (32 lanes x 16 columns per instruction), + bias, -> bf16/fp32 global rows
// Accumulators never touch registers until the epilogue; operands never touch registers at all.
#include "tc_common.cuh"

namespace nextou {

// ------------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------------
constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;   // one 128-byte swizzle row of bf16
constexpr int GEMM_THREADS = 192;
constexpr int MAX_TAPS = 64;

struct GemmParams {
  // problem
  int M, N;            // output rows (tokens / voxels) and columns (Cout)
  int kblocks;         // 64-wide K blocks per tap (ceil(Cin / 64))
  int taps;            // 1 for a plain GEMM
  int block_n;         // UMMA N (multiple of 16, <= 256)
  int tmem_cols;       // power of two >= block_n
  int stages;
  // output
  void* C;
  long long ldc;       // row pitch of C in elements; columns [N, ldc) are written as zeros
  int out_dtype;
  const float* bias;   // [N] or NULL
  // implicit-GEMM geometry (is_conv)
  int is_conv;
  int D, H, W;         // the i-grid the CTAs tile (output volume of a forward conv; per-class sub-grid of a data gradient)
  int td, th, tw;      // brick of the i-grid per CTA (td*th*tw <= 128)
  int nd, nh, nw;      // bricks per axis
  // generalised tap geometry: tap t reads the input at coordinate i*es + tap_off[t] (es = element stride of the input
  // tensor map) and uses weight block tap_wi[t]; grid point i is written to output voxel i*os + oo of a [Do][Ho][Wo] volume
  int es_d, es_h, es_w;
  int os_d, os_h, os_w, oo_d, oo_h, oo_w;
  int Do, Ho, Wo;
  int last_ksteps;     // K16 steps of the last (ragged) 64-channel block
  int zero_fill;       // no taps at all: write bias / zeros
  int row32;           // every output row starts 32-byte aligned (256-bit stores)
  int store_cols;      // columns [0, store_cols) of an output row are written (<= ldc: the row may continue with other data)
  const float* scale;  // per-column scale [N] or NULL (= 1); out = lrelu(acc * scale + bias, slope)
  float slope;
  signed char tap_dd[MAX_TAPS], tap_dh[MAX_TAPS], tap_dw[MAX_TAPS];
  short tap_wi[MAX_TAPS];
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB] [stages][B block_n*128 B] [barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = GEMM_BM * GEMM_BK * 2;
  const int b_bytes = p.block_n * GEMM_BK * 2;
  uint8_t* smA = smem;
  uint8_t* smB = smem + (size_t)p.stages * a_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smB + (size_t)p.stages * b_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_base_holder = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  __shared__ __align__(16) float sbias[512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.block_n;
  stage_bias(sbias, p.bias, n0, p.N, p.block_n, p.scale);

  // output brick of this CTA
  int m0 = blockIdx.x * GEMM_BM;  // plain GEMM: first row
  int bn = 0, d0 = 0, h0 = 0, w0 = 0;
  if (p.is_conv) {
    int t = blockIdx.x;
    const int wt = t % p.nw; t /= p.nw;
    const int ht = t % p.nh; t /= p.nh;
    const int dt = t % p.nd; t /= p.nd;
    bn = t;
    d0 = dt * p.td; h0 = ht * p.th; w0 = wt * p.tw;
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_holder, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_holder;

  const int total_kb = p.taps * p.kblocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int a_tx = p.is_conv ? p.td * p.th * p.tw * GEMM_BK * 2 : a_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < total_kb; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], (uint32_t)(a_tx + b_bytes));
        const int tap = it / p.kblocks, cb = it - tap * p.kblocks;
        if (p.is_conv) {
          tma_load_5d(smA + (size_t)stage * a_bytes, &tmA, &full_bar[stage], cb * GEMM_BK, w0 * p.es_w + p.tap_dw[tap],
                      h0 * p.es_h + p.tap_dh[tap], d0 * p.es_d + p.tap_dd[tap], bn);
          tma_load_2d(smB + (size_t)stage * b_bytes, &tmB, &full_bar[stage], (p.tap_wi[tap] * p.kblocks + cb) * GEMM_BK, n0);
        } else {
          tma_load_2d(smA + (size_t)stage * a_bytes, &tmA, &full_bar[stage], cb * GEMM_BK, m0);
          tma_load_2d(smB + (size_t)stage * b_bytes, &tmB, &full_bar[stage], it * GEMM_BK, n0);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {  // whole warp: warp-uniform control flow (descriptors stay in uniform registers), one elected lane issues
      const uint32_t idesc = make_idesc_bf16(GEMM_BM, p.block_n);
      int stage = 0, cb = 0;
      uint32_t phase = 0;
      for (int it = 0; it < total_kb; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smA + (size_t)stage * a_bytes));
        const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smB + (size_t)stage * b_bytes));
        const int ksteps = (cb == p.kblocks - 1) ? p.last_ksteps : GEMM_BK / 16;   // skip all-zero K16 steps of the padding
        if (++cb == p.kblocks) cb = 0;
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128-byte swizzle row: +2 in the (>>4) start-address field
            umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (total_kb > 0 && elect_one()) umma_commit(tmem_full_bar);        // accumulator complete
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quarters warp%4) =====================
    if (!p.zero_fill) mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row inside the tile == TMEM lane
    long long out_row = -1;
    if (p.is_conv) {
      const int wx = r % p.tw, hy = (r / p.tw) % p.th, dz = r / (p.tw * p.th);
      const int d = d0 + dz, h = h0 + hy, w = w0 + wx;
      if (dz < p.td && d < p.D && h < p.H && w < p.W)
        out_row = (((long long)bn * p.Do + d * p.os_d + p.oo_d) * p.Ho + h * p.os_h + p.oo_h) * p.Wo + w * p.os_w + p.oo_w;
    } else if (m0 + r < p.M) {
      out_row = m0 + r;
    }
    if (p.out_dtype == NEXTOU_BF16 && !p.zero_fill) {
      __nv_bfloat16* dst = out_row >= 0 ? reinterpret_cast<__nv_bfloat16*>(p.C) + out_row * p.ldc + n0 : nullptr;
      const long long left = p.store_cols - n0;
      epilogue_row_bf16(tmem_base + ((uint32_t)(q * 32) << 16), p.block_n, sbias, dst,
                        (int)(left < p.block_n ? left : p.block_n), p.row32 != 0, p.slope);
    } else
    for (int c = 0; c < p.block_n; c += 16) {
      uint32_t raw[16];
      if (!p.zero_fill) {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, raw);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) raw[j] = 0u;
      }
      if (out_row >= 0) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = n0 + c + j;
          const float x = affine_act(raw[j], sbias[256 + c + j], sbias[c + j], p.slope);
          v[j] = col < p.N ? x : 0.f;
        }
        if (n0 + c < p.store_cols) {
          if (p.out_dtype == NEXTOU_BF16)
            store_chunk16(reinterpret_cast<__nv_bfloat16*>(p.C) + out_row * p.ldc, n0 + c, v, p.store_cols);
          else
            store_chunk16(reinterpret_cast<float*>(p.C) + out_row * p.ldc, n0 + c, v, p.store_cols);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, long long m_tiles, cudaStream_t st) {
  p.block_n = pick_block_n(p.N);
  p.tmem_cols = pow2_cols(p.block_n);
  p.row32 = (p.ldc % 16 == 0 && ((uintptr_t)p.C & 31) == 0) ? 1 : 0;
  if (p.store_cols <= 0 || p.store_cols > p.ldc) p.store_cols = (int)p.ldc;
  const int per_stage = GEMM_BM * GEMM_BK * 2 + p.block_n * GEMM_BK * 2;
  int stages = (200 * 1024) / per_stage;
  if (stages > 4) stages = 4;
  if (stages < 2) stages = 2;
  const int total_kb = p.taps * p.kblocks;
  if (stages > total_kb) stages = total_kb < 2 ? 2 : total_kb;
  p.stages = stages;
  const size_t smem = 1024 + (size_t)stages * per_stage + (2 * stages + 1) * sizeof(uint64_t) + 16;
  int rc = ensure_smem(gemm_tcgen05_kernel, smem);
  if (rc) return rc;
  const int n_tiles = (p.N + p.block_n - 1) / p.block_n;
  if (m_tiles > 2147483647LL || n_tiles > 65535) {
    set_error("gemm: grid too large");
    return NEXTOU_ERR_INVALID;
  }
  dim3 grid((unsigned)m_tiles, (unsigned)n_tiles);
  gemm_tcgen05_kernel<<<grid, GEMM_THREADS, smem, st>>>(tmA, tmB, p);
  return check_launch("gemm_tcgen05_kernel");
}

}  // namespace nextou

using namespace nextou;

// ======================================================================================================
// Persistent plain GEMM for the 1x1 layers: C[M][ldc] = A[M][K] * B[N][K]^T (+ bias).
// The 1x1 convolutions of NexToU have tiny K (132..1296) and N (14..1296): one 128-row tile is only 3..21 K blocks of MMA
// work, so a one-tile-per-CTA kernel is dominated by launch / TMEM-allocation / barrier-init / pipeline-fill overhead
// (measured 8.7x off the HBM roofline).  Here a CTA keeps its [block_n x K] weight slice RESIDENT in shared memory
// (loaded once), walks the row tiles m = blockIdx.x, blockIdx.x + gridDim.x, ..., streams only the activation tiles
// through a TMA ring and double-buffers the accumulator in tensor memory so the epilogue of tile i overlaps tile i+1.
// 320 threads: warp 0 TMA producer, warp 1 MMA issuer (warp-uniform, elected lane), warps 2-9 epilogue: the epilogue of a
// 128 x block_n tile is a latency chain per warp (TMEM load -> convert -> store), so two warps share each TMEM lane quadrant
// and take alternate 32-column chunks (ncu: 4 epilogue warps = 1 per scheduler ran the 86 016 x 132 -> 528 GEMM at 4 us per tile).
// ======================================================================================================
namespace nextou {

constexpr int PG_MAX_STAGES = 8;
constexpr int PG_THREADS = 320;   // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue (two per TMEM lane quadrant)

struct PGemmParams {
  int M, N, K;
  int kblocks, last_ksteps;
  int block_n, tmem_cols, a_stages;
  int b_resident;          // weights of this Cout tile stay in smem; else they share the ring stage with A
  long long m_tiles;
  void* C;
  long long ldc;
  int out_dtype;
  const float* bias;
  int row32;
  const float* scale;
  float slope;
  // scatter mode (kernel == stride transposed convolution, ED:273-276, 321): row m is INPUT voxel (b, d, h, w) of a [B][D][H][W]
  // grid; the N tile holds `cpt` output parity classes of `cpb` columns each (class t = (a, i, j) of the kd x kh x kw kernel), and
  // class t of row m is written to output voxel (d*kd + a, h*kh + i, w*kw + j) of the [B][D*kd][H*kh][W*kw] volume
  int scatter, cpt, cpb, cp_store;
  int D, H, W, kd, kh, kw;
};

__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(PG_THREADS, 1)
    gemm_pers_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const PGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = GEMM_BM * GEMM_BK * 2;
  const int b_bytes = p.block_n * GEMM_BK * 2;
  const int stage_bytes = p.b_resident ? a_bytes : a_bytes + b_bytes;
  uint8_t* ring = smem;
  uint8_t* smBres = smem + (size_t)p.a_stages * stage_bytes;                       // [kblocks][block_n x 64] when resident
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smBres + (p.b_resident ? (size_t)p.kblocks * b_bytes : 0));
  uint64_t* empty_bar = full_bar + p.a_stages;
  uint64_t* bres_bar = empty_bar + p.a_stages;
  uint64_t* tmem_full = bres_bar + 1;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  __shared__ __align__(16) float sbias[512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.block_n;
  if (p.scatter) {     // every class segment carries the same per-channel bias
    for (int i = threadIdx.x; i < p.block_n; i += blockDim.x) {
      const int co = i % p.cpb;
      sbias[i] = (p.bias != nullptr && co < p.N) ? p.bias[co] : 0.f;
      sbias[256 + i] = 1.f;
    }
  } else {
    stage_bias(sbias, p.bias, n0, p.N, p.block_n, p.scale);
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(bres_bar, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      // B tile: rows n0 .. n0 + block_n of a plain [N][K] operand, or (scatter) `cpt` classes x `cpb` rows of [class][Cout][K]
      const int bn1 = p.scatter ? 0 : n0, bn2 = p.scatter ? blockIdx.y * p.cpt : 0;
      if (p.b_resident) {
        mbar_expect_tx(bres_bar, (uint32_t)(p.kblocks * b_bytes));
        for (int kb = 0; kb < p.kblocks; ++kb) tma_load_3d(smBres + (size_t)kb * b_bytes, &tmB, bres_bar, kb * GEMM_BK, bn1, bn2);
      }
      int st = 0;
      uint32_t ph = 0;
      for (long long mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
        const int m0 = (int)(mt * GEMM_BM);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&empty_bar[st], ph ^ 1);
          mbar_expect_tx(&full_bar[st], (uint32_t)stage_bytes);
          tma_load_2d(ring + (size_t)st * stage_bytes, &tmA, &full_bar[st], kb * GEMM_BK, m0);
          if (!p.b_resident) tma_load_3d(ring + (size_t)st * stage_bytes + a_bytes, &tmB, &full_bar[st], kb * GEMM_BK, bn1, bn2);
          if (++st == p.a_stages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(GEMM_BM, p.block_n);
    if (p.b_resident) mbar_wait(bres_bar, 0);
    int st = 0, it = 0;
    uint32_t ph = 0;
    for (long long mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(acc * p.block_n);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + (size_t)st * stage_bytes);
        const uint64_t adesc = make_kmajor_sw128_desc(sa);
        const uint64_t bdesc = make_kmajor_sw128_desc(p.b_resident ? smem_u32(smBres + (size_t)kb * b_bytes) : sa + a_bytes);
        const int ksteps = (kb == p.kblocks - 1) ? p.last_ksteps : 4;
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k)
            umma_f16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[st]);
        }
        __syncwarp();
        if (++st == p.a_stages) { st = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;     // which of the two warps of this lane quadrant
    const bool plain = p.scale == nullptr && p.slope == 1.f;
    int it = 0;
    for (long long mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++it) {
      const long long row = mt * GEMM_BM + q * 32 + lane;
      const int acc = it & 1;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      if (p.scatter) {
        // input voxel of this row -> the output voxel of every class of the tile; cp_store columns per class (zero padded)
        long long t = row;
        const int w = (int)(t % p.W); t /= p.W;
        const int h = (int)(t % p.H); t /= p.H;
        const int d = (int)(t % p.D);
        const long long b = t / p.D;
        const int Ho = p.H * p.kh, Wo = p.W * p.kw;
        for (int tl = 0; tl < p.cpt; ++tl) {
          const int cls = blockIdx.y * p.cpt + tl;
          const int j = cls % p.kw, i = (cls / p.kw) % p.kh, a = cls / (p.kw * p.kh);
          const long long orow = ((b * (p.D * p.kd) + d * p.kd + a) * Ho + h * p.kh + i) * Wo + w * p.kw + j;
          __nv_bfloat16* dst = row < p.M ? reinterpret_cast<__nv_bfloat16*>(p.C) + orow * p.ldc : nullptr;
          epilogue_row_bf16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n + tl * p.cpb), p.cpb, sbias + tl * p.cpb,
                            dst, p.cp_store, false, p.slope, part, 2, plain);
        }
      } else if (p.out_dtype == NEXTOU_BF16) {
        __nv_bfloat16* dst = row < p.M ? reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + n0 : nullptr;
        const long long left = p.ldc - n0;
        epilogue_row_bf16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n), p.block_n, sbias, dst,
                          (int)(left < p.block_n ? left : p.block_n), p.row32 != 0, p.slope, part, 2, plain);
      } else
      for (int c = part * 16; c < p.block_n; c += 32) {
        uint32_t raw[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n + c), raw);
        tmem_ld_wait();
        if (row < p.M && n0 + c < p.ldc) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = n0 + c + j;
            const float x = affine_act(raw[j], sbias[256 + c + j], sbias[c + j], p.slope);
            v[j] = col < p.N ? x : 0.f;
          }
          if (p.out_dtype == NEXTOU_BF16)
            store_chunk16(reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc, n0 + c, v, p.ldc);
          else
            store_chunk16(reinterpret_cast<float*>(p.C) + row * p.ldc, n0 + c, v, p.ldc);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace nextou

// C[M][ldc] = A[M][K] * B[N][K]^T (+ bias): A, B bf16 with K contiguous (row pitches lda / ldb elements, multiples of 8,
// 16-byte aligned bases); C bf16 or fp32 with ldc % 8 == 0; columns [N, ldc) of C are zero-filled.
extern "C" int nextou_gemm_bf16_tn(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc,
                                   int M, int N, int K, const float* bias, int out_dtype, void* stream) {
  return nextou_gemm_bf16_tn_affine(A, lda, B, ldb, C, ldc, M, N, K, nullptr, bias, 1.f, out_dtype, stream);
}

// C = lrelu((A * B^T) * scale[N] + shift[N], slope): the inference form — an eval-mode BatchNorm (+ LeakyReLU) behind a
// 1x1 convolution is folded into the epilogue (scale = gamma * rsqrt(var + eps), shift = beta + (bias - mean) * scale).
extern "C" int nextou_gemm_bf16_tn_affine(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc,
                                          int M, int N, int K, const float* scale, const float* bias, float slope,
                                          int out_dtype, void* stream) {
  NEXTOU_REQUIRE(A && B && C, "gemm_bf16_tn: null pointer");
  NEXTOU_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16_tn: bad shape M=%d N=%d K=%d", M, N, K);
  NEXTOU_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 && lda >= K && ldb >= K && ldc >= N,
                 "gemm_bf16_tn: row pitches must be multiples of 8 elements and cover K / N (lda=%lld ldb=%lld ldc=%lld)", lda, ldb, ldc);
  NEXTOU_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)C & 15) == 0, "gemm_bf16_tn: 16-byte alignment");
  NEXTOU_REQUIRE(out_dtype == NEXTOU_BF16 || out_dtype == NEXTOU_F32, "gemm_bf16_tn: bad out dtype");
  PGemmParams p = {};
  p.M = M; p.N = N; p.K = K;
  p.kblocks = (K + GEMM_BK - 1) / GEMM_BK;
  p.last_ksteps = (K - (p.kblocks - 1) * GEMM_BK + 15) / 16;
  p.block_n = pick_block_n(N);
  p.tmem_cols = pow2_cols(2 * p.block_n);
  p.m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  p.C = C; p.ldc = ldc; p.out_dtype = out_dtype; p.bias = bias; p.scale = scale; p.slope = slope;
  p.row32 = (ldc % 16 == 0 && ((uintptr_t)C & 31) == 0) ? 1 : 0;
  const int a_bytes = GEMM_BM * GEMM_BK * 2, b_bytes = p.block_n * GEMM_BK * 2;
  const int budget = 200 * 1024;
  const long long bres = (long long)p.kblocks * b_bytes;
  p.b_resident = bres <= budget - 3 * a_bytes ? 1 : 0;
  {
    const int per_stage = p.b_resident ? a_bytes : a_bytes + b_bytes;
    int st = (int)((budget - (p.b_resident ? bres : 0)) / per_stage);
    if (st > PG_MAX_STAGES) st = PG_MAX_STAGES;
    NEXTOU_REQUIRE(st >= 2, "gemm_bf16_tn: tile does not fit in shared memory (N tile %d, K %d)", p.block_n, K);
    p.a_stages = st;
  }
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t str[1] = {(cuuint64_t)lda * 2};
    cuuint32_t box[2] = {GEMM_BK, GEMM_BM};
    int rc = encode_bf16_map(&tmA, A, 2, dims, str, box, "A");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, 1};
    cuuint64_t str[2] = {(cuuint64_t)ldb * 2, (cuuint64_t)ldb * 2 * N};
    cuuint32_t box[3] = {GEMM_BK, (cuuint32_t)p.block_n, 1};
    int rc = encode_bf16_map(&tmB, B, 3, dims, str, box, "B");
    if (rc) return rc;
  }
  const size_t smem = 1024 + (size_t)p.a_stages * (p.b_resident ? a_bytes : a_bytes + b_bytes) + (p.b_resident ? (size_t)bres : 0) +
                      (2 * p.a_stages + 5) * sizeof(uint64_t) + 16;
  int rc = ensure_smem(gemm_pers_tcgen05_kernel, smem);
  if (rc) return rc;
  const int n_tiles = (N + p.block_n - 1) / p.block_n;
  int per_sm = (int)((220 * 1024) / smem);
  if (per_sm > 512 / p.tmem_cols) per_sm = 512 / p.tmem_cols;
  if (per_sm < 1) per_sm = 1;
  // one resident wave: rounding UP here put 150 CTAs of a 3-tile GEMM on 148 SMs, i.e. a second wave for 2 CTAs (2x kernel time)
  long long ctas = ((long long)num_sms() * per_sm) / n_tiles;
  if (ctas > p.m_tiles) ctas = p.m_tiles;
  if (ctas < 1) ctas = 1;
  NEXTOU_REQUIRE(n_tiles <= 65535, "gemm_bf16_tn: grid too large");
  dim3 grid((unsigned)ctas, (unsigned)n_tiles);
  gemm_pers_tcgen05_kernel<<<grid, PG_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, p);
  return check_launch("gemm_pers_tcgen05_kernel");
}

// Forward of a kernel == stride transposed convolution (decoder up-sampling, NexToU_Encoder_Decoder.py:273-276, 321) as ONE
// persistent GEMM: every input voxel owns a disjoint kd x kh x kw block of output voxels, so
//   out[(d*kd + a, h*kh + i, w*kw + j)][co] = bias[co] + sum_ci x[(d, h, w)][ci] * W[ci][co][a][i][j]
// is x[M][Cin] times the [classes * Cout][Cin] operand, with the epilogue scattering the class segments of a row to their
// output voxels.  The input is read once (the per-class launches of nextou_conv3d_ndhwc_strided_dgrad re-read it per class).
// x: bf16 tokens [B*D*H*W][ldx]; wpack_t: bf16 [Cout][taps][cin_pad] (taps in (kd, kh, kw) order, cin_pad = ceil(Cin/64)*64: the
// Bt pack of nextou_pack_weight on the (Cin, Cout, *k) weight); out: bf16 [B*D*kd*H*kh*W*kw][ldo], columns [0, store_cols) of
// every row written (store_cols % 8 == 0, Cout <= store_cols <= ldo; [Cout, store_cols) zero-filled).
extern "C" int nextou_convtranspose_scatter_fwd_supported(int Cout, int store_cols, int kd, int kh, int kw) {
  const int cpb = (store_cols + 15) / 16 * 16;
  return Cout > 0 && store_cols % 8 == 0 && store_cols >= Cout && cpb <= 256 && kd >= 1 && kh >= 1 && kw >= 1 && kd * kh * kw <= 64;
}

extern "C" int nextou_convtranspose_scatter_fwd(const void* x, long long ldx, int B, int D, int H, int W, int Cin, const void* wpack_t,
                                                int Cout, int kd, int kh, int kw, const float* bias, void* out, long long ldo,
                                                int store_cols, void* stream) {
  NEXTOU_REQUIRE(x && wpack_t && out, "convtranspose_scatter_fwd: null pointer");
  NEXTOU_REQUIRE(nextou_convtranspose_scatter_fwd_supported(Cout, store_cols, kd, kh, kw), "convtranspose_scatter_fwd: unsupported layer");
  NEXTOU_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && ldx % 8 == 0 && ldx >= Cin && ldo % 8 == 0 && ldo >= store_cols,
                 "convtranspose_scatter_fwd: bad shape / pitches");
  NEXTOU_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)wpack_t & 15) == 0 && ((uintptr_t)out & 15) == 0, "convtranspose_scatter_fwd: 16-byte alignment");
  const long long M = (long long)B * D * H * W;
  NEXTOU_REQUIRE(M < 2147483647LL, "convtranspose_scatter_fwd: too many voxels");
  const int taps = kd * kh * kw;
  PGemmParams p = {};
  p.M = (int)M; p.N = Cout; p.K = Cin;
  p.kblocks = (Cin + GEMM_BK - 1) / GEMM_BK;
  p.last_ksteps = (Cin - (p.kblocks - 1) * GEMM_BK + 15) / 16;
  p.scatter = 1;
  p.cpb = (store_cols + 15) / 16 * 16;                 // tensor-memory columns / weight rows per class (UMMA N granularity)
  p.cp_store = store_cols;
  p.cpt = 1;
  for (int c = taps; c >= 1; --c)                       // most classes per tile with cpt | taps and cpt * cpb <= 256
    if (taps % c == 0 && c * p.cpb <= 256) { p.cpt = c; break; }
  p.block_n = p.cpt * p.cpb;
  p.tmem_cols = pow2_cols(2 * p.block_n);
  p.m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  p.C = out; p.ldc = ldo; p.out_dtype = NEXTOU_BF16; p.bias = bias; p.scale = nullptr; p.slope = 1.f; p.row32 = 0;
  p.D = D; p.H = H; p.W = W; p.kd = kd; p.kh = kh; p.kw = kw;
  const int a_bytes = GEMM_BM * GEMM_BK * 2, b_bytes = p.block_n * GEMM_BK * 2;
  const int budget = 200 * 1024;
  const long long bres = (long long)p.kblocks * b_bytes;
  p.b_resident = bres <= budget - 3 * a_bytes ? 1 : 0;
  {
    const int per_stage = p.b_resident ? a_bytes : a_bytes + b_bytes;
    int st = (int)((budget - (p.b_resident ? bres : 0)) / per_stage);
    if (st > PG_MAX_STAGES) st = PG_MAX_STAGES;
    NEXTOU_REQUIRE(st >= 2, "convtranspose_scatter_fwd: tile does not fit in shared memory");
    p.a_stages = st;
  }
  const int cin_pad = p.kblocks * GEMM_BK;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)M};
    cuuint64_t str[1] = {(cuuint64_t)ldx * 2};
    cuuint32_t box[2] = {GEMM_BK, GEMM_BM};
    int rc = encode_bf16_map(&tmA, x, 2, dims, str, box, "convtranspose x");
    if (rc) return rc;
  }
  {
    // [Cout][taps][cin_pad] viewed as (ci, co, class): a box {64, cpb, cpt} lands as rows (class, co) = the N axis of the tile;
    // rows co >= Cout are zero-filled by the TMA unit
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)taps};
    cuuint64_t str[2] = {(cuuint64_t)taps * cin_pad * 2, (cuuint64_t)cin_pad * 2};
    cuuint32_t box[3] = {GEMM_BK, (cuuint32_t)p.cpb, (cuuint32_t)p.cpt};
    int rc = encode_bf16_map(&tmB, wpack_t, 3, dims, str, box, "convtranspose weights");
    if (rc) return rc;
  }
  const size_t smem = 1024 + (size_t)p.a_stages * (p.b_resident ? a_bytes : a_bytes + b_bytes) + (p.b_resident ? (size_t)bres : 0) +
                      (2 * p.a_stages + 5) * sizeof(uint64_t) + 16;
  int rc = ensure_smem(gemm_pers_tcgen05_kernel, smem);
  if (rc) return rc;
  const int n_tiles = taps / p.cpt;
  int per_sm = (int)((220 * 1024) / smem);
  if (per_sm > 512 / p.tmem_cols) per_sm = 512 / p.tmem_cols;
  if (per_sm < 1) per_sm = 1;
  // one resident wave: rounding UP here put 150 CTAs of a 3-tile GEMM on 148 SMs, i.e. a second wave for 2 CTAs (2x kernel time)
  long long ctas = ((long long)num_sms() * per_sm) / n_tiles;
  if (ctas > p.m_tiles) ctas = p.m_tiles;
  if (ctas < 1) ctas = 1;
  dim3 grid((unsigned)ctas, (unsigned)n_tiles);
  gemm_pers_tcgen05_kernel<<<grid, PG_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, p);
  return check_launch("gemm_pers_tcgen05_kernel");
}

// ------------------------------------------------------------------------------------------------------
// General per-tap implicit GEMM: out[i*os + oo][:] = bias + sum_t  x[i*es + off_t][:] * Wblock[wi_t]^T   for i in an i-grid.
// Serves (a) strided forward convolutions (es = stride, os = 1), (b) the data gradient of a strided convolution and the
// forward of a kernel == stride transposed convolution, one launch per output parity class (es = 1, os = stride, oo = class).
// ------------------------------------------------------------------------------------------------------
namespace nextou {
struct TapList {
  int n = 0;
  int dd[MAX_TAPS], dh[MAX_TAPS], dw[MAX_TAPS], wi[MAX_TAPS];
};

static int launch_conv_general(const void* x, long long ldx, int B, int Di, int Hi, int Wi, int Cin, const void* wpack,
                               int w_taps_total, int Cout, const TapList& taps, int gd, int gh, int gw,   // i-grid
                               int es_d, int es_h, int es_w, int os_d, int os_h, int os_w, int oo_d, int oo_h, int oo_w,
                               int Do, int Ho, int Wo, const float* bias, void* out, long long ldo, int out_dtype,
                               cudaStream_t stream, int store_cols = 0, const float* scale = nullptr, float slope = 1.f) {
  GemmParams p = {};
  p.store_cols = store_cols;
  p.scale = scale; p.slope = slope;
  p.M = 0; p.N = Cout; p.kblocks = (Cin + GEMM_BK - 1) / GEMM_BK; p.taps = taps.n;
  p.last_ksteps = (Cin - (p.kblocks - 1) * GEMM_BK + 15) / 16;
  p.zero_fill = taps.n == 0 ? 1 : 0;
  p.C = out; p.ldc = ldo; p.out_dtype = out_dtype; p.bias = bias; p.is_conv = 1;
  p.D = gd; p.H = gh; p.W = gw;
  p.es_d = es_d; p.es_h = es_h; p.es_w = es_w;
  p.os_d = os_d; p.os_h = os_h; p.os_w = os_w; p.oo_d = oo_d; p.oo_h = oo_h; p.oo_w = oo_w;
  p.Do = Do; p.Ho = Ho; p.Wo = Wo;
  for (int t = 0; t < taps.n; ++t) {
    p.tap_dd[t] = (signed char)taps.dd[t]; p.tap_dh[t] = (signed char)taps.dh[t]; p.tap_dw[t] = (signed char)taps.dw[t];
    p.tap_wi[t] = (short)taps.wi[t];
  }
  // brick (td x th x tw <= 128 grid points) minimising the number of CTAs; ties -> the widest W extent (coalescing);
  // the TMA box of a strided read spans t*es source elements per axis (<= 256)
  int tw = 1, th = 1, td = 1;
  long long best = -1;
  for (int a = 1; a <= (gw < 128 ? gw : 128); ++a)
    for (int b = 1; a * b <= 128 && b <= gh; ++b) {
      int c = 128 / (a * b);
      if (c > gd) c = gd;
      if (a * es_w > 256 || b * es_h > 256 || c * es_d > 256) continue;
      const long long tiles = (long long)((gw + a - 1) / a) * ((gh + b - 1) / b) * ((gd + c - 1) / c);
      if (best < 0 || tiles < best || (tiles == best && a > tw && a <= 32)) {
        best = tiles; tw = a; th = b; td = c;
      }
    }
  p.tw = tw; p.th = th; p.td = td;
  p.nw = (gw + tw - 1) / tw; p.nh = (gh + th - 1) / th; p.nd = (gd + td - 1) / td;
  const int bn = pick_block_n(Cout);
  const int cin_pad = p.kblocks * GEMM_BK;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Di, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * Wi, (cuuint64_t)ldx * 2 * Wi * Hi,
                         (cuuint64_t)ldx * 2 * Wi * Hi * Di};
    cuuint32_t box[5] = {GEMM_BK, (cuuint32_t)(tw * es_w), (cuuint32_t)(th * es_h), (cuuint32_t)(td * es_d), 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)es_w, (cuuint32_t)es_h, (cuuint32_t)es_d, 1};
    int rc = encode_bf16_map(&tmA, x, 5, dims, str, box, "conv input", estr);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)w_taps_total * cin_pad, (cuuint64_t)Cout};
    cuuint64_t str[1] = {(cuuint64_t)w_taps_total * cin_pad * 2};
    cuuint32_t box[2] = {GEMM_BK, (cuuint32_t)bn};
    int rc = encode_bf16_map(&tmB, wpack, 2, dims, str, box, "conv weights");
    if (rc) return rc;
  }
  return launch_gemm(tmA, tmB, p, (long long)B * p.nd * p.nh * p.nw, stream);
}

static int conv_common_checks(const char* who, const void* x, const void* w, const void* out, long long ldx, int Cin,
                              long long ldo, int Cout, int out_dtype) {
  NEXTOU_REQUIRE(x && w && out, "%s: null pointer", who);
  NEXTOU_REQUIRE(Cin > 0 && Cout > 0, "%s: bad channel counts", who);
  NEXTOU_REQUIRE(ldx % 8 == 0 && ldx >= Cin && ldo % 8 == 0 && ldo >= Cout, "%s: pitches must be multiples of 8", who);
  NEXTOU_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)out & 15) == 0, "%s: 16-byte alignment", who);
  NEXTOU_REQUIRE(out_dtype == NEXTOU_BF16 || out_dtype == NEXTOU_F32, "%s: bad out dtype", who);
  return 0;
}
}  // namespace nextou

// Strided convolution forward (StackedConvBlocks down-sampling convs, ED:134-141): out[o] = b + sum_k x[o*s + k - p] W[k].
// x: bf16 NDHWC [B][Di][Hi][Wi][ldx]; wpack: bf16 [Cout][taps*cin_pad] (taps ordered (kd, kh, kw), cin_pad = ceil(Cin/64)*64);
// out: [B*Do*Ho*Wo][ldo] with Do = (Di + 2 pd - kd) / sd + 1 (likewise H, W); columns [Cout, ldo) are zero-filled.
// The strided gather is done by the TMA unit (tensor map with element strides), out-of-range taps are zero-filled.
extern "C" int nextou_conv3d_ndhwc_strided_fwd(const void* x, long long ldx, int B, int Di, int Hi, int Wi, int Cin,
                                               const void* wpack, int Cout, int kd, int kh, int kw, int sd, int sh, int sw,
                                               int pd, int ph, int pw, const float* bias, void* out, long long ldo,
                                               int out_dtype, void* stream) {
  return nextou_conv3d_ndhwc_strided_fwd_affine(x, ldx, B, Di, Hi, Wi, Cin, wpack, Cout, kd, kh, kw, sd, sh, sw, pd, ph, pw, nullptr,
                                                bias, 1.f, out, ldo, out_dtype, stream);
}

// out = lrelu(conv(x) * scale[Cout] + shift[Cout], slope): inference form with the eval-mode norm (+ LeakyReLU) folded in
extern "C" int nextou_conv3d_ndhwc_strided_fwd_affine(const void* x, long long ldx, int B, int Di, int Hi, int Wi, int Cin,
                                                      const void* wpack, int Cout, int kd, int kh, int kw, int sd, int sh,
                                                      int sw, int pd, int ph, int pw, const float* scale, const float* bias,
                                                      float slope, void* out, long long ldo, int out_dtype, void* stream) {
  int rc = conv_common_checks("conv3d_ndhwc_strided_fwd", x, wpack, out, ldx, Cin, ldo, Cout, out_dtype);
  if (rc) return rc;
  NEXTOU_REQUIRE(B > 0 && Di > 0 && Hi > 0 && Wi > 0, "conv3d_ndhwc_strided_fwd: bad shape");
  NEXTOU_REQUIRE(kd > 0 && kh > 0 && kw > 0 && kd * kh * kw <= MAX_TAPS, "conv3d_ndhwc_strided_fwd: at most %d taps", MAX_TAPS);
  NEXTOU_REQUIRE(sd >= 1 && sh >= 1 && sw >= 1 && sd <= 8 && sh <= 8 && sw <= 8, "conv3d_ndhwc_strided_fwd: strides must be in [1, 8]");
  NEXTOU_REQUIRE(pd >= 0 && ph >= 0 && pw >= 0 && pd < 64 && ph < 64 && pw < 64, "conv3d_ndhwc_strided_fwd: bad padding");
  const int Do = (Di + 2 * pd - kd) / sd + 1, Ho = (Hi + 2 * ph - kh) / sh + 1, Wo = (Wi + 2 * pw - kw) / sw + 1;
  NEXTOU_REQUIRE(Do > 0 && Ho > 0 && Wo > 0, "conv3d_ndhwc_strided_fwd: empty output");
  TapList taps;
  for (int a = 0; a < kd; ++a)
    for (int b = 0; b < kh; ++b)
      for (int c = 0; c < kw; ++c) {
        const int t = taps.n++;
        taps.dd[t] = a - pd; taps.dh[t] = b - ph; taps.dw[t] = c - pw; taps.wi[t] = t;
      }
  return launch_conv_general(x, ldx, B, Di, Hi, Wi, Cin, wpack, kd * kh * kw, Cout, taps, Do, Ho, Wo, sd, sh, sw, 1, 1, 1, 0, 0,
                             0, Do, Ho, Wo, bias, out, ldo, out_dtype, (cudaStream_t)stream, 0, scale, slope);
}

// Stride-1 'same' convolution (odd kernels): the per-tap variant of nextou_conv3d_ndhwc_halo_fwd.
extern "C" int nextou_conv3d_ndhwc_fwd(const void* x, long long ldx, int B, int D, int H, int W, int Cin,
                                       const void* wpack, int Cout, int kd, int kh, int kw, const float* bias,
                                       void* out, long long ldo, int out_dtype, void* stream) {
  NEXTOU_REQUIRE(kd % 2 == 1 && kh % 2 == 1 && kw % 2 == 1, "conv3d_ndhwc_fwd: odd kernel sizes only");
  return nextou_conv3d_ndhwc_strided_fwd(x, ldx, B, D, H, W, Cin, wpack, Cout, kd, kh, kw, 1, 1, 1, kd / 2, kh / 2, kw / 2, bias,
                                         out, ldo, out_dtype, stream);
}

// Data gradient of a strided convolution, and (pd = ph = pw = 0, kernel == stride) the FORWARD of a transposed
// convolution (decoder up-sampling, ED:273-276, 321):
//   dx[u][ci] = bias[ci] + sum_{(v, k): v*s + k - p == u} dy[v][:] . wpack_t[ci][k][:]
// dy: bf16 NDHWC [B][Do][Ho][Wo][ldy] (Cout channels); wpack_t: bf16 [Cin][taps*cout_pad] (taps in (kd, kh, kw) order, NOT
// flipped); dx: [B*Di*Hi*Wi][ldx].  One launch per output parity class u mod s: each class is a small stride-1 convolution
// of dy whose results are written to every s-th voxel of dx; a class without taps (k < s) is written as bias / zeros.
extern "C" int nextou_conv3d_ndhwc_strided_dgrad(const void* dy, long long ldy, int B, int Do, int Ho, int Wo, int Cout,
                                                 const void* wpack_t, int Cin, int kd, int kh, int kw, int sd, int sh, int sw,
                                                 int pd, int ph, int pw, const float* bias, void* dx, long long ldx, int Di,
                                                 int Hi, int Wi, int out_dtype, void* stream) {
  return nextou_conv3d_ndhwc_strided_dgrad_cols(dy, ldy, B, Do, Ho, Wo, Cout, wpack_t, Cin, kd, kh, kw, sd, sh, sw, pd, ph, pw, bias,
                                                dx, ldx, (int)ldx, Di, Hi, Wi, out_dtype, stream);
}

// Same, writing only columns [0, store_cols) of every dx row (store_cols % 8 == 0, Cin <= store_cols <= ldx; columns
// [Cin, store_cols) are zero-filled): the decoder writes the up-sampled half of torch.cat((up, skip), 1) (ED:322) straight
// into the concatenation buffer, whose rows continue with the skip channels.
extern "C" int nextou_conv3d_ndhwc_strided_dgrad_cols(const void* dy, long long ldy, int B, int Do, int Ho, int Wo, int Cout,
                                                      const void* wpack_t, int Cin, int kd, int kh, int kw, int sd, int sh,
                                                      int sw, int pd, int ph, int pw, const float* bias, void* dx,
                                                      long long ldx, int store_cols, int Di, int Hi, int Wi, int out_dtype,
                                                      void* stream) {
  NEXTOU_REQUIRE(store_cols % 8 == 0 && store_cols >= Cin && store_cols <= ldx, "conv3d_ndhwc_strided_dgrad: bad store_cols %d", store_cols);
  int rc = conv_common_checks("conv3d_ndhwc_strided_dgrad", dy, wpack_t, dx, ldy, Cout, ldx, Cin, out_dtype);
  if (rc) return rc;
  NEXTOU_REQUIRE(B > 0 && Do > 0 && Ho > 0 && Wo > 0 && Di > 0 && Hi > 0 && Wi > 0, "conv3d_ndhwc_strided_dgrad: bad shape");
  NEXTOU_REQUIRE(kd > 0 && kh > 0 && kw > 0 && kd * kh * kw <= MAX_TAPS, "conv3d_ndhwc_strided_dgrad: at most %d taps", MAX_TAPS);
  NEXTOU_REQUIRE(sd >= 1 && sh >= 1 && sw >= 1 && sd <= 8 && sh <= 8 && sw <= 8, "conv3d_ndhwc_strided_dgrad: strides must be in [1, 8]");
  NEXTOU_REQUIRE(pd >= 0 && ph >= 0 && pw >= 0 && pd < 64 && ph < 64 && pw < 64, "conv3d_ndhwc_strided_dgrad: bad padding");
  for (int od = 0; od < sd && od < Di; ++od)
    for (int oh = 0; oh < sh && oh < Hi; ++oh)
      for (int ow = 0; ow < sw && ow < Wi; ++ow) {
        TapList taps;
        for (int a = 0; a < kd; ++a) {
          const int ta = od + pd - a;
          if (((ta % sd) + sd) % sd) continue;
          for (int b = 0; b < kh; ++b) {
            const int tb = oh + ph - b;
            if (((tb % sh) + sh) % sh) continue;
            for (int c = 0; c < kw; ++c) {
              const int tc = ow + pw - c;
              if (((tc % sw) + sw) % sw) continue;
              const int t = taps.n++;
              taps.dd[t] = ta / sd; taps.dh[t] = tb / sh; taps.dw[t] = tc / sw;
              taps.wi[t] = (a * kh + b) * kw + c;
            }
          }
        }
        const int gd = (Di - od + sd - 1) / sd, gh = (Hi - oh + sh - 1) / sh, gw = (Wi - ow + sw - 1) / sw;
        rc = launch_conv_general(dy, ldy, B, Do, Ho, Wo, Cout, wpack_t, kd * kh * kw, Cin, taps, gd, gh, gw, 1, 1, 1, sd, sh, sw,
                                 od, oh, ow, Di, Hi, Wi, bias, dx, ldx, out_dtype, (cudaStream_t)stream, store_cols);
        if (rc) return rc;
      }
  return 0;
}

// ======================================================================================================
// Weight gradient:  dW[co][tap][ci] += sum_v dY[v][co] * X[v + tap - pad][ci]
// GEMM view: M = Cout (128 per CTA), N = Cin tile (<= 256), K = voxels.  The voxel axis is the row axis of both
// token-major operands, so both are MN-MAJOR tcgen05 operands: a TMA box {64 channels, 64-voxel brick} lands as
// 64 rows (K) of 128 B (64 channels of M or N) — the canonical SWIZZLE_128B MN-major atom.  Per 64-voxel K block the
// dY brick is fetched once and multiplied with the tap-shifted X bricks of a whole tap group, each tap owning its
// own fp32 accumulator slab in tensor memory; the K axis is split over CTAs and the partial dW tiles are reduced
// with fp32 global atomics (dW is zero-filled by the caller).
// ======================================================================================================
namespace nextou {

constexpr int WG_BRICK = 64;               // voxels per K block
constexpr int WG_BOX_BYTES = WG_BRICK * 128;  // one {64 ch x 64 voxel} box
constexpr int WG_THREADS = 192;

struct WgradParams {
  int Cout, Cin;
  int D, H, W, B;
  int td, th, tw, nd, nh, nw;     // 64-voxel brick and bricks per axis
  int kd, kh, kw, pd, ph, pw, taps;
  int sd, sh, sw;                 // X is read at brick_origin * s + tap - pad (strided convolution / transposed convolution)
  int n_tile;                     // N per CTA (multiple of 16, <= 256)
  int n_boxes;                    // ceil(n_tile / 64)
  int tap_group;                  // taps per CTA (tap_group * n_tile <= 512 TMEM columns)
  int n_groups, n_mtiles, n_ntiles, ksplit;
  int tmem_cols, stages;
  long long total_bricks;
  float* dW;                      // [Cout][taps][cin_stride]
  int cin_stride;
};

// MN-major SWIZZLE_128B descriptor: 64-element (128 B) MN blocks `lbo` bytes apart, 8-row K groups 1024 B apart
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
    wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                         const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = (2 + p.tap_group * p.n_boxes) * WG_BOX_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_base_holder = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;   // grid = (tiles, ksplit): the tap groups / channel tiles of one brick range run together and share L2
  const int nt = t % p.n_ntiles; t /= p.n_ntiles;
  const int mt = t % p.n_mtiles; t /= p.n_mtiles;
  const int grp = t;
  const int m0 = mt * 128, n0 = nt * p.n_tile;
  const int tap0 = grp * p.tap_group;
  const int ntaps = min(p.tap_group, p.taps - tap0);
  const int ks = blockIdx.y;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_holder, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_holder;

  // bricks of this K split: ks, ks + ksplit, ...
  const long long nb = p.total_bricks > ks ? (p.total_bricks - ks + p.ksplit - 1) / p.ksplit : 0;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long i = 0; i < nb; ++i) {
        long long b = ks + i * p.ksplit;
        const int wt = (int)(b % p.nw); b /= p.nw;
        const int ht = (int)(b % p.nh); b /= p.nh;
        const int dt = (int)(b % p.nd); b /= p.nd;
        const int bn = (int)b;
        const int w0 = wt * p.tw, h0 = ht * p.th, d0 = dt * p.td;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], (uint32_t)((2 + ntaps * p.n_boxes) * WG_BOX_BYTES));
        uint8_t* st = smem + (size_t)stage * stage_bytes;
        tma_load_5d(st, &tmDY, &full_bar[stage], m0, w0, h0, d0, bn);
        tma_load_5d(st + WG_BOX_BYTES, &tmDY, &full_bar[stage], m0 + 64, w0, h0, d0, bn);
        for (int tp = 0; tp < ntaps; ++tp) {
          const int tap = tap0 + tp;
          const int kw_ = tap % p.kw, kh_ = (tap / p.kw) % p.kh, kd_ = tap / (p.kw * p.kh);
          for (int j = 0; j < p.n_boxes; ++j)
            tma_load_5d(st + (size_t)(2 + tp * p.n_boxes + j) * WG_BOX_BYTES, &tmX, &full_bar[stage], n0 + 64 * j,
                        w0 * p.sw + kw_ - p.pw, h0 * p.sh + kh_ - p.ph, d0 * p.sd + kd_ - p.pd, bn);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // warp-uniform control flow, one elected lane issues (descriptors stay in uniform registers)
    if (nb > 0) {
      // D fp32, A/B bf16, both MN-major (bits 15, 16)
      const uint32_t idesc = make_idesc_bf16(128, p.n_tile) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0;
      for (long long i = 0; i < nb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)stage * stage_bytes);
        if (elect_one()) {
          for (int tp = 0; tp < ntaps; ++tp) {
            const uint32_t xb = st + (uint32_t)(2 + tp * p.n_boxes) * WG_BOX_BYTES;
#pragma unroll
            for (int k = 0; k < WG_BRICK / 16; ++k) {
              // 16 voxels = 16 rows of 128 B = 2048 B along K
              const uint64_t adesc = make_mnmajor_sw128_desc(st + k * 2048, WG_BOX_BYTES);
              const uint64_t bdesc = make_mnmajor_sw128_desc(xb + k * 2048, WG_BOX_BYTES);
              umma_f16(tmem_base + (uint32_t)(tp * p.n_tile), adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(tmem_full_bar);
      __syncwarp();
    }
  } else if (nb > 0) {
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int co = m0 + q * 32 + lane;
    const bool vec = (p.cin_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dW) & 15) == 0;
    for (int tp = 0; tp < ntaps; ++tp) {
      for (int c = 0; c < p.n_tile; c += 16) {
        uint32_t raw[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tp * p.n_tile + c), raw);
        tmem_ld_wait();
        if (co < p.Cout)   // columns in [Cin, cin_stride) receive exact zeros (X channels beyond Cin are zero-filled by TMA)
          red_add_row16(p.dW + ((long long)co * p.taps + (tap0 + tp)) * p.cin_stride + n0 + c, raw, p.cin_stride - n0 - c, vec);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace nextou

// General weight gradient:  dW[m][tap][n] += sum_i dy[i][m] * x[i*s + tap - pad][n]   (fp32, the caller zero-fills
// dW[Cout][taps][cin_stride]).  dy: bf16 tokens of the DENSE grid [B][D][H][W][ldy] (Cout channels = M); x: bf16 tokens of
// the strided-read volume [B][Dx][Hx][Wx][ldx] (Cin channels = N).  Strided convolution: dy = output gradient, x = input.
// Transposed convolution (kernel == stride, pad 0): dy := the layer INPUT, x := the output gradient, dW = [Cin_t][tap][Cout_t].
extern "C" int nextou_conv3d_ndhwc_strided_wgrad(const void* dy, long long ldy, const void* x, long long ldx, int B, int D,
                                                 int H, int W, int Dx, int Hx, int Wx, int Cin, int Cout, int kd, int kh,
                                                 int kw, int sd, int sh, int sw, int pd, int ph, int pw, float* dW,
                                                 int cin_stride, void* stream) {
  NEXTOU_REQUIRE(dy && x && dW, "conv3d_ndhwc_wgrad: null pointer");
  NEXTOU_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && Dx > 0 && Hx > 0 && Wx > 0 && Cin > 0 && Cout > 0 && cin_stride >= Cin,
                 "conv3d_ndhwc_wgrad: bad shape");
  NEXTOU_REQUIRE(kd > 0 && kh > 0 && kw > 0 && sd >= 1 && sh >= 1 && sw >= 1 && sd <= 4 && sh <= 4 && sw <= 4 && pd >= 0 &&
                 ph >= 0 && pw >= 0, "conv3d_ndhwc_wgrad: bad kernel / stride / padding");
  NEXTOU_REQUIRE(ldy % 8 == 0 && ldy >= Cout && ldx % 8 == 0 && ldx >= Cin, "conv3d_ndhwc_wgrad: pitches must be multiples of 8");
  NEXTOU_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0, "conv3d_ndhwc_wgrad: 16-byte alignment");
  WgradParams p = {};
  p.Cout = Cout; p.Cin = Cin; p.D = D; p.H = H; p.W = W; p.B = B;
  p.kd = kd; p.kh = kh; p.kw = kw; p.pd = pd; p.ph = ph; p.pw = pw; p.taps = kd * kh * kw;
  p.sd = sd; p.sh = sh; p.sw = sw;
  p.dW = dW; p.cin_stride = cin_stride;
  // 64-voxel brick with power-of-two edges minimising the brick count (ties: widest along W)
  long long best = -1;
  for (int a = 64; a >= 1; a >>= 1)
    for (int b = 64 / a; b >= 1; b >>= 1) {
      const int c = 64 / (a * b);
      const long long bricks = (long long)((W + a - 1) / a) * ((H + b - 1) / b) * ((D + c - 1) / c);
      if (best < 0 || bricks < best) { best = bricks; p.tw = a; p.th = b; p.td = c; }
    }
  p.nw = (W + p.tw - 1) / p.tw; p.nh = (H + p.th - 1) / p.th; p.nd = (D + p.td - 1) / p.td;
  p.total_bricks = (long long)B * p.nd * p.nh * p.nw;
  p.n_ntiles = (Cin + 255) / 256;
  p.n_tile = ((Cin + p.n_ntiles - 1) / p.n_ntiles + 15) / 16 * 16;
  p.n_boxes = (p.n_tile + 63) / 64;
  p.n_mtiles = (Cout + 127) / 128;
  // taps per CTA: TMEM (512 columns) and shared memory (2 stages) limits
  int tg = 512 / p.n_tile;
  const int smem_budget = 200 * 1024;
  while (tg > 1 && 2 * (2 + tg * p.n_boxes) * WG_BOX_BYTES > smem_budget) --tg;
  if (tg > p.taps) tg = p.taps;
  NEXTOU_REQUIRE(tg >= 1 && 2 * (2 + tg * p.n_boxes) * WG_BOX_BYTES <= smem_budget, "conv3d_ndhwc_wgrad: tile does not fit");
  p.tap_group = tg;
  p.n_groups = (p.taps + tg - 1) / tg;
  p.tmem_cols = pow2_cols(tg * p.n_tile);
  const long long tiles = (long long)p.n_groups * p.n_mtiles * p.n_ntiles;
  long long ksplit = (2LL * num_sms() + tiles - 1) / tiles;
  if (ksplit > p.total_bricks / 24) ksplit = p.total_bricks / 24;   // >= 24 K blocks per CTA amortise prologue + reduction
  if (ksplit < 1) ksplit = 1;
  p.ksplit = (int)ksplit;
  p.stages = 2;   // small stages keep several CTAs resident per SM (measured better than a deeper ring here)
  CUtensorMap tmDY, tmX;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)ldy * 2, (cuuint64_t)ldy * 2 * W, (cuuint64_t)ldy * 2 * W * H,
                         (cuuint64_t)ldy * 2 * W * H * D};
    cuuint32_t box[5] = {64, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.td, 1};
    int rc = encode_bf16_map(&tmDY, dy, 5, dims, str, box, "wgrad dY");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)Wx, (cuuint64_t)Hx, (cuuint64_t)Dx, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * Wx, (cuuint64_t)ldx * 2 * Wx * Hx,
                         (cuuint64_t)ldx * 2 * Wx * Hx * Dx};
    cuuint32_t box[5] = {64, (cuuint32_t)(p.tw * sw), (cuuint32_t)(p.th * sh), (cuuint32_t)(p.td * sd), 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)sw, (cuuint32_t)sh, (cuuint32_t)sd, 1};
    int rc = encode_bf16_map(&tmX, x, 5, dims, str, box, "wgrad X", estr);
    if (rc) return rc;
  }
  const size_t smem = 1024 + (size_t)p.stages * (2 + tg * p.n_boxes) * WG_BOX_BYTES + (2 * p.stages + 1) * sizeof(uint64_t) + 16;
  int rc = ensure_smem(wgrad_tcgen05_kernel, smem);
  if (rc) return rc;
  NEXTOU_REQUIRE(tiles <= 65535, "conv3d_ndhwc_wgrad: too many tiles");
  dim3 grid((unsigned)tiles, (unsigned)p.ksplit);
  wgrad_tcgen05_kernel<<<grid, WG_THREADS, smem, (cudaStream_t)stream>>>(tmDY, tmX, p);
  return check_launch("wgrad_tcgen05_kernel");
}

// Stride-1 'same' convolution (odd kernels) / 1x1 layer (kd = kh = kw = 1): see nextou_conv3d_ndhwc_strided_wgrad.
extern "C" int nextou_conv3d_ndhwc_wgrad(const void* dy, long long ldy, const void* x, long long ldx, int B, int D, int H,
                                         int W, int Cin, int Cout, int kd, int kh, int kw, float* dW, int cin_stride,
                                         void* stream) {
  NEXTOU_REQUIRE(kd % 2 == 1 && kh % 2 == 1 && kw % 2 == 1, "conv3d_ndhwc_wgrad: odd kernel sizes only");
  return nextou_conv3d_ndhwc_strided_wgrad(dy, ldy, x, ldx, B, D, H, W, D, H, W, Cin, Cout, kd, kh, kw, 1, 1, 1, kd / 2, kh / 2,
                                           kw / 2, dW, cin_stride, stream);
}
