// DRAFT (not built, not run on a GPU yet — see drafts/README.md).
//
// Stride-1 'same' convolution with a 3x3 in-plane kernel (1x3x3 / 3x3x3 of StackedConvBlocks, reference ED:125-141, 281-300)
// for SMALL Cout (<= 80: the 33- and 66-channel layers), with the three kw taps FOLDED INTO THE N DIMENSION of the MMA.
//
// conv_halo_tcgen05_kernel issues one tcgen05.mma (M128 x Npad x K16) per (kd, kh, kw, K16 step); at Npad = 48 / 80 each costs
// 62-84 cycles (shared-memory operand reads + row-shifted A start) where the math needs 24 / 40.  Here
//     acc[r][(kw, co)] = sum_{kd, kh, ci}  X[d + kd - pd][h0 + i + kh - 1][w0 - 1 + j][ci] * W[co][kd][kh][kw][ci],   r = 8 i + j
// is ONE mma of N = 3 * Npad per (kd, kh, K16 step) on the UN-SHIFTED rows of the haloed box {64 ch, 8 (W), 16 + 2 (H)}: the kh
// tap moves the A start by 8 rows = 1 024 B (a whole swizzle atom: no shifted-start penalty), and
//     y[h0 + i][w0 + wl][co] = acc[8 i + wl][(0, co)] + acc[8 i + wl + 1][(1, co)] + acc[8 i + wl + 2][(2, co)],   wl = 0 .. 5
// is formed in the epilogue: accumulator row = TMEM lane = epilogue lane, so the +1 / +2 rows are a 1- / 2-lane shuffle inside
// the warp (8 i + wl + 2 <= 8 i + 7).  A CTA tile is 16 (H) x 6 (W) output voxels (every NexToU width is a multiple of 6).
//
// wfold: bf16 [3 * Npad][kd * 3 * cin_pad], row (kw * Npad + co), column ((kd_ * 3 + kh_) * cin_pad + ci), zero padded
// (Npad = Cout rounded up to 16, cin_pad = Cin rounded up to 64).  Everything else as conv_halo_tcgen05_kernel: persistent
// CTAs, TMA producer / MMA issuer / weight producer / 4 epilogue warps, double-buffered accumulator in tensor memory.
#include "../tc_common.cuh"

namespace nextou {

constexpr int CF_TH = 16, CF_BW = 8, CF_TW = 6;   // output rows, box columns, valid output columns per tile
constexpr int CF_THREADS = 224;
constexpr int CF_A_BYTES = CF_BW * (CF_TH + 2) * 128;   // 18 432 B: one haloed box (1 024-byte multiple)

struct ConvFoldParams {
  int N, Cin, npad, nfold, tmem_cols;
  int kblocks, kd, pd;
  int D, H, W, B, nh, nw;
  long long total_tiles;
  int a_stages, b_stages, b_resident;     // weight tiles: one per (kd, kh, cb); resident = all of them stay in smem
  void* C;
  long long ldc;
  const float* bias;
};

__device__ __forceinline__ void cf_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(CF_THREADS, 1)
    conv_fold_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const ConvFoldParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.nfold * 128;                       // one (kd, kh, cb) weight tile: nfold rows x 64 channels
  uint8_t* smA = smem;
  uint8_t* smB = smem + (size_t)p.a_stages * CF_A_BYTES;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smB + (size_t)p.b_stages * b_bytes);
  uint64_t* emptyA = fullA + p.a_stages;
  uint64_t* fullB = emptyA + p.a_stages;
  uint64_t* emptyB = fullB + p.b_stages;
  uint64_t* tmem_full = emptyB + p.b_stages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  __shared__ __align__(16) float sbias[96];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.npad; i += blockDim.x) sbias[i] = (p.bias != nullptr && i < p.N) ? p.bias[i] : 0.f;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int slabs = p.kd * p.kblocks;          // activation boxes per tile; each feeds 3 kh taps

  if (warp == 0) {
    // ---------------- activation boxes ----------------
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        long long t = tile;
        const int wt = (int)(t % p.nw); t /= p.nw;
        const int ht = (int)(t % p.nh); t /= p.nh;
        const int d0 = (int)(t % p.D);
        const int bn = (int)(t / p.D);
        int kd_ = 0, cb = 0;
        for (int s = 0; s < slabs; ++s) {
          mbar_wait(&emptyA[st], ph ^ 1);
          mbar_expect_tx(&fullA[st], (uint32_t)CF_A_BYTES);
          tma_load_5d(smA + (size_t)st * CF_A_BYTES, &tmA, &fullA[st], cb * 64, wt * CF_TW - 1, ht * CF_TH - 1, d0 + kd_ - p.pd, bn);
          if (++st == p.a_stages) { st = 0; ph ^= 1; }
          if (++cb == p.kblocks) { cb = 0; ++kd_; }
        }
      }
    }
  } else if (warp == 2) {
    // ---------------- weight tiles: one per (kd, cb, kh), in the order the MMA warp consumes them ----------------
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      const int cin_pad = p.kblocks * 64;
      bool first = true;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        if (p.b_resident && !first) break;
        first = false;
        int kd_ = 0, cb = 0;
        for (int s = 0; s < slabs; ++s) {
          for (int kh_ = 0; kh_ < 3; ++kh_) {
            mbar_wait(&emptyB[st], ph ^ 1);
            mbar_expect_tx(&fullB[st], (uint32_t)b_bytes);
            tma_load_2d(smB + (size_t)st * b_bytes, &tmB, &fullB[st], (kd_ * 3 + kh_) * cin_pad + cb * 64, 0);
            if (++st == p.b_stages) { st = 0; ph ^= 1; }
          }
          if (++cb == p.kblocks) { cb = 0; ++kd_; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (warp-uniform control flow, one elected lane issues) ----------------
    const uint32_t idesc = make_idesc_bf16(128, p.nfold);
    const int last_ksteps = (p.Cin - (p.kblocks - 1) * 64 + 15) / 16;
    int sa = 0, sb = 0, it = 0;
    uint32_t pa = 0, pb = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(acc * p.nfold);
      uint32_t accum = 0;
      int cb = 0;
      for (int s = 0; s < slabs; ++s) {
        mbar_wait(&fullA[sa], pa);
        const uint64_t ad0 = make_kmajor_sw128_desc(smem_u32(smA + (size_t)sa * CF_A_BYTES));
        const int ksteps = (cb == p.kblocks - 1) ? last_ksteps : 4;
        for (int kh_ = 0; kh_ < 3; ++kh_) {
          if (!(p.b_resident && it > 0)) mbar_wait(&fullB[sb], pb);
          tc_fence_after();
          const uint64_t bd0 = make_kmajor_sw128_desc(smem_u32(smB + (size_t)sb * b_bytes));
          if (elect_one()) {
            // the kh tap starts 8 box rows (= 1 024 B = 64 descriptor units) further: a whole swizzle atom, stays aligned
            for (int k = 0; k < ksteps; ++k) {
              umma_f16(tacc, ad0 + (uint64_t)(kh_ * 64 + 2 * k), bd0 + (uint64_t)(2 * k), idesc, accum);
              accum = 1;
            }
            if (!p.b_resident) umma_commit(&emptyB[sb]);
          }
          __syncwarp();
          accum = 1;
          if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
        }
        if (elect_one()) umma_commit(&emptyA[sa]);
        __syncwarp();
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        if (++cb == p.kblocks) cb = 0;
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);
      __syncwarp();
    }
  } else {
    // ---------------- epilogue: warps 3..6 -> TMEM lane quarters warp % 4; row R = 8 i + j of the tile ----------------
    const int q = warp & 3;
    const int R = q * 32 + lane;
    const int ih = R >> 3, j = R & 7;
    int it = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      long long t = tile;
      const int wt = (int)(t % p.nw); t /= p.nw;
      const int ht = (int)(t % p.nh); t /= p.nh;
      const int d0 = (int)(t % p.D);
      const int bn = (int)(t / p.D);
      const int h = ht * CF_TH + ih, w = wt * CF_TW + j;
      const bool valid = j < CF_TW && h < p.H && w < p.W;
      __nv_bfloat16* dst = valid ? reinterpret_cast<__nv_bfloat16*>(p.C) + ((((long long)bn * p.D + d0) * p.H + h) * p.W + w) * p.ldc
                                 : nullptr;
      const int acc = it & 1;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.nfold);
      for (int c = 0; c < p.npad; c += 16) {
        uint32_t g0[16], g1[16], g2[16];
        tmem_ld16(trow + (uint32_t)c, g0);
        tmem_ld16(trow + (uint32_t)(p.npad + c), g1);
        tmem_ld16(trow + (uint32_t)(2 * p.npad + c), g2);
        tmem_ld_wait();
        uint32_t wv[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float v0 = __uint_as_float(g0[e]) + __shfl_down_sync(0xffffffffu, __uint_as_float(g1[e]), 1) +
                     __shfl_down_sync(0xffffffffu, __uint_as_float(g2[e]), 2) + sbias[c + e];
          float v1 = __uint_as_float(g0[e + 1]) + __shfl_down_sync(0xffffffffu, __uint_as_float(g1[e + 1]), 1) +
                     __shfl_down_sync(0xffffffffu, __uint_as_float(g2[e + 1]), 2) + sbias[c + e + 1];
          wv[e >> 1] = pack_bf16x2(v0, v1);
        }
        if (dst != nullptr) {   // columns [N, ldc) come out as zeros (zero weight rows, zero bias)
          if (c + 8 <= p.ldc) st_global_128(dst + c, wv[0], wv[1], wv[2], wv[3]);
          if (c + 16 <= p.ldc) st_global_128(dst + c + 8, wv[4], wv[5], wv[6], wv[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) cf_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace nextou

using namespace nextou;

// Same contract as nextou_conv3d_ndhwc_halo_fwd (bf16 output only), restricted to kh = kw = 3 and Cout <= 80; wfold as above.
extern "C" int nextou_conv3d_ndhwc_fold_fwd(const void* x, long long ldx, int B, int D, int H, int W, int Cin,
                                            const void* wfold, int Cout, int kd, const float* bias, void* out, long long ldo,
                                            void* stream) {
  NEXTOU_REQUIRE(x && wfold && out, "conv3d_ndhwc_fold_fwd: null pointer");
  NEXTOU_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && Cout <= 80 && kd % 2 == 1 && kd <= 7,
                 "conv3d_ndhwc_fold_fwd: bad shape (Cout <= 80, odd kd)");
  NEXTOU_REQUIRE(ldx % 8 == 0 && ldx >= Cin && ldo % 8 == 0 && ldo >= Cout, "conv3d_ndhwc_fold_fwd: pitches must be multiples of 8");
  NEXTOU_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)wfold & 15) == 0 && ((uintptr_t)out & 15) == 0, "conv3d_ndhwc_fold_fwd: 16-byte alignment");
  ConvFoldParams p = {};
  p.N = Cout; p.Cin = Cin;
  p.npad = (Cout + 15) / 16 * 16;
  p.nfold = 3 * p.npad;
  p.tmem_cols = pow2_cols(2 * p.nfold);
  NEXTOU_REQUIRE(p.tmem_cols <= 512 && p.npad <= 96, "conv3d_ndhwc_fold_fwd: 3 * Cout_pad must fit 256 columns");
  p.kblocks = (Cin + 63) / 64;
  p.kd = kd; p.pd = kd / 2;
  p.D = D; p.H = H; p.W = W; p.B = B;
  p.nh = (H + CF_TH - 1) / CF_TH; p.nw = (W + CF_TW - 1) / CF_TW;
  p.total_tiles = (long long)B * D * p.nh * p.nw;
  p.C = out; p.ldc = ldo; p.bias = bias;
  const int b_bytes = p.nfold * 128;
  const int budget = 200 * 1024;
  const int ntiles_b = kd * p.kblocks * 3;
  p.a_stages = 3;
  if ((long long)ntiles_b * b_bytes <= budget - 3 * CF_A_BYTES && ntiles_b <= 32) {
    p.b_resident = 1; p.b_stages = ntiles_b;
    int as = (int)((budget - (long long)ntiles_b * b_bytes) / CF_A_BYTES);
    p.a_stages = as > 8 ? 8 : (as < 3 ? 3 : as);
  } else {
    p.b_resident = 0;
    int st = (budget - p.a_stages * CF_A_BYTES) / b_bytes;
    if (st > 6) st = 6;
    NEXTOU_REQUIRE(st >= 2, "conv3d_ndhwc_fold_fwd: weight tile does not fit");
    p.b_stages = st;
  }
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * W, (cuuint64_t)ldx * 2 * W * H, (cuuint64_t)ldx * 2 * W * H * D};
    cuuint32_t box[5] = {64, CF_BW, CF_TH + 2, 1, 1};
    int rc = encode_bf16_map(&tmA, x, 5, dims, str, box, "conv fold input");
    if (rc) return rc;
  }
  {
    const int cin_pad = p.kblocks * 64;
    cuuint64_t dims[2] = {(cuuint64_t)kd * 3 * cin_pad, (cuuint64_t)p.nfold};
    cuuint64_t str[1] = {(cuuint64_t)kd * 3 * cin_pad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)p.nfold};
    int rc = encode_bf16_map(&tmB, wfold, 2, dims, str, box, "conv fold weights");
    if (rc) return rc;
  }
  const size_t smem = 1024 + (size_t)p.a_stages * CF_A_BYTES + (size_t)p.b_stages * b_bytes +
                      (2 * p.a_stages + 2 * p.b_stages + 4) * sizeof(uint64_t) + 16;
  int rc = ensure_smem(conv_fold_tcgen05_kernel, smem);
  if (rc) return rc;
  long long ctas = num_sms();
  if (ctas > p.total_tiles) ctas = p.total_tiles;
  conv_fold_tcgen05_kernel<<<(unsigned)ctas, CF_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, p);
  return check_launch("conv_fold_tcgen05_kernel");
}
