// Stride-1 'same' convolution with kh, kw in {1, 3} (every StackedConvBlocks convolution of NexToU that does not
// down-sample: 1x3x3 and 3x3x3, reference ED:125-141, 281-300) as a persistent HALO-REUSE implicit GEMM for sm_100a.
//
// csrc/gemm_tcgen05.cu fetches the activation brick once per tap (27 TMA boxes of 16 KB per 64-channel slab).  Here a CTA
// owns 16 (H) x 8 (W) output bricks of one depth slice and fetches, per depth tap and 64-channel slab, ONE haloed box
// {64 ch, 8+2, 16+2, 1} (23 KB).  All 9 in-plane taps read that box in place: the 8 voxels of an output row are 8
// consecutive 128-byte rows of the box, so tap (kh, kw) is the same SWIZZLE_128B K-major operand with its start address
// moved by (kh*10 + kw) rows and its 8-row group stride (SBO) set to the box row pitch of 10 rows = 1280 B (the swizzle
// is a function of the absolute smem address, so the shifted start stays consistent with what TMA wrote).
//
// Measured on B200 (tools/mma_rate.py): one tcgen05.mma M128 x N x K16 costs max(46, N/2) cycles, a ready mbarrier
// try_wait 100-200 cycles in the issuing thread.  The kernel is therefore organised to wait rarely:
//   * persistent CTAs walk many bricks; the accumulator is double buffered in tensor memory (epilogue of brick i
//     overlaps the MMAs of brick i+1);
//   * weights are RESIDENT in shared memory when the whole packed filter fits (the 33/66-channel layers: 54-108 KB),
//     otherwise they stream through a ring of tap groups (3 taps = one kernel row per barrier);
//   * all-zero K slabs of the channel padding are skipped (Cin = 33 needs 3 of the 4 K16 steps of its 64-wide slab).
//
// 224 threads: warp 0 activation-box producer, warp 1 MMA issuer, warp 2 weight producer (+ TMEM allocation),
// warps 3-6 epilogue (TMEM -> registers -> + bias -> global rows).
#include "tc_common.cuh"

namespace nextou {

constexpr int CV_TH = 16, CV_TW = 8;      // output brick (rows x cols) = 128 voxels = UMMA M
constexpr int CV_THREADS = 224;
constexpr int CV_A_MAX_STAGES = 8;
constexpr int CV_B_MAX_STAGES = 32;

struct ConvParams {
  int N;                 // Cout
  int Cin;
  int block_n, tmem_cols;
  int kblocks;           // 64-channel slabs of Cin
  int kd, kh, kw, pd, ph, pw;
  int D, H, W, B;
  int nh, nw;            // bricks per slice
  long long total_tiles;
  int box_w, box_rows;   // haloed box: (8 + kw - 1) wide, box_w * (16 + kh - 1) rows
  int a_stage_bytes;     // box_rows * 128 rounded up to 1024
  int a_stages;          // depth of the activation-box ring
  int b_group;           // in-plane taps per weight stage (one mbarrier per group)
  int b_groups;          // groups per slab = ceil(kh*kw / b_group)
  int b_stages;          // depth of the weight ring (== all groups of all slabs when resident)
  int b_resident;        // weights loaded once per CTA and kept
  int halves;            // 1: 16 x 8 bricks; 2: 16 x 16 bricks = two M = 128 accumulators that share every streamed weight tile
  void* C;
  long long ldc;
  int out_dtype;
  const float* bias;
  const float* scale;    // per-column scale [N] or NULL; out = lrelu(acc * scale + bias, slope)
  float slope;
  int row32;             // every output row starts 32-byte aligned (256-bit stores)
  long long* dbg;        // optional [16] cycle counters written by CTA (0,0); NULL in production
};

// K-major SWIZZLE_128B descriptor whose start may sit on any 128-byte row of a 1024-byte-aligned tile and whose 8-row
// groups are `sbo_bytes` apart (any multiple of 128).  Measured on B200: the UMMA unit derives the swizzle XOR from the
// ABSOLUTE shared-memory address bits (like TMA does when it writes the box), so a row-shifted start needs no
// correction — the base-offset field (bits 49..51) must stay 0 (filling it with (addr >> 7) & 7 gives wrong results).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc_rows(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Issue the MMAs of NT consecutive in-plane taps of a 3x3 kernel (NT = 9: the whole plane, 3: one kernel row, 1: one tap)
// x KS K16-steps with compile-time descriptor deltas: per MMA two 64-bit adds + the tcgen05.mma, nothing else.
// adesc0 / bdesc0 = descriptors of the group's first tap at k = 0; bstep = one tap's weight tile in 16-byte units.
template <int KS, int NT, int HALVES>
__device__ __forceinline__ void issue_group_3x3(uint32_t tacc, uint32_t block_n, uint64_t adesc0, uint64_t bdesc0, uint32_t bstep,
                                                uint32_t idesc, uint32_t accum_first) {
  constexpr int BOXW = CV_TW * HALVES + 2;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const uint32_t aoff = (uint32_t)((j / 3) * BOXW + (j % 3)) * 8u;   // (kh * box_w + kw) rows of 128 B, in 16-byte units
#pragma unroll
    for (int h = 0; h < HALVES; ++h)       // the second half of a wide brick starts 8 voxels (8 rows of 128 B) further along W
#pragma unroll
      for (int k = 0; k < KS; ++k)
        umma_f16(tacc + (uint32_t)h * block_n, adesc0 + (uint64_t)(aoff + 64 * h + 2 * k), bdesc0 + (uint64_t)(j * bstep + 2 * k), idesc,
                 (j | k) != 0 ? 1u : accum_first);
  }
}
template <int KS, int HALVES>
__device__ __forceinline__ void issue_group_3x3_nt(int nt, uint32_t tacc, uint32_t block_n, uint64_t adesc0, uint64_t bdesc0,
                                                   uint32_t bstep, uint32_t idesc, uint32_t accum_first) {
  if (nt == 9) issue_group_3x3<KS, 9, HALVES>(tacc, block_n, adesc0, bdesc0, bstep, idesc, accum_first);
  else if (nt == 3) issue_group_3x3<KS, 3, HALVES>(tacc, block_n, adesc0, bdesc0, bstep, idesc, accum_first);
  else issue_group_3x3<KS, 1, HALVES>(tacc, block_n, adesc0, bdesc0, bstep, idesc, accum_first);
}
template <int HALVES>
__device__ __forceinline__ void issue_group_3x3_ks(int ksteps, int nt, uint32_t tacc, uint32_t block_n, uint64_t ad, uint64_t bd,
                                                   uint32_t bstep, uint32_t idesc, uint32_t accum_first) {
  switch (ksteps) {
    case 4: issue_group_3x3_nt<4, HALVES>(nt, tacc, block_n, ad, bd, bstep, idesc, accum_first); break;
    case 3: issue_group_3x3_nt<3, HALVES>(nt, tacc, block_n, ad, bd, bstep, idesc, accum_first); break;
    case 2: issue_group_3x3_nt<2, HALVES>(nt, tacc, block_n, ad, bd, bstep, idesc, accum_first); break;
    default: issue_group_3x3_nt<1, HALVES>(nt, tacc, block_n, ad, bd, bstep, idesc, accum_first); break;
  }
}

__global__ void __launch_bounds__(CV_THREADS, 1)
    conv_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.block_n * 128;             // one tap's [block_n x 64] weight tile
  const int b_stage_bytes = p.b_group * b_bytes;
  uint8_t* smA = smem;
  uint8_t* smB = smem + (size_t)p.a_stages * p.a_stage_bytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smB + (size_t)p.b_stages * b_stage_bytes);
  uint64_t* emptyA = fullA + p.a_stages;
  uint64_t* fullB = emptyA + p.a_stages;
  uint64_t* emptyB = fullB + p.b_stages;
  uint64_t* tmem_full = emptyB + p.b_stages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  __shared__ __align__(16) float sbias[512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.block_n;
  stage_bias(sbias, p.bias, n0, p.N, p.block_n, p.scale);

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

  const int slabs = p.kd * p.kblocks;          // activation boxes per brick
  const int inplane = p.kh * p.kw;             // taps served by one box

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
        const int h0 = ht * CV_TH, w0 = wt * CV_TW * p.halves;
        int kd_ = 0, cb = 0;
        for (int s = 0; s < slabs; ++s) {
          mbar_wait(&emptyA[st], ph ^ 1);
          mbar_expect_tx(&fullA[st], (uint32_t)(p.box_rows * 128));
          tma_load_5d(smA + (size_t)st * p.a_stage_bytes, &tmA, &fullA[st], cb * 64, w0 - p.pw, h0 - p.ph,
                      d0 + kd_ - p.pd, bn);
          if (++st == p.a_stages) { st = 0; ph ^= 1; }
          if (++cb == p.kblocks) { cb = 0; ++kd_; }
        }
      }
    }
  } else if (warp == 2) {
    // ---------------- weight tiles: groups of b_group in-plane taps of one (depth tap, slab) per barrier ----------------
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      const int cin_pad = p.kblocks * 64;
      bool first = true;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        if (p.b_resident && !first) break;       // resident filter: loaded once
        first = false;
        int kd_ = 0, cb = 0;
        for (int s = 0; s < slabs; ++s) {
          for (int g = 0; g < p.b_groups; ++g) {
            const int tp0 = g * p.b_group;
            const int nt = min(p.b_group, inplane - tp0);
            mbar_wait(&emptyB[st], ph ^ 1);
            mbar_expect_tx(&fullB[st], (uint32_t)(nt * b_bytes));
            for (int j = 0; j < nt; ++j)
              tma_load_2d(smB + (size_t)st * b_stage_bytes + (size_t)j * b_bytes, &tmB, &fullB[st],
                          (kd_ * inplane + tp0 + j) * cin_pad + cb * 64, n0);
            if (++st == p.b_stages) { st = 0; ph ^= 1; }
          }
          if (++cb == p.kblocks) { cb = 0; ++kd_; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: the whole warp runs the (warp-uniform) control flow so that descriptors live in
    // uniform registers; one elected lane issues the tcgen05.mma / commit instructions ----------------
    {
      const uint32_t idesc = make_idesc_bf16(128, p.block_n);
      const uint32_t sbo = (uint32_t)p.box_w * 128;
      const int last_ksteps = (p.Cin - (p.kblocks - 1) * 64 + 15) / 16;   // K16 steps of the last (ragged) slab
      const bool fast33 = p.kh == 3 && p.kw == 3 && (p.b_group == 9 || p.b_group == 3 || p.b_group == 1);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
      long long w_acc = 0, w_a = 0, w_b = 0, c0 = 0;
      const long long t_all = clock64();
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        if (dbg) c0 = clock64();
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
        if (dbg) w_acc += clock64() - c0;
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * p.halves * p.block_n);
        uint32_t first_mma = 1;
        int cb = 0;
        for (int s = 0; s < slabs; ++s) {
          if (dbg) c0 = clock64();
          mbar_wait(&fullA[sa], pa);
          if (dbg) w_a += clock64() - c0;
          const uint32_t abase = smem_u32(smA + (size_t)sa * p.a_stage_bytes);
          const int ksteps = (cb == p.kblocks - 1) ? last_ksteps : 4;
          int kh_ = 0, kw_ = 0;
          for (int g = 0; g < p.b_groups; ++g) {
            const int nt = min(p.b_group, inplane - g * p.b_group);
            if (dbg) c0 = clock64();
            if (!(p.b_resident && it > 0)) mbar_wait(&fullB[sb], pb);   // resident weights: waited once, on brick 0
            if (dbg) w_b += clock64() - c0;
            tc_fence_after();
            const uint32_t bbase = smem_u32(smB + (size_t)sb * b_stage_bytes);
            if (fast33) {
              if (elect_one()) {
                const uint64_t ad = make_kmajor_sw128_desc_rows(abase + (uint32_t)(kh_ * p.box_w + kw_) * 128, sbo);
                const uint64_t bd = make_kmajor_sw128_desc(bbase);
                const uint32_t bstep = (uint32_t)b_bytes >> 4;
                if (p.halves == 2) issue_group_3x3_ks<2>(ksteps, nt, tacc, (uint32_t)p.block_n, ad, bd, bstep, idesc, first_mma ^ 1u);
                else issue_group_3x3_ks<1>(ksteps, nt, tacc, (uint32_t)p.block_n, ad, bd, bstep, idesc, first_mma ^ 1u);
                if (!p.b_resident) umma_commit(&emptyB[sb]);
              }
            } else if (elect_one()) {
              int kh2 = kh_, kw2 = kw_;
              for (int j = 0; j < nt; ++j) {
                const uint32_t astart = abase + (uint32_t)(kh2 * p.box_w + kw2) * 128;
                const uint64_t bdesc = make_kmajor_sw128_desc(bbase + (uint32_t)(j * b_bytes));
                for (int k = 0; k < ksteps; ++k) {
                  for (int h = 0; h < p.halves; ++h)
                    umma_f16(tacc + (uint32_t)(h * p.block_n), make_kmajor_sw128_desc_rows(astart + h * 1024 + k * 32, sbo),
                             bdesc + (uint64_t)(2 * k), idesc, first_mma ^ 1u);
                  first_mma = 0;
                }
                if (++kw2 == p.kw) { kw2 = 0; ++kh2; }
              }
              if (!p.b_resident) umma_commit(&emptyB[sb]);
            }
            __syncwarp();
            first_mma = 0;
            kw_ += nt;
            while (kw_ >= p.kw) { kw_ -= p.kw; ++kh_; }
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
      if (dbg && lane == 0) { p.dbg[0] = clock64() - t_all; p.dbg[1] = w_acc; p.dbg[2] = w_a; p.dbg[3] = w_b; p.dbg[4] = it; }
    }
  } else {
    // ---------------- epilogue: warps 3..6 -> TMEM lane quarters warp % 4 ----------------
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int hy = r >> 3, wx = r & 7;
    int it = 0;
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && warp == 3 && lane == 0;
    long long e_wait = 0, c0 = 0;
    const long long e_all = clock64();
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      long long t = tile;
      const int wt = (int)(t % p.nw); t /= p.nw;
      const int ht = (int)(t % p.nh); t /= p.nh;
      const int d0 = (int)(t % p.D);
      const int bn = (int)(t / p.D);
      const int acc = it & 1;
      if (dbg) c0 = clock64();
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      if (dbg) e_wait += clock64() - c0;
      tc_fence_after();
      for (int half = 0; half < p.halves; ++half) {
        const int h = ht * CV_TH + hy, w = (wt * p.halves + half) * CV_TW + wx;
        const long long out_row = (h < p.H && w < p.W) ? (((long long)bn * p.D + d0) * p.H + h) * p.W + w : -1;
        const uint32_t tcol = (uint32_t)((acc * p.halves + half) * p.block_n);
        if (p.out_dtype == NEXTOU_BF16) {
          __nv_bfloat16* dst = out_row >= 0 ? reinterpret_cast<__nv_bfloat16*>(p.C) + out_row * p.ldc + n0 : nullptr;
          const long long left = p.ldc - n0;
          epilogue_row_bf16(tmem_base + ((uint32_t)(q * 32) << 16) + tcol, p.block_n, sbias, dst,
                            (int)(left < p.block_n ? left : p.block_n), p.row32 != 0, p.slope);
        } else
        for (int c = 0; c < p.block_n; c += 16) {
          uint32_t raw[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + tcol + (uint32_t)c, raw);
          tmem_ld_wait();
          if (out_row >= 0 && n0 + c < p.ldc) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = n0 + c + j;
              const float x = affine_act(raw[j], sbias[256 + c + j], sbias[c + j], p.slope);
              v[j] = col < p.N ? x : 0.f;
            }
            store_chunk16(reinterpret_cast<float*>(p.C) + out_row * p.ldc, n0 + c, v, p.ldc);
          }
        }
      }
      // accumulator drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (dbg) { p.dbg[5] = clock64() - e_all; p.dbg[6] = e_wait; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace nextou

using namespace nextou;

static long long* g_conv_dbg = nullptr;
// diagnostics: device buffer of 16 int64 that CTA (0,0) fills with per-role cycle counters (NULL disables)
extern "C" void nextou_debug_set_conv_counters(long long* dev_buf) { g_conv_dbg = dev_buf; }

// Same contract as nextou_conv3d_ndhwc_fwd, restricted to kh, kw in {1, 3} (any odd kd).
extern "C" int nextou_conv3d_ndhwc_halo_fwd(const void* x, long long ldx, int B, int D, int H, int W, int Cin,
                                            const void* wpack, int Cout, int kd, int kh, int kw, const float* bias,
                                            void* out, long long ldo, int out_dtype, void* stream) {
  return nextou_conv3d_ndhwc_halo_fwd_affine(x, ldx, B, D, H, W, Cin, wpack, Cout, kd, kh, kw, nullptr, bias, 1.f, out, ldo,
                                             out_dtype, stream);
}

// out = lrelu(conv(x) * scale[Cout] + shift[Cout], slope): inference form with the eval-mode norm (+ LeakyReLU) folded in
extern "C" int nextou_conv3d_ndhwc_halo_fwd_affine(const void* x, long long ldx, int B, int D, int H, int W, int Cin,
                                                   const void* wpack, int Cout, int kd, int kh, int kw, const float* scale,
                                                   const float* bias, float slope, void* out, long long ldo, int out_dtype,
                                                   void* stream) {
  NEXTOU_REQUIRE(x && wpack && out, "conv3d_ndhwc_halo_fwd: null pointer");
  NEXTOU_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3d_ndhwc_halo_fwd: bad shape");
  NEXTOU_REQUIRE((kh == 1 || kh == 3) && (kw == 1 || kw == 3) && kd % 2 == 1 && kd <= 7,
                 "conv3d_ndhwc_halo_fwd: kh, kw must be 1 or 3 (got %d x %d x %d)", kd, kh, kw);
  NEXTOU_REQUIRE(ldx % 8 == 0 && ldx >= Cin && ldo % 8 == 0 && ldo >= Cout, "conv3d_ndhwc_halo_fwd: pitches must be multiples of 8");
  NEXTOU_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)wpack & 15) == 0 && ((uintptr_t)out & 15) == 0, "conv3d_ndhwc_halo_fwd: 16-byte alignment");
  NEXTOU_REQUIRE(out_dtype == NEXTOU_BF16 || out_dtype == NEXTOU_F32, "conv3d_ndhwc_halo_fwd: bad out dtype");
  ConvParams p = {};
  p.N = Cout; p.Cin = Cin;
  p.block_n = pick_block_n(Cout);
  p.kblocks = (Cin + 63) / 64;
  p.kd = kd; p.kh = kh; p.kw = kw; p.pd = kd / 2; p.ph = kh / 2; p.pw = kw / 2;
  p.D = D; p.H = H; p.W = W; p.B = B;
  p.C = out; p.ldc = ldo; p.out_dtype = out_dtype; p.bias = bias; p.scale = scale; p.slope = slope; p.dbg = g_conv_dbg;
  p.row32 = (ldo % 16 == 0 && ((uintptr_t)out & 31) == 0) ? 1 : 0;
  const int taps = kd * kh * kw, inplane = kh * kw;
  const int cin_pad = p.kblocks * 64;
  const int b_bytes = p.block_n * 128;
  const int total_budget = 200 * 1024;
  const long long all_b = (long long)taps * p.kblocks * b_bytes;
  // Filters that do not fit shared memory are streamed through a ring, once per brick: 553 KB per 128 voxels for the 66 -> 66
  // 3x3x3 layers, which makes the kernel bound by weight traffic (L2 -> shared memory writes contend with the operand reads
  // of the MMAs; measured 107 cycles per MMA instead of 84).  Then a CTA takes 16 x 16 bricks: two M = 128 accumulators that
  // consume every streamed weight tile twice (needs 4 x block_n tensor-memory columns for the double buffer).
  const bool resident_narrow = all_b <= total_budget - 3 * (((CV_TW + kw - 1) * (CV_TH + kh - 1) * 128 + 1023) / 1024 * 1024) &&
                               kd * p.kblocks <= CV_B_MAX_STAGES;
  p.halves = (!resident_narrow && 4 * p.block_n <= 512 && W > CV_TW && kw == 3 && kh == 3) ? 2 : 1;
  p.tmem_cols = pow2_cols(2 * p.halves * p.block_n);   // double-buffered accumulator(s)
  p.nh = (H + CV_TH - 1) / CV_TH; p.nw = (W + CV_TW * p.halves - 1) / (CV_TW * p.halves);
  p.box_w = CV_TW * p.halves + kw - 1;
  p.box_rows = p.box_w * (CV_TH + kh - 1);
  p.a_stage_bytes = (p.box_rows * 128 + 1023) / 1024 * 1024;
  p.a_stages = p.halves == 2 ? 2 : 3;
  const int b_budget = total_budget - p.a_stages * p.a_stage_bytes;
  if (resident_narrow) {
    // whole filter resident: one stage per (depth tap, slab) holding all in-plane taps
    p.b_resident = 1; p.b_group = inplane; p.b_groups = 1; p.b_stages = kd * p.kblocks;
    int as = (int)((total_budget - all_b) / p.a_stage_bytes);   // spend the rest on a deeper activation ring
    if (as > CV_A_MAX_STAGES) as = CV_A_MAX_STAGES;
    if (as > p.a_stages) p.a_stages = as;
  } else {
    p.b_resident = 0;
    p.b_group = (inplane == 9 && 3 * b_bytes * 2 <= b_budget) ? 3 : 1;
    p.b_groups = (inplane + p.b_group - 1) / p.b_group;
    int st = b_budget / (p.b_group * b_bytes);
    if (st > 6) st = 6;
    NEXTOU_REQUIRE(st >= 2, "conv3d_ndhwc_halo_fwd: weight tile does not fit (Cout tile %d)", p.block_n);
    p.b_stages = st;
  }
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * W, (cuuint64_t)ldx * 2 * W * H,
                         (cuuint64_t)ldx * 2 * W * H * D};
    cuuint32_t box[5] = {64, (cuuint32_t)p.box_w, (cuuint32_t)(CV_TH + kh - 1), 1, 1};
    int rc = encode_bf16_map(&tmA, x, 5, dims, str, box, "conv halo input");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)taps * cin_pad, (cuuint64_t)Cout};
    cuuint64_t str[1] = {(cuuint64_t)taps * cin_pad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)p.block_n};
    int rc = encode_bf16_map(&tmB, wpack, 2, dims, str, box, "conv halo weights");
    if (rc) return rc;
  }
  const size_t smem = 1024 + (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * p.b_group * b_bytes +
                      (2 * p.a_stages + 2 * p.b_stages + 4) * sizeof(uint64_t) + 16;
  int rc = ensure_smem(conv_halo_tcgen05_kernel, smem);
  if (rc) return rc;
  p.total_tiles = (long long)B * D * p.nh * p.nw;
  const int n_tiles = (Cout + p.block_n - 1) / p.block_n;
  NEXTOU_REQUIRE(n_tiles <= 65535, "conv3d_ndhwc_halo_fwd: grid too large");
  // persistent grid: as many CTAs as can be resident (smem / TMEM bound), split over the Cout tiles
  int per_sm = (int)((220 * 1024) / smem);
  if (per_sm > 512 / p.tmem_cols) per_sm = 512 / p.tmem_cols;
  if (per_sm < 1) per_sm = 1;
  // one resident wave: rounding UP here put 150 CTAs of a 3-tile GEMM on 148 SMs, i.e. a second wave for 2 CTAs (2x kernel time)
  long long ctas = ((long long)num_sms() * per_sm) / n_tiles;
  if (ctas > p.total_tiles) ctas = p.total_tiles;
  if (ctas < 1) ctas = 1;
  dim3 grid((unsigned)ctas, (unsigned)n_tiles);
  conv_halo_tcgen05_kernel<<<grid, CV_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, p);
  return check_launch("conv_halo_tcgen05_kernel");
}

// ======================================================================================================
// Weight gradient with halo reuse (kh, kw in {1, 3}):  dW[co][tap][ci] += sum_v dY[v][co] * X[v + tap - pad][ci]
// K block = an 8 x 8 voxel brick of one depth slice.  Per K block a CTA fetches the dY brick (M = 128 channels, two
// {64 ch, 8, 8} boxes) and ONE haloed X box {64 ch, 8+kw-1, 8+kh-1} per 64-channel slab of its Cin tile; every
// in-plane tap multiplies a row-shifted view of that box (MN-major operand: rows = voxels = K, 8-voxel groups one box
// row pitch apart) into its own fp32 accumulator slab in tensor memory.  Depth taps are separate CTAs (grid), the
// voxel axis is split over CTAs and reduced with fp32 atomics into the zero-filled dW.
// ======================================================================================================
namespace nextou {

constexpr int WH_THREADS = 192;
constexpr int WH_MAX_STAGES = 8;

struct WgradHaloParams {
  int Cout, Cin;
  int D, H, W, B;
  int nh, nw;                 // 8x8 bricks per slice
  int kd, kh, kw, pd, ph, pw;
  int n_tile, n_boxes;        // Cin tile per CTA (multiple of 16), 64-channel boxes it spans
  int box_w, xbox_bytes;      // haloed X box width (rows per H line) and its 1024-aligned size
  int n_mtiles, n_ntiles, ksplit, tmem_cols, stages;
  int grp_rows, n_rgroups;    // kernel rows (kh) per CTA and number of such groups: a CTA owns grp_rows * kw in-plane taps
  long long total_bricks;
  float* dW;
  int cin_stride;
};

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc2(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// MMAs of NROWS kernel rows x 3 kernel columns x 4 K16-steps with compile-time descriptor deltas
template <int NROWS>
__device__ __forceinline__ void wgrad_issue_rows(uint32_t tmem_base, int n_tile, uint64_t ad0, uint64_t bd0, uint32_t idesc,
                                                 uint32_t accum0) {
#pragma unroll
  for (int tp = 0; tp < 3 * NROWS; ++tp) {
    const uint32_t toff = (uint32_t)((tp / 3) * 10 + (tp % 3)) * 8u;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_f16(tmem_base + (uint32_t)(tp * n_tile), ad0 + (uint64_t)(k * 128), bd0 + (uint64_t)(toff + k * 160), idesc,
               k != 0 ? 1u : accum0);
  }
}

__global__ void __launch_bounds__(WH_THREADS, 1)
    wgrad_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                              const WgradHaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = 2 * 8192 + p.n_boxes * p.xbox_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid = (tiles, ksplit): the tiles (kernel depth / row group / channel tiles) of one brick range are launched back to back and
  // sweep the same dY / X bricks at the same time, so all but the first fetch hit L2 (with the tile index as the slow grid
  // axis every depth tap re-read both tensors from DRAM: 3x the algorithmic traffic on the 3x3x3 layers)
  int t = blockIdx.x;
  const int nt = t % p.n_ntiles; t /= p.n_ntiles;
  const int mt = t % p.n_mtiles; t /= p.n_mtiles;
  const int rg = t % p.n_rgroups; t /= p.n_rgroups;
  const int kd_ = t;
  const int m0 = mt * 128, n0 = nt * p.n_tile;
  const int ks = blockIdx.y;
  const int r0 = rg * p.grp_rows;                          // first kernel row of this CTA
  const int nrows = min(p.grp_rows, p.kh - r0);
  const int inplane = nrows * p.kw;                        // in-plane taps of this CTA: rows r0 .. r0 + nrows - 1

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const long long nb = p.total_bricks > ks ? (p.total_bricks - ks + p.ksplit - 1) / p.ksplit : 0;
  const int box_rows = p.box_w * (8 + p.grp_rows - 1);   // the X box spans only the kernel rows of this group

  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (long long i = 0; i < nb; ++i) {
        long long b = ks + i * p.ksplit;
        const int wt = (int)(b % p.nw); b /= p.nw;
        const int ht = (int)(b % p.nh); b /= p.nh;
        const int d0 = (int)(b % p.D);
        const int bn = (int)(b / p.D);
        const int w0 = wt * 8, h0 = ht * 8;
        mbar_wait(&empty_bar[st], ph ^ 1);
        mbar_expect_tx(&full_bar[st], (uint32_t)(2 * 8192 + p.n_boxes * box_rows * 128));
        uint8_t* sp = smem + (size_t)st * stage_bytes;
        tma_load_5d(sp, &tmDY, &full_bar[st], m0, w0, h0, d0, bn);
        tma_load_5d(sp + 8192, &tmDY, &full_bar[st], m0 + 64, w0, h0, d0, bn);
        for (int j = 0; j < p.n_boxes; ++j)
          tma_load_5d(sp + 2 * 8192 + (size_t)j * p.xbox_bytes, &tmX, &full_bar[st], n0 + 64 * j, w0 - p.pw, h0 - p.ph + r0,
                      d0 + kd_ - p.pd, bn);
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole warp runs the warp-uniform control flow (descriptors in uniform registers); one elected lane issues
    if (nb > 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.n_tile) | (1u << 15) | (1u << 16);  // both operands MN-major
      const uint32_t sbo_x = (uint32_t)p.box_w * 128;
      const bool fast33 = p.kh == 3 && p.kw == 3;
      int st = 0;
      uint32_t ph = 0;
      for (long long i = 0; i < nb; ++i) {
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        const uint32_t sp = smem_u32(smem + (size_t)st * stage_bytes);
        const uint32_t xb = sp + 2 * 8192;
        if (fast33) {
          if (elect_one()) {
            // compile-time descriptor deltas: tap (r, kw) starts (r*10 + kw) rows into the haloed box, a K16 step
            // advances dY by 16 rows (2048 B) and X by two 8-voxel groups (2 box rows of 10 x 128 B)
            const uint64_t ad0 = make_mnmajor_sw128_desc2(sp, 8192, 1024);
            const uint64_t bd0 = make_mnmajor_sw128_desc2(xb, (uint32_t)p.xbox_bytes, 1280);
            const uint32_t accum0 = i != 0 ? 1u : 0u;
            if (nrows == 3) wgrad_issue_rows<3>(tmem_base, p.n_tile, ad0, bd0, idesc, accum0);
            else if (nrows == 2) wgrad_issue_rows<2>(tmem_base, p.n_tile, ad0, bd0, idesc, accum0);
            else wgrad_issue_rows<1>(tmem_base, p.n_tile, ad0, bd0, idesc, accum0);
            umma_commit(&empty_bar[st]);
          }
        } else if (elect_one()) {
          int kh_ = 0, kw_ = 0;
          for (int tp = 0; tp < inplane; ++tp) {
            const uint32_t xstart = xb + (uint32_t)(kh_ * p.box_w + kw_) * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 16 voxels = two 8-voxel groups per MMA
              const uint64_t adesc = make_mnmajor_sw128_desc2(sp + k * 2048, 8192, 1024);
              const uint64_t bdesc = make_mnmajor_sw128_desc2(xstart + k * 2 * sbo_x, (uint32_t)p.xbox_bytes, sbo_x);
              umma_f16(tmem_base + (uint32_t)(tp * p.n_tile), adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
            }
            if (++kw_ == p.kw) { kw_ = 0; ++kh_; }
          }
          umma_commit(&empty_bar[st]);
        }
        __syncwarp();
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(tmem_full);
      __syncwarp();
    }
  } else if (nb > 0) {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int co = m0 + q * 32 + lane;
    const bool vec = (p.cin_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dW) & 15) == 0;
    for (int tp = 0; tp < inplane; ++tp) {
      const int tap = kd_ * (p.kh * p.kw) + r0 * p.kw + tp;
      for (int c = 0; c < p.n_tile; c += 16) {
        uint32_t raw[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tp * p.n_tile + c), raw);
        tmem_ld_wait();
        if (co < p.Cout)   // columns in [Cin, cin_stride) receive exact zeros (X channels beyond Cin are zero-filled by TMA)
          red_add_row16(p.dW + ((long long)co * (p.kd * p.kh * p.kw) + tap) * p.cin_stride + n0 + c, raw, p.cin_stride - n0 - c, vec);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace nextou

extern "C" int nextou_conv3d_ndhwc_halo_wgrad(const void* dy, long long ldy, const void* x, long long ldx, int B, int D,
                                              int H, int W, int Cin, int Cout, int kd, int kh, int kw, float* dW,
                                              int cin_stride, void* stream) {
  NEXTOU_REQUIRE(dy && x && dW, "conv3d_ndhwc_halo_wgrad: null pointer");
  NEXTOU_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && cin_stride >= Cin, "conv3d_ndhwc_halo_wgrad: bad shape");
  NEXTOU_REQUIRE((kh == 1 || kh == 3) && (kw == 1 || kw == 3) && kd % 2 == 1 && kd <= 7,
                 "conv3d_ndhwc_halo_wgrad: kh, kw must be 1 or 3 (got %d x %d x %d)", kd, kh, kw);
  NEXTOU_REQUIRE(ldy % 8 == 0 && ldy >= Cout && ldx % 8 == 0 && ldx >= Cin, "conv3d_ndhwc_halo_wgrad: pitches must be multiples of 8");
  NEXTOU_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0, "conv3d_ndhwc_halo_wgrad: 16-byte alignment");
  WgradHaloParams p = {};
  p.Cout = Cout; p.Cin = Cin; p.D = D; p.H = H; p.W = W; p.B = B;
  p.kd = kd; p.kh = kh; p.kw = kw; p.pd = kd / 2; p.ph = kh / 2; p.pw = kw / 2;
  p.dW = dW; p.cin_stride = cin_stride;
  p.nh = (H + 7) / 8; p.nw = (W + 7) / 8;
  p.total_bricks = (long long)B * D * p.nh * p.nw;
  // Every in-plane tap of a CTA owns n_tile fp32 columns of tensor memory (512 in all).  One MMA costs ~62 cycles up to
  // N = 64 and grows slowly beyond (shared-memory operand reads), so the cheapest plan has the FEWEST Cin tiles: give a CTA
  // fewer kernel rows (more tap groups, each re-reading dY) when that lets one wide tile cover Cin.
  p.grp_rows = kh;
  {
    double best = 1e30;
    for (int rows = kh; rows >= 1; --rows) {
      if (rows != kh && !(kh == 3 && kw == 3)) break;           // row groups only on the unrolled 3x3 path
      int max_n = (512 / (rows * kw)) / 16 * 16;
      if (max_n > 256) max_n = 256;
      if (max_n < 16) continue;
      const int nt = (Cin + max_n - 1) / max_n;
      const int ntile = ((Cin + nt - 1) / nt + 15) / 16 * 16;
      const int groups = (kh + rows - 1) / rows;
      const double mma = ntile <= 64 ? 62.0 : 44.0 + 0.2 * ntile + (ntile > 160 ? 0.3 * (ntile - 160) : 0.0);
      const double cost = (double)nt * (kh * kw * 4 * mma + groups * 350.0);   // + per-iteration barrier / dY re-read cost
      if (cost < best * 0.97) { best = cost; p.grp_rows = rows; }
    }
  }
  p.n_rgroups = (kh + p.grp_rows - 1) / p.grp_rows;
  const int inplane = p.grp_rows * kw;       // taps per CTA
  int max_n = (512 / inplane) / 16 * 16;
  if (max_n > 256) max_n = 256;
  p.n_ntiles = (Cin + max_n - 1) / max_n;
  p.n_tile = ((Cin + p.n_ntiles - 1) / p.n_ntiles + 15) / 16 * 16;
  p.n_boxes = (p.n_tile + 63) / 64;
  p.n_mtiles = (Cout + 127) / 128;
  p.box_w = 8 + kw - 1;
  p.xbox_bytes = (p.box_w * (8 + p.grp_rows - 1) * 128 + 1023) / 1024 * 1024;
  p.tmem_cols = pow2_cols(inplane * p.n_tile);
  const long long tiles = (long long)kd * p.n_rgroups * p.n_mtiles * p.n_ntiles;
  long long ksplit = (3LL * num_sms() + tiles - 1) / tiles;
  if (ksplit > p.total_bricks / 16) ksplit = p.total_bricks / 16;   // >= 16 K blocks per CTA amortise prologue + atomics
  if (ksplit < 1) ksplit = 1;
  p.ksplit = (int)ksplit;
  CUtensorMap tmDY, tmX;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)ldy * 2, (cuuint64_t)ldy * 2 * W, (cuuint64_t)ldy * 2 * W * H,
                         (cuuint64_t)ldy * 2 * W * H * D};
    cuuint32_t box[5] = {64, 8, 8, 1, 1};
    int rc = encode_bf16_map(&tmDY, dy, 5, dims, str, box, "wgrad halo dY");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * W, (cuuint64_t)ldx * 2 * W * H,
                         (cuuint64_t)ldx * 2 * W * H * D};
    cuuint32_t box[5] = {64, (cuuint32_t)p.box_w, (cuuint32_t)(8 + p.grp_rows - 1), 1, 1};
    int rc = encode_bf16_map(&tmX, x, 5, dims, str, box, "wgrad halo X");
    if (rc) return rc;
  }
  {
    const int per_stage = 2 * 8192 + p.n_boxes * p.xbox_bytes;
    int st = (200 * 1024) / per_stage;
    if (st > WH_MAX_STAGES) st = WH_MAX_STAGES;
    if (st < 2) st = 2;
    p.stages = st;
  }
  const size_t smem = 1024 + (size_t)p.stages * (2 * 8192 + p.n_boxes * p.xbox_bytes) + (2 * p.stages + 1) * sizeof(uint64_t) + 16;
  int rc = ensure_smem(wgrad_halo_tcgen05_kernel, smem);
  if (rc) return rc;
  NEXTOU_REQUIRE(tiles <= 65535, "conv3d_ndhwc_halo_wgrad: too many tiles");
  dim3 grid((unsigned)tiles, (unsigned)p.ksplit);
  wgrad_halo_tcgen05_kernel<<<grid, WH_THREADS, smem, (cudaStream_t)stream>>>(tmDY, tmX, p);
  return check_launch("wgrad_halo_tcgen05_kernel");
}

// ======================================================================================================
// Weight gradient of the DOWN-SAMPLING convolution of an encoder stage (3x3 in-plane kernel, in-plane stride 2, padding 1;
// NexToU_Encoder_Decoder.py:125-141, stages 1-5) with halo reuse:  dW[co][tap][ci] += sum_o dY[o][co] * X[2o + tap - 1][ci].
// The per-tap kernel of gemm_tcgen05.cu gathers X with an element-strided TMA box per tap (every voxel a separate fetch, 9
// boxes per brick and depth tap: 895 us for enc s1 conv0, 672 MB of DRAM traffic against 273 MB algorithmic).  Here the input
// is addressed as its four (h, w)-PARITY PLANES — plain tensor maps over the same memory with doubled row / column strides,
// no copy — because tap (kh, kw) of output voxel (ho, wo) reads plane (kh != 1, kw != 1) at (ho - (kh == 0), wo - (kw == 0)):
// inside a plane the taps of an 8 x 8 brick of output voxels are row-shifted views of ONE haloed 9 x 9 box, exactly like the
// stride-1 kernel above.  Per brick a CTA fetches the dY brick and the four plane boxes and issues all nine in-plane taps.
// Requires ceil16(Cin) * 9 <= 512 tensor-memory columns (Cin <= 48: the full-resolution layer that dominates); depth taps and
// Cout tiles are separate CTAs, the voxel axis is split over CTAs and reduced with fp32 vector reds.
// ======================================================================================================
namespace nextou {

struct WgradPlanesParams {
  int Cout, Cin;
  int Do, Ho, Wo, B;          // output (dY) grid
  int nh, nw;                 // 8 x 8 bricks per output slice
  int kd, pd, sd;             // depth taps, padding and stride (in-plane: 3 x 3, padding 1, stride 2)
  int n_tile, n_mtiles, ksplit, tmem_cols, stages;
  int xbox_bytes;             // one plane box {64 ch, 9, 9} rounded up to 1024
  long long total_bricks;
  float* dW;
  int cin_stride;
};

__global__ void __launch_bounds__(WH_THREADS, 1)
    wgrad_planes_tcgen05_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX00,
                                const __grid_constant__ CUtensorMap tmX01, const __grid_constant__ CUtensorMap tmX10,
                                const __grid_constant__ CUtensorMap tmX11, const WgradPlanesParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = 2 * 8192 + 4 * p.xbox_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x % p.n_mtiles, kd_ = blockIdx.x / p.n_mtiles;   // tiles fastest: see wgrad_halo_tcgen05_kernel
  const int m0 = mt * 128;
  const int ks = blockIdx.y;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX00); prefetch_tmap(&tmX01); prefetch_tmap(&tmX10); prefetch_tmap(&tmX11);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const long long nb = p.total_bricks > ks ? (p.total_bricks - ks + p.ksplit - 1) / p.ksplit : 0;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (long long i = 0; i < nb; ++i) {
        long long b = ks + i * p.ksplit;
        const int wt = (int)(b % p.nw); b /= p.nw;
        const int ht = (int)(b % p.nh); b /= p.nh;
        const int d0 = (int)(b % p.Do);
        const int bn = (int)(b / p.Do);
        const int w0 = wt * 8, h0 = ht * 8;
        const int di = d0 * p.sd + kd_ - p.pd;                 // out-of-range depth: the TMA unit zero-fills (= padding)
        mbar_wait(&empty_bar[st], ph ^ 1);
        mbar_expect_tx(&full_bar[st], (uint32_t)(2 * 8192 + 4 * 81 * 128));
        uint8_t* sp = smem + (size_t)st * stage_bytes;
        tma_load_5d(sp, &tmDY, &full_bar[st], m0, w0, h0, d0, bn);
        tma_load_5d(sp + 8192, &tmDY, &full_bar[st], m0 + 64, w0, h0, d0, bn);
        uint8_t* xb = sp + 2 * 8192;
        // plane (ph, pw) holds input voxels (2i + ph, 2j + pw); every box starts one plane row / column before the brick
        tma_load_5d(xb, &tmX00, &full_bar[st], 0, w0 - 1, h0 - 1, di, bn);
        tma_load_5d(xb + p.xbox_bytes, &tmX01, &full_bar[st], 0, w0 - 1, h0 - 1, di, bn);
        tma_load_5d(xb + 2 * p.xbox_bytes, &tmX10, &full_bar[st], 0, w0 - 1, h0 - 1, di, bn);
        tma_load_5d(xb + 3 * p.xbox_bytes, &tmX11, &full_bar[st], 0, w0 - 1, h0 - 1, di, bn);
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (nb > 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.n_tile) | (1u << 15) | (1u << 16);  // both operands MN-major
      const uint32_t xbox16 = (uint32_t)p.xbox_bytes >> 4;
      int st = 0;
      uint32_t ph = 0;
      for (long long i = 0; i < nb; ++i) {
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        const uint32_t sp = smem_u32(smem + (size_t)st * stage_bytes);
        if (elect_one()) {
          const uint64_t ad0 = make_mnmajor_sw128_desc2(sp, 8192, 1024);
          const uint64_t bd0 = make_mnmajor_sw128_desc2(sp + 2 * 8192, (uint32_t)p.xbox_bytes, 9 * 128);
          const uint32_t accum0 = i != 0 ? 1u : 0u;
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            const int kh = tp / 3, kw = tp % 3;
            // tap (kh, kw): plane (kh != 1, kw != 1); inside the 9 x 9 box (origin = brick origin - 1) the view starts at row
            // (kh != 0), column (kw != 0); rows of 128 B, in 16-byte units
            const uint32_t toff = (uint32_t)((kh != 1 ? 2 : 0) + (kw != 1 ? 1 : 0)) * xbox16 +
                                  (uint32_t)((kh != 0 ? 9 : 0) + (kw != 0 ? 1 : 0)) * 8u;
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 16 voxels = two 8-voxel groups = 2 box rows of 9 x 128 B (144 units); dY: 2048 B
              umma_f16(tmem_base + (uint32_t)(tp * p.n_tile), ad0 + (uint64_t)(k * 128), bd0 + (uint64_t)(toff + k * 144), idesc,
                       k != 0 ? 1u : accum0);
          }
          umma_commit(&empty_bar[st]);
        }
        __syncwarp();
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(tmem_full);
      __syncwarp();
    }
  } else if (nb > 0) {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int co = m0 + q * 32 + lane;
    const bool vec = (p.cin_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dW) & 15) == 0;
    for (int tp = 0; tp < 9; ++tp) {
      const int tap = kd_ * 9 + tp;
      for (int c = 0; c < p.n_tile; c += 16) {
        uint32_t raw[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tp * p.n_tile + c), raw);
        tmem_ld_wait();
        if (co < p.Cout)
          red_add_row16(p.dW + ((long long)co * (p.kd * 9) + tap) * p.cin_stride + c, raw, p.cin_stride - c, vec);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace nextou

// 1 if nextou_conv3d_ndhwc_planes_wgrad covers the layer (else: nextou_conv3d_ndhwc_strided_wgrad)
extern "C" int nextou_conv3d_ndhwc_planes_wgrad_supported(int Cin, int kd, int kh, int kw, int sd, int sh, int sw, int pd, int ph, int pw) {
  const int n_tile = (Cin + 15) / 16 * 16;
  return kh == 3 && kw == 3 && sh == 2 && sw == 2 && ph == 1 && pw == 1 && (kd == 1 || kd == 3) && pd == kd / 2 && (sd == 1 || sd == 2) &&
         n_tile * 9 <= 512;
}

// dW[Cout][kd*9][cin_stride] (fp32, zero-filled by the caller) += sum_o dy[o][co] * x[(do*sd + a - pd, 2ho + kh - 1, 2wo + kw - 1)][ci]
// dy: bf16 tokens of the output grid [B][Do][Ho][Wo][ldy]; x: bf16 tokens of the input volume [B][Di][Hi][Wi][ldx]
extern "C" int nextou_conv3d_ndhwc_planes_wgrad(const void* dy, long long ldy, const void* x, long long ldx, int B, int Do, int Ho,
                                                int Wo, int Di, int Hi, int Wi, int Cin, int Cout, int kd, int sd, int pd, float* dW,
                                                int cin_stride, void* stream) {
  NEXTOU_REQUIRE(dy && x && dW, "conv3d_ndhwc_planes_wgrad: null pointer");
  NEXTOU_REQUIRE(nextou_conv3d_ndhwc_planes_wgrad_supported(Cin, kd, 3, 3, sd, 2, 2, pd, 1, 1), "conv3d_ndhwc_planes_wgrad: unsupported layer");
  NEXTOU_REQUIRE(B > 0 && Do > 0 && Ho > 0 && Wo > 0 && Di > 0 && Hi > 0 && Wi > 0 && Cout > 0 && cin_stride >= Cin, "conv3d_ndhwc_planes_wgrad: bad shape");
  NEXTOU_REQUIRE(Ho == (Hi + 2 - 3) / 2 + 1 && Wo == (Wi + 2 - 3) / 2 + 1 && Do == (Di + 2 * pd - kd) / sd + 1, "conv3d_ndhwc_planes_wgrad: grids do not match");
  NEXTOU_REQUIRE(ldy % 8 == 0 && ldy >= Cout && ldx % 8 == 0 && ldx >= Cin, "conv3d_ndhwc_planes_wgrad: pitches must be multiples of 8");
  NEXTOU_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0, "conv3d_ndhwc_planes_wgrad: 16-byte alignment");
  WgradPlanesParams p = {};
  p.Cout = Cout; p.Cin = Cin; p.Do = Do; p.Ho = Ho; p.Wo = Wo; p.B = B; p.kd = kd; p.pd = pd; p.sd = sd;
  p.dW = dW; p.cin_stride = cin_stride;
  p.nh = (Ho + 7) / 8; p.nw = (Wo + 7) / 8;
  p.total_bricks = (long long)B * Do * p.nh * p.nw;
  p.n_tile = (Cin + 15) / 16 * 16;
  p.n_mtiles = (Cout + 127) / 128;
  p.xbox_bytes = (81 * 128 + 1023) / 1024 * 1024;
  p.tmem_cols = pow2_cols(9 * p.n_tile);
  const long long tiles = (long long)kd * p.n_mtiles;
  long long ksplit = (3LL * num_sms() + tiles - 1) / tiles;
  if (ksplit > p.total_bricks / 16) ksplit = p.total_bricks / 16;
  if (ksplit < 1) ksplit = 1;
  p.ksplit = (int)ksplit;
  p.stages = 3;
  CUtensorMap tmDY, tmX[4];
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)Do, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)ldy * 2, (cuuint64_t)ldy * 2 * Wo, (cuuint64_t)ldy * 2 * Wo * Ho, (cuuint64_t)ldy * 2 * Wo * Ho * Do};
    cuuint32_t box[5] = {64, 8, 8, 1, 1};
    int rc = encode_bf16_map(&tmDY, dy, 5, dims, str, box, "wgrad planes dY");
    if (rc) return rc;
  }
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      // parity plane (ph, pw) of the input: the same memory with doubled H / W strides, starting at voxel (ph, pw)
      const int Hp = (Hi - ph + 1) / 2, Wp = (Wi - pw + 1) / 2;
      if (Hp <= 0 || Wp <= 0) {   // degenerate (1-voxel extents): an empty plane contributes nothing; alias plane (0, 0) out of range
        set_error("conv3d_ndhwc_planes_wgrad: extents < 2 are not supported");
        return NEXTOU_ERR_UNSUPPORTED;
      }
      cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)Di, (cuuint64_t)B};
      cuuint64_t str[4] = {(cuuint64_t)ldx * 2 * 2, (cuuint64_t)ldx * 2 * Wi * 2, (cuuint64_t)ldx * 2 * Wi * Hi,
                           (cuuint64_t)ldx * 2 * Wi * Hi * Di};
      cuuint32_t box[5] = {64, 9, 9, 1, 1};
      const char* base = reinterpret_cast<const char*>(x) + ((long long)ph * Wi + pw) * ldx * 2;
      int rc = encode_bf16_map(&tmX[ph * 2 + pw], base, 5, dims, str, box, "wgrad planes X");
      if (rc) return rc;
    }
  const size_t smem = 1024 + (size_t)p.stages * (2 * 8192 + 4 * p.xbox_bytes) + (2 * p.stages + 1) * sizeof(uint64_t) + 16;
  int rc = ensure_smem(wgrad_planes_tcgen05_kernel, smem);
  if (rc) return rc;
  NEXTOU_REQUIRE(tiles <= 65535, "conv3d_ndhwc_planes_wgrad: too many tiles");
  dim3 grid((unsigned)tiles, (unsigned)p.ksplit);
  wgrad_planes_tcgen05_kernel<<<grid, WH_THREADS, smem, (cudaStream_t)stream>>>(tmDY, tmX[0], tmX[1], tmX[2], tmX[3], p);
  return check_launch("wgrad_planes_tcgen05_kernel");
}

// ======================================================================================================
// Diagnostics: issue / completion cost of back-to-back tcgen05.mma (M = 128, K = 16, bf16, SS operands) for a given N.
// out[0] = cycles to ISSUE `reps` MMAs, out[1] = cycles until the last one has COMPLETED (commit + wait).
// ======================================================================================================
namespace nextou {
// flags: 1 = warps 4..7 stream tcgen05.ld from the other half of tensor memory while the MMAs are issued;
//        2 = warps 4..7 spin on mbarrier try_wait instead;  4 = per-MMA varying A start row (like the halo conv).
__global__ void __launch_bounds__(256, 1) mma_rate_probe_kernel(int n, int reps, int flags, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2, never;
  __shared__ uint32_t holder;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < (32768 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&never, 1); stop = 0; fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(&holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = holder;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    const long long t0 = clock64();
    for (int i = 0; i < reps / 4; ++i) {
      tc_fence_after();
      const uint32_t astart = a + ((flags & 4) ? (uint32_t)((i % 9) / 3 * 10 + (i % 3)) * 128 : 0u);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm, make_kmajor_sw128_desc_rows(astart + k * 32, (flags & 4) ? 1280 : 1024),
                   make_kmajor_sw128_desc(b) + (uint64_t)(2 * k), idesc, 1u);
        umma_commit(&bar2);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    stop = 1;
  } else if (warp >= 4) {
    const int q = warp & 3;
    if (flags & 1) {
      uint32_t raw[16];
      uint32_t acc = 0;
      while (!stop) {
        tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + 256u, raw);
        tmem_ld_wait();
        acc += raw[0];
      }
      if (acc == 0x12345678u) out[3] = acc;
    } else if (flags & 2) {
      const uint32_t addr = smem_u32(&never);
      uint32_t done = 0;
      while (!stop && !done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(0u) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}
}  // namespace nextou

extern "C" int nextou_debug_mma_rate(int n, int reps, int flags, long long* out_dev, void* stream) {
  NEXTOU_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && reps > 0 && out_dev, "debug_mma_rate: bad args");
  const size_t smem = 1024 + 32768 + 32768;
  int rc = ensure_smem(mma_rate_probe_kernel, smem);
  if (rc) return rc;
  mma_rate_probe_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(n, reps, flags, out_dev);
  return check_launch("mma_rate_probe_kernel");
}

// Diagnostics: sustained cost of TMA box loads issued by one thread into a 4-deep smem ring (no consumer work).
namespace nextou {
__global__ void __launch_bounds__(32, 1) tma_rate_probe_kernel(const __grid_constant__ CUtensorMap tm, int rank, int box_bytes,
                                                               int reps, int cstep, int wmax, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[4];
  const int slot = (box_bytes + 1023) / 1024 * 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const int s = i & 3;
      if (i >= 4) mbar_wait(&bars[s], ((i >> 2) - 1) & 1);
      mbar_expect_tx(&bars[s], (uint32_t)box_bytes);
      const int w0 = (i * cstep) % wmax;
      if (rank == 5) tma_load_5d(smem + (size_t)s * slot, &tm, &bars[s], 0, w0 - 1, ((i * 7) % 13) * 16 - 1, (i % 60), 0);
      else tma_load_2d(smem + (size_t)s * slot, &tm, &bars[s], 0, w0);
    }
    for (int i = reps; i < reps + 4; ++i) mbar_wait(&bars[i & 3], ((i >> 2) - 1) & 1);
    out[blockIdx.x] = clock64() - t0;
  }
}
}  // namespace nextou

// mode 0: conv halo box {64, 10, 18, 1, 1} over a [64][224][192][pitch] volume with `cin` valid channels;
// mode 1: 2-D box {64, 128} over a [rows][pitch] matrix.
extern "C" int nextou_debug_tma_rate(const void* base, int mode, int cin, int pitch, int reps, int ctas, long long* out_dev,
                                     void* stream) {
  CUtensorMap tm;
  int box_bytes, rank;
  if (mode == 0) {
    cuuint64_t dims[5] = {(cuuint64_t)cin, 192, 224, 64, 1};
    cuuint64_t str[4] = {(cuuint64_t)pitch * 2, (cuuint64_t)pitch * 2 * 192, (cuuint64_t)pitch * 2 * 192 * 224,
                         (cuuint64_t)pitch * 2 * 192 * 224 * 64};
    cuuint32_t box[5] = {64, 10, 18, 1, 1};
    int rc = encode_bf16_map(&tm, base, 5, dims, str, box, "probe5d");
    if (rc) return rc;
    box_bytes = 180 * 128; rank = 5;
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)64 * 224 * 192};
    cuuint64_t str[1] = {(cuuint64_t)pitch * 2};
    cuuint32_t box[2] = {64, 128};
    int rc = encode_bf16_map(&tm, base, 2, dims, str, box, "probe2d");
    if (rc) return rc;
    box_bytes = 128 * 128; rank = 2;
  }
  const size_t smem = 1024 + 4 * 24 * 1024;
  int rc = ensure_smem(tma_rate_probe_kernel, smem);
  if (rc) return rc;
  tma_rate_probe_kernel<<<ctas, 32, smem, (cudaStream_t)stream>>>(tm, rank, box_bytes, reps, mode == 0 ? 8 : 128,
                                                                 mode == 0 ? 184 : 64 * 224 * 192 - 128, out_dev);
  return check_launch("tma_rate_probe_kernel");
}
