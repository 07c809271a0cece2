// Stride-1 'same' convolution with kh, kw in {1, 3} (every StackedConvBlocks convolution of NexToU that does not
// down-sample: 1x3x3 and 3x3x3, reference ED:125-141, 281-300) as a HALO-REUSE implicit GEMM for sm_100a.
//
// csrc/gemm_tcgen05.cu fetches the activation brick once per tap (27 TMA boxes of 16 KB per 64-channel slab), which
// makes the 33/66-channel layers L2-bandwidth bound.  Here a CTA owns an 16 (H) x 8 (W) output brick of one depth slice
// and fetches, per depth tap and 64-channel slab, ONE haloed box {64 ch, 8+2, 16+2, 1} (23 KB).  All 9 in-plane taps
// read that box in place: the 8 voxels of an output row are 8 consecutive 128-byte rows of the box, so tap (kh, kw) is
// the same SWIZZLE_128B K-major operand with its start address moved by (kh*10 + kw) rows and its 8-row group stride
// (SBO) set to the box row pitch of 10 rows = 1280 B (the swizzle is a function of the absolute smem address, so the
// shifted start stays consistent with what TMA wrote).  Activation traffic from L2 drops 9 x 128/180 = 6.4x; the weights stream through their own ring.
//
// 224 threads: warp 0 activation-box producer, warp 1 MMA issuer, warp 2 weight producer (+ TMEM allocation),
// warps 3-6 epilogue (TMEM -> registers -> + bias -> global rows).
#include "tc_common.cuh"

namespace nextou {

constexpr int CV_TH = 16, CV_TW = 8;      // output brick (rows x cols) = 128 voxels = UMMA M
constexpr int CV_THREADS = 224;
constexpr int CV_A_STAGES = 2;
constexpr int CV_B_STAGES = 4;

struct ConvParams {
  int N;                 // Cout
  int block_n, tmem_cols;
  int kblocks;           // 64-channel slabs of Cin
  int kd, kh, kw, pd, ph, pw;
  int D, H, W, B;
  int nh, nw;            // bricks per slice
  int box_w, box_rows;   // haloed box: (8 + kw - 1) wide, box_w * (16 + kh - 1) rows
  int a_stage_bytes;     // box_rows * 128 rounded up to 1024
  void* C;
  long long ldc;
  int out_dtype;
  const float* bias;
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

__global__ void __launch_bounds__(CV_THREADS, 1)
    conv_halo_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.block_n * 128;
  uint8_t* smA = smem;
  uint8_t* smB = smem + (size_t)CV_A_STAGES * p.a_stage_bytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smB + (size_t)CV_B_STAGES * b_bytes);
  uint64_t* emptyA = fullA + CV_A_STAGES;
  uint64_t* fullB = emptyA + CV_A_STAGES;
  uint64_t* emptyB = fullB + CV_B_STAGES;
  uint64_t* tmem_full = emptyB + CV_B_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.block_n;
  int t = blockIdx.x;
  const int wt = t % p.nw; t /= p.nw;
  const int ht = t % p.nh; t /= p.nh;
  const int d0 = t % p.D;
  const int bn = t / p.D;
  const int h0 = ht * CV_TH, w0 = wt * CV_TW;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < CV_A_STAGES; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < CV_B_STAGES; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int slabs = p.kd * p.kblocks;          // activation boxes per tile
  const int inplane = p.kh * p.kw;             // taps served by one box

  if (warp == 0) {
    // ---------------- activation boxes ----------------
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int s = 0; s < slabs; ++s) {
        const int kd_ = s / p.kblocks, cb = s - kd_ * p.kblocks;
        mbar_wait(&emptyA[st], ph ^ 1);
        mbar_expect_tx(&fullA[st], (uint32_t)(p.box_rows * 128));
        tma_load_5d(smA + (size_t)st * p.a_stage_bytes, &tmA, &fullA[st], cb * 64, w0 - p.pw, h0 - p.ph, d0 + kd_ - p.pd, bn);
        if (++st == CV_A_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ---------------- weight tiles: one [block_n x 64] K-major tile per (depth tap, slab, in-plane tap) ----------------
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      const int cin_pad = p.kblocks * 64;
      for (int s = 0; s < slabs; ++s) {
        const int kd_ = s / p.kblocks, cb = s - kd_ * p.kblocks;
        for (int tp = 0; tp < inplane; ++tp) {
          mbar_wait(&emptyB[st], ph ^ 1);
          mbar_expect_tx(&fullB[st], (uint32_t)b_bytes);
          tma_load_2d(smB + (size_t)st * b_bytes, &tmB, &fullB[st], (kd_ * inplane + tp) * cin_pad + cb * 64, n0);
          if (++st == CV_B_STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.block_n);
      const uint32_t sbo = (uint32_t)p.box_w * 128;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int s = 0; s < slabs; ++s) {
        mbar_wait(&fullA[sa], pa);
        const uint32_t abase = smem_u32(smA + (size_t)sa * p.a_stage_bytes);
        for (int tp = 0; tp < inplane; ++tp) {
          const int kh_ = tp / p.kw, kw_ = tp - kh_ * p.kw;
          mbar_wait(&fullB[sb], pb);
          tc_fence_after();
          const uint32_t astart = abase + (uint32_t)(kh_ * p.box_w + kw_) * 128;
          const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smB + (size_t)sb * b_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = make_kmajor_sw128_desc_rows(astart + k * 32, sbo);
            umma_f16(tmem_base, adesc, bdesc + (uint64_t)(2 * k), idesc, (s | tp | k) != 0 ? 1u : 0u);
          }
          umma_commit(&emptyB[sb]);
          if (++sb == CV_B_STAGES) { sb = 0; pb ^= 1; }
        }
        umma_commit(&emptyA[sa]);
        if (++sa == CV_A_STAGES) { sa = 0; pa ^= 1; }
      }
      umma_commit(tmem_full);
    }
  } else {
    // ---------------- epilogue: warps 3..6 -> TMEM lane quarters warp % 4 ----------------
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int hy = r >> 3, wx = r & 7;
    const int h = h0 + hy, w = w0 + wx;
    const long long out_row = (h < p.H && w < p.W) ? (((long long)bn * p.D + d0) * p.H + h) * p.W + w : -1;
    for (int c = 0; c < p.block_n; c += 16) {
      uint32_t raw[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, raw);
      tmem_ld_wait();
      if (out_row >= 0 && n0 + c < p.ldc) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = n0 + c + j;
          float x = __uint_as_float(raw[j]);
          if (p.bias != nullptr && col < p.N) x += p.bias[col];
          v[j] = col < p.N ? x : 0.f;
        }
        if (p.out_dtype == NEXTOU_BF16)
          store_chunk16(reinterpret_cast<__nv_bfloat16*>(p.C) + out_row * p.ldc, n0 + c, v, p.ldc);
        else
          store_chunk16(reinterpret_cast<float*>(p.C) + out_row * p.ldc, n0 + c, v, p.ldc);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace nextou

using namespace nextou;

// Same contract as nextou_conv3d_ndhwc_fwd, restricted to kh, kw in {1, 3} (any odd kd).
extern "C" int nextou_conv3d_ndhwc_halo_fwd(const void* x, long long ldx, int B, int D, int H, int W, int Cin,
                                            const void* wpack, int Cout, int kd, int kh, int kw, const float* bias,
                                            void* out, long long ldo, int out_dtype, void* stream) {
  NEXTOU_REQUIRE(x && wpack && out, "conv3d_ndhwc_halo_fwd: null pointer");
  NEXTOU_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3d_ndhwc_halo_fwd: bad shape");
  NEXTOU_REQUIRE((kh == 1 || kh == 3) && (kw == 1 || kw == 3) && kd % 2 == 1 && kd <= 7,
                 "conv3d_ndhwc_halo_fwd: kh, kw must be 1 or 3 (got %d x %d x %d)", kd, kh, kw);
  NEXTOU_REQUIRE(ldx % 8 == 0 && ldx >= Cin && ldo % 8 == 0 && ldo >= Cout, "conv3d_ndhwc_halo_fwd: pitches must be multiples of 8");
  NEXTOU_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)wpack & 15) == 0 && ((uintptr_t)out & 15) == 0, "conv3d_ndhwc_halo_fwd: 16-byte alignment");
  NEXTOU_REQUIRE(out_dtype == NEXTOU_BF16 || out_dtype == NEXTOU_F32, "conv3d_ndhwc_halo_fwd: bad out dtype");
  ConvParams p = {};
  p.N = Cout;
  p.block_n = pick_block_n(Cout);
  p.tmem_cols = pow2_cols(p.block_n);
  p.kblocks = (Cin + 63) / 64;
  p.kd = kd; p.kh = kh; p.kw = kw; p.pd = kd / 2; p.ph = kh / 2; p.pw = kw / 2;
  p.D = D; p.H = H; p.W = W; p.B = B;
  p.nh = (H + CV_TH - 1) / CV_TH; p.nw = (W + CV_TW - 1) / CV_TW;
  p.box_w = CV_TW + kw - 1;
  p.box_rows = p.box_w * (CV_TH + kh - 1);
  p.a_stage_bytes = (p.box_rows * 128 + 1023) / 1024 * 1024;
  p.C = out; p.ldc = ldo; p.out_dtype = out_dtype; p.bias = bias;
  const int taps = kd * kh * kw;
  const int cin_pad = p.kblocks * 64;
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
  const size_t smem = 1024 + (size_t)CV_A_STAGES * p.a_stage_bytes + (size_t)CV_B_STAGES * p.block_n * 128 +
                      (2 * CV_A_STAGES + 2 * CV_B_STAGES + 1) * sizeof(uint64_t) + 16;
  int rc = ensure_smem(conv_halo_tcgen05_kernel, smem);
  if (rc) return rc;
  const long long tiles = (long long)B * D * p.nh * p.nw;
  const int n_tiles = (Cout + p.block_n - 1) / p.block_n;
  NEXTOU_REQUIRE(tiles <= 2147483647LL && n_tiles <= 65535, "conv3d_ndhwc_halo_fwd: grid too large");
  dim3 grid((unsigned)tiles, (unsigned)n_tiles);
  conv_halo_tcgen05_kernel<<<grid, CV_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, p);
  return check_launch("conv_halo_tcgen05_kernel");
}
