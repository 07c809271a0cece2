// First convolution of the network: Cin = the number of image modalities (1 for CT, <= 4 for MR), Cout = 33
// (StackedConvBlocks of encoder stage 0, reference NexToU_Encoder_Decoder.py:125-141; kernel 1x3x3 in 3d_fullres_nextou).
// As an implicit GEMM its K dimension is taps x Cin = 9: padding Cin to one 64-channel tensor-core slab wastes 98 % of the
// MMA work and of the activation traffic (measured: 216 us forward, 295 us weight gradient on the 64x224x192 patch, both
// far above what their 265 MB of HBM traffic needs).  These are plain CUDA-core kernels instead: the layer is a streaming
// pass (read Cin values per voxel, write Cout), 2*taps*Cin*Cout = 594 FLOP per voxel.
//   forward        : one thread = one voxel x 8 output channels (one 16-byte store); weights (bf16-rounded like the tensor-core
//                    path rounds its operand packs) and bias live in shared memory
//   weight gradient: the same thread mapping keeps 8 x taps x Cin partial sums in registers over a grid-stride loop, then
//                    shared-memory and global fp32 atomics (dW is zero-filled by the caller)
#include "common.cuh"

namespace nextou {

constexpr int CS_MAX_CIN = 4;
constexpr int CS_RUN = 8;          // consecutive voxels along W handled by one thread (sliding tap window)

struct ConvSmallGeom {
  int B, D, H, W, kd, kh, kw, Cin, Cout, taps;
  int ldx, ldo;         // row pitches (elements) of x [V][ldx] and out / dy [V][ldo]
  int chunks;           // ldo / 8: 16-byte chunks per output row
  int runs_per_row;     // ceil(W / CS_RUN)
  unsigned n_runs;      // B * D * H * runs_per_row
};

// The tap window of one run: win[r][j][c] = x[(d + od_r, h + oh_r, w0 - KW/2 + j)][c] for the ROWS = kd*kh tap rows and the
// CS_RUN + KW - 1 columns the run's voxels touch (zero outside the volume = the convolution's zero padding).  Every value is
// used by up to KW voxels; the coordinates are decomposed once per run, with 32-bit arithmetic (the host checks the range).
template <int CIN, int ROWS, int KW>
__device__ __forceinline__ void load_window(const ConvSmallGeom& g, const __nv_bfloat16* __restrict__ x, unsigned run, int& w0,
                                            unsigned& v0, float (&win)[ROWS][CS_RUN + KW - 1][CIN]) {
  const unsigned rw = run % (unsigned)g.runs_per_row;
  unsigned t = run / (unsigned)g.runs_per_row;           // flat (b, d, h)
  const int h = (int)(t % (unsigned)g.H);
  const unsigned bd = t / (unsigned)g.H;
  const int d = (int)(bd % (unsigned)g.D);
  w0 = (int)rw * CS_RUN;
  v0 = t * (unsigned)g.W + (unsigned)w0;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int od = r / g.kh - g.kd / 2, oh = r % g.kh - g.kh / 2;
    const bool row_in = (unsigned)(d + od) < (unsigned)g.D && (unsigned)(h + oh) < (unsigned)g.H;
    const int base = ((int)v0 + (od * g.H + oh) * g.W - KW / 2) * g.ldx;
#pragma unroll
    for (int j = 0; j < CS_RUN + KW - 1; ++j) {
      const bool in = row_in && (unsigned)(w0 - KW / 2 + j) < (unsigned)g.W;
#pragma unroll
      for (int c = 0; c < CIN; ++c) win[r][j][c] = in ? __bfloat162float(x[base + j * g.ldx + c]) : 0.f;
    }
  }
}

// Forward: one thread = one run x one 8-channel chunk; its 8 x taps x CIN weights sit in registers for the whole run.
// sw: [ROWS*KW][CIN][ldo] fp32 (bf16-rounded values, like the tensor-core path rounds its operand packs), sb: [ldo].
template <int CIN, int ROWS, int KW>
__global__ void __launch_bounds__(128)
    conv_small_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                          __nv_bfloat16* __restrict__ out, ConvSmallGeom g) {
  extern __shared__ __align__(16) float smem[];
  constexpr int TAPS = ROWS * KW;
  float* sw = smem;
  float* sb = smem + TAPS * CIN * g.ldo;
  for (int i = threadIdx.x; i < TAPS * CIN * g.ldo; i += blockDim.x) {
    const int co = i % g.ldo, c = (i / g.ldo) % CIN, t = i / (g.ldo * CIN);
    sw[i] = co < g.Cout ? __bfloat162float(__float2bfloat16_rn(w[(co * CIN + c) * TAPS + t])) : 0.f;
  }
  for (int i = threadIdx.x; i < g.ldo; i += blockDim.x) sb[i] = (bias != nullptr && i < g.Cout) ? bias[i] : 0.f;
  __syncthreads();
  const unsigned items = g.n_runs * (unsigned)g.chunks;
  for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
    const unsigned run = it / (unsigned)g.chunks;
    const int c0 = (int)(it - run * (unsigned)g.chunks) * 8;
    float wreg[TAPS * CIN][8], breg[8];
#pragma unroll
    for (int q = 0; q < TAPS * CIN; ++q) {
      const float4 w0 = *reinterpret_cast<const float4*>(sw + q * g.ldo + c0);
      const float4 w1 = *reinterpret_cast<const float4*>(sw + q * g.ldo + c0 + 4);
      wreg[q][0] = w0.x; wreg[q][1] = w0.y; wreg[q][2] = w0.z; wreg[q][3] = w0.w;
      wreg[q][4] = w1.x; wreg[q][5] = w1.y; wreg[q][6] = w1.z; wreg[q][7] = w1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) breg[j] = sb[c0 + j];
    float win[ROWS][CS_RUN + KW - 1][CIN];
    int w0;
    unsigned v0;
    load_window<CIN, ROWS, KW>(g, x, run, w0, v0, win);
#pragma unroll
    for (int l = 0; l < CS_RUN; ++l) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = breg[j];
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int k = 0; k < KW; ++k)
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            const float xv = win[r][l + k][c];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, wreg[(r * KW + k) * CIN + c][j], acc[j]);
          }
      if (w0 + l < g.W) {
        uint4 o;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[0], acc[1]), p1 = __floats2bfloat162_rn(acc[2], acc[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(acc[4], acc[5]), p3 = __floats2bfloat162_rn(acc[6], acc[7]);
        o.x = *reinterpret_cast<unsigned*>(&p0); o.y = *reinterpret_cast<unsigned*>(&p1);
        o.z = *reinterpret_cast<unsigned*>(&p2); o.w = *reinterpret_cast<unsigned*>(&p3);
        *reinterpret_cast<uint4*>(out + (size_t)(v0 + l) * g.ldo + c0) = o;      // columns >= Cout come out as exact zeros
      }
    }
  }
}

// Weight gradient: dW[co][tap][ci] (fp32) += sum_v dy[v][co] * x[v + tap][ci].  Block = a whole number of warps per 8-channel
// chunk (<= 320 threads): warp w owns chunk w % chunks, its lanes walk consecutive runs; the 8 x taps x CIN partial sums stay
// in registers over the grid-stride loop and are reduced with warp shuffles, then one global atomic per warp and value.
template <int CIN, int ROWS, int KW>
__global__ void __launch_bounds__(320)
    conv_small_wgrad_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, float* __restrict__ dW,
                            int cin_stride, ConvSmallGeom g) {
  constexpr int TAPS = ROWS * KW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = (warp % g.chunks) * 8;
  const unsigned runs_per_cta = (blockDim.x >> 5) / g.chunks * 32;
  float acc[8][TAPS * CIN];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < TAPS * CIN; ++q) acc[j][q] = 0.f;
  for (unsigned run = blockIdx.x * runs_per_cta + (warp / g.chunks) * 32 + lane; run < g.n_runs; run += gridDim.x * runs_per_cta) {
    float win[ROWS][CS_RUN + KW - 1][CIN];
    int w0;
    unsigned v0;
    load_window<CIN, ROWS, KW>(g, x, run, w0, v0, win);
    uint4 raw[CS_RUN];                                    // all dY vectors of the run in flight before any arithmetic
#pragma unroll
    for (int l = 0; l < CS_RUN; ++l)
      raw[l] = (w0 + l < g.W) ? *reinterpret_cast<const uint4*>(dy + (size_t)(v0 + l) * g.ldo + c0) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int l = 0; l < CS_RUN; ++l) {
      float d[8];
      const unsigned u[4] = {raw[l].x, raw[l].y, raw[l].z, raw[l].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        d[2 * j] = __uint_as_float(u[j] << 16);
        d[2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u);
      }
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int k = 0; k < KW; ++k)
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            const float xv = win[r][l + k][c];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j][(r * KW + k) * CIN + c] = fmaf(d[j], xv, acc[j][(r * KW + k) * CIN + c]);
          }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < TAPS * CIN; ++q) {
      float s = acc[j][q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0 && c0 + j < g.Cout) atomicAdd(dW + ((long long)(c0 + j) * TAPS + q / CIN) * cin_stride + q % CIN, s);
    }
}

static int small_geom(ConvSmallGeom& g, int B, int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, long long ldx,
                      long long ldo, const char* who) {
  NEXTOU_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "%s: bad shape", who);
  NEXTOU_REQUIRE(Cin >= 1 && Cin <= CS_MAX_CIN && Cout >= 1 && Cout <= 64, "%s: needs Cin <= %d and Cout <= 64 (got %d -> %d)", who,
                 CS_MAX_CIN, Cin, Cout);
  NEXTOU_REQUIRE(kd % 2 == 1 && kh % 2 == 1 && kw % 2 == 1, "%s: odd kernel sizes only", who);
  NEXTOU_REQUIRE(ldx >= Cin && ldo >= Cout && ldo % 8 == 0 && ldo <= 64, "%s: output pitch must be a multiple of 8 (<= 64)", who);
  NEXTOU_REQUIRE((long long)B * D * H * (W + CS_RUN) * (ldx > ldo ? ldx : ldo) < 2147483647LL, "%s: volume too large for 32-bit indexing", who);
  g.B = B; g.D = D; g.H = H; g.W = W; g.kd = kd; g.kh = kh; g.kw = kw; g.Cin = Cin; g.Cout = Cout; g.taps = kd * kh * kw;
  g.ldx = (int)ldx; g.ldo = (int)ldo; g.chunks = (int)(ldo / 8);
  g.runs_per_row = (W + CS_RUN - 1) / CS_RUN;
  g.n_runs = (unsigned)((long long)B * D * H * g.runs_per_row);
  return 0;
}

}  // namespace nextou

using namespace nextou;

// (Cin, tap rows = kd*kh, kw) combinations with taps * Cin <= 12 partial sums per output channel: 1x3x3 / 3x3 kernels on one
// modality, 3-tap and 1-tap kernels on up to four; everything else keeps the tensor-core path
#define CS_DISPATCH(cin, rows, kwid, ...)                                                                  \
  if ((rows) == 3 && (kwid) == 3 && (cin) == 1) { constexpr int CIN = 1, ROWS = 3, KW = 3; __VA_ARGS__ }   \
  else if ((rows) == 1 && (kwid) == 3 && (cin) == 1) { constexpr int CIN = 1, ROWS = 1, KW = 3; __VA_ARGS__ } \
  else if ((rows) == 1 && (kwid) == 3 && (cin) == 2) { constexpr int CIN = 2, ROWS = 1, KW = 3; __VA_ARGS__ } \
  else if ((rows) == 1 && (kwid) == 3 && (cin) == 3) { constexpr int CIN = 3, ROWS = 1, KW = 3; __VA_ARGS__ } \
  else if ((rows) == 1 && (kwid) == 3 && (cin) == 4) { constexpr int CIN = 4, ROWS = 1, KW = 3; __VA_ARGS__ } \
  else if ((rows) == 3 && (kwid) == 1 && (cin) == 1) { constexpr int CIN = 1, ROWS = 3, KW = 1; __VA_ARGS__ } \
  else if ((rows) == 3 && (kwid) == 1 && (cin) == 2) { constexpr int CIN = 2, ROWS = 3, KW = 1; __VA_ARGS__ } \
  else if ((rows) == 3 && (kwid) == 1 && (cin) == 3) { constexpr int CIN = 3, ROWS = 3, KW = 1; __VA_ARGS__ } \
  else if ((rows) == 3 && (kwid) == 1 && (cin) == 4) { constexpr int CIN = 4, ROWS = 3, KW = 1; __VA_ARGS__ } \
  else if ((rows) == 1 && (kwid) == 1 && (cin) == 1) { constexpr int CIN = 1, ROWS = 1, KW = 1; __VA_ARGS__ } \
  else if ((rows) == 1 && (kwid) == 1 && (cin) == 2) { constexpr int CIN = 2, ROWS = 1, KW = 1; __VA_ARGS__ } \
  else if ((rows) == 1 && (kwid) == 1 && (cin) == 3) { constexpr int CIN = 3, ROWS = 1, KW = 1; __VA_ARGS__ } \
  else if ((rows) == 1 && (kwid) == 1 && (cin) == 4) { constexpr int CIN = 4, ROWS = 1, KW = 1; __VA_ARGS__ } \
  else {                                                                                                   \
    set_error("conv3d_small_cin: %d tap rows x %d x %d input channels not instantiated", (rows), (kwid), (cin)); \
    return NEXTOU_ERR_UNSUPPORTED;                                                                         \
  }

// 1 if nextou_conv3d_small_cin_{fwd,wgrad} cover this layer (callers route everything else to the tensor-core kernels)
extern "C" int nextou_conv3d_small_cin_supported(int Cin, int Cout, int kd, int kh, int kw, long long ldo) {
  const int rows = kd * kh;
  const bool combo = (rows == 3 && kw == 3 && Cin == 1) || (((rows == 1 && kw == 3) || (rows == 3 && kw == 1) || (rows == 1 && kw == 1)) &&
                                                            Cin >= 1 && Cin <= 4);
  return combo && Cout <= 64 && kd % 2 == 1 && kh % 2 == 1 && kw % 2 == 1 && ldo % 8 == 0 && ldo <= 64;
}

// out[v][0..ldo) = bias + sum_{tap, ci} x[v + tap - pad][ci] * bf16(w[co][ci][tap]);  x, out bf16 token-major; w fp32 master weight
// (Cout, Cin, kd, kh, kw) as nn.Conv stores it; columns [Cout, ldo) are written as zeros.
extern "C" int nextou_conv3d_small_cin_fwd(const void* x, long long ldx, int B, int D, int H, int W, int Cin, const float* w,
                                           int Cout, int kd, int kh, int kw, const float* bias, void* out, long long ldo,
                                           void* stream) {
  NEXTOU_REQUIRE(x && w && out, "conv3d_small_cin_fwd: null pointer");
  NEXTOU_REQUIRE(((uintptr_t)out & 15) == 0, "conv3d_small_cin_fwd: 16-byte alignment");
  ConvSmallGeom g;
  int rc = small_geom(g, B, D, H, W, Cin, Cout, kd, kh, kw, ldx, ldo, "conv3d_small_cin_fwd");
  if (rc) return rc;
  const size_t smem = sizeof(float) * ((size_t)g.taps * Cin * ldo + ldo);
  long long blocks = ((long long)g.n_runs * g.chunks + 127) / 128;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  CS_DISPATCH(Cin, kd * kh, kw, (conv_small_fwd_kernel<CIN, ROWS, KW><<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(
                                    (const __nv_bfloat16*)x, w, bias, (__nv_bfloat16*)out, g));)
  return check_launch("conv_small_fwd_kernel");
}

// dW[Cout][taps][cin_stride] (fp32, zero-filled by the caller) += sum_v dy[v][co] * x[v + tap - pad][ci]
extern "C" int nextou_conv3d_small_cin_wgrad(const void* dy, long long ldy, const void* x, long long ldx, int B, int D, int H,
                                             int W, int Cin, int Cout, int kd, int kh, int kw, float* dW, int cin_stride,
                                             void* stream) {
  NEXTOU_REQUIRE(dy && x && dW && cin_stride >= Cin, "conv3d_small_cin_wgrad: bad arguments");
  NEXTOU_REQUIRE(((uintptr_t)dy & 15) == 0, "conv3d_small_cin_wgrad: 16-byte alignment");
  ConvSmallGeom g;
  int rc = small_geom(g, B, D, H, W, Cin, Cout, kd, kh, kw, ldx, ldy, "conv3d_small_cin_wgrad");
  if (rc) return rc;
  const int wpc = g.chunks <= 5 ? 2 : 1;                  // warps per 8-channel chunk: block <= 320 threads (chunks <= 8)
  const int threads = 32 * wpc * g.chunks;
  long long blocks = ((long long)g.n_runs + 32 * wpc - 1) / (32 * wpc);
  if (blocks > 2LL * num_sms()) blocks = 2LL * num_sms();
  CS_DISPATCH(Cin, kd * kh, kw, (conv_small_wgrad_kernel<CIN, ROWS, KW><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
                                    (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, dW, cin_stride, g));)
  return check_launch("conv_small_wgrad_kernel");
}
