// Pool-GNN down/up-sampling on token-major (NDHWC) activations.
// Replaces nn.MaxPool3d(return_indices) / nn.MaxUnpool3d / F.avg_pool3d in PoolDyGraphConv.forward
// (reference network_architecture/NexToU_Encoder_Decoder.py:511-512, 524-528, 536-549) and their backward.
// All pools are non-overlapping (kernel == stride).  The arg-max is stored as the uint8 CHILD index
// (dz*ph*pw + dy*pw + dx) instead of torch's int64 flat voxel index; ties keep the first child in
// (dz, dy, dx) scan order like ATen's max_pool3d_with_indices.
// HBM-bound elementwise kernels: one thread per (voxel, channel), channels fastest (coalesced).
#include "common.cuh"

namespace nextou {

struct PoolGeom {
  int B, D, H, W, pd, ph, pw, Dp, Hp, Wp;
};

// pooled voxel p (flat over B,Dp,Hp,Wp) + child index -> flat input voxel
__device__ __forceinline__ long long child_voxel(const PoolGeom& g, long long p, int child) {
  const int xw = (int)(p % g.Wp);
  long long t = p / g.Wp;
  const int yh = (int)(t % g.Hp);
  t /= g.Hp;
  const int zd = (int)(t % g.Dp);
  const int b = (int)(t / g.Dp);
  const int dx = child % g.pw, dy = (child / g.pw) % g.ph, dz = child / (g.pw * g.ph);
  return (((long long)b * g.D + zd * g.pd + dz) * g.H + yh * g.ph + dy) * g.W + xw * g.pw + dx;
}
// input voxel v -> pooled voxel + child
__device__ __forceinline__ void parent_of(const PoolGeom& g, long long v, long long& p, int& child) {
  const int x = (int)(v % g.W);
  long long t = v / g.W;
  const int y = (int)(t % g.H);
  t /= g.H;
  const int z = (int)(t % g.D);
  const int b = (int)(t / g.D);
  p = (((long long)b * g.Dp + z / g.pd) * g.Hp + y / g.ph) * g.Wp + x / g.pw;
  child = ((z % g.pd) * g.ph + (y % g.ph)) * g.pw + (x % g.pw);
}

template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, long long ldx, int C, PoolGeom g, T* __restrict__ out,
                                   long long ldo, uint8_t* __restrict__ arg, long long total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  const long long p = t / C;
  const int nchild = g.pd * g.ph * g.pw;
  float best = to_f(x[child_voxel(g, p, 0) * ldx + c]);
  int bi = 0;
  for (int ch = 1; ch < nchild; ++ch) {
    const float v = to_f(x[child_voxel(g, p, ch) * ldx + c]);
    if (v > best || (v != v && best == best)) {  // NaN propagates like ATen
      best = v;
      bi = ch;
    }
  }
  out[p * ldo + c] = from_f<T>(best);
  arg[p * C + c] = (uint8_t)bi;
}

template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ dout, long long ldo, const uint8_t* __restrict__ arg, int C,
                                   PoolGeom g, T* __restrict__ dx, long long ldx, long long total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  const long long v = t / C;
  long long p;
  int child;
  parent_of(g, v, p, child);
  dx[v * ldx + c] = (arg[p * C + c] == child) ? dout[p * ldo + c] : from_f<T>(0.f);
}

template <typename T>
__global__ void avgpool_fwd_kernel(const T* __restrict__ x, long long ldx, int C, PoolGeom g, T* __restrict__ out,
                                   long long ldo, long long total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  const long long p = t / C;
  const int nchild = g.pd * g.ph * g.pw;
  float s = 0.f;
  for (int ch = 0; ch < nchild; ++ch) s = __fadd_rn(s, to_f(x[child_voxel(g, p, ch) * ldx + c]));
  out[p * ldo + c] = from_f<T>(__fdiv_rn(s, (float)nchild));
}

template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dout, long long ldo, int C, PoolGeom g, T* __restrict__ dx,
                                   long long ldx, long long total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  const long long v = t / C;
  long long p;
  int child;
  parent_of(g, v, p, child);
  dx[v * ldx + c] = from_f<T>(__fdiv_rn(to_f(dout[p * ldo + c]), (float)(g.pd * g.ph * g.pw)));
}

// unpool: out[v][j] = (arg[parent(v)][j % Carg] == child(v)) ? gsrc[parent(v)][j] : 0      (ED:536-549)
template <typename T>
__global__ void maxunpool_fwd_kernel(const T* __restrict__ gsrc, long long ldg, const uint8_t* __restrict__ arg,
                                     int Carg, int C2, PoolGeom g, T* __restrict__ out, long long ldo,
                                     long long total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int j = (int)(t % C2);
  const long long v = t / C2;
  long long p;
  int child;
  parent_of(g, v, p, child);
  out[v * ldo + j] = (arg[p * Carg + (j % Carg)] == child) ? gsrc[p * ldg + j] : from_f<T>(0.f);
}

template <typename T>
__global__ void maxunpool_bwd_kernel(const T* __restrict__ dout, long long ldo, const uint8_t* __restrict__ arg,
                                     int Carg, int C2, PoolGeom g, T* __restrict__ dg, long long ldg,
                                     long long total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int j = (int)(t % C2);
  const long long p = t / C2;
  const long long v = child_voxel(g, p, arg[p * Carg + (j % Carg)]);
  dg[p * ldg + j] = dout[v * ldo + j];
}

// ---- 16-byte-vector variants (VEC channels of one voxel per thread: one set of index divisions per vector) ----
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T>
__global__ void maxunpool_fwd_vec_kernel(const T* __restrict__ gsrc, long long ldg, const uint8_t* __restrict__ arg,
                                         int Carg, int C2, PoolGeom g, T* __restrict__ out, long long ldo, long long total) {
  constexpr int V = Vec16<T>::N;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int groups = C2 / V;
  const int j0 = (int)(t % groups) * V;
  const long long v = t / groups;
  long long p;
  int child;
  parent_of(g, v, p, child);
  uint4 src = *reinterpret_cast<const uint4*>(gsrc + p * ldg + j0);
  T* e = reinterpret_cast<T*>(&src);
  const uint8_t* a = arg + p * Carg;
  int ja = j0 % Carg;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    if (a[ja] != child) e[i] = from_f<T>(0.f);
    if (++ja == Carg) ja = 0;
  }
  *reinterpret_cast<uint4*>(out + v * ldo + j0) = src;
}

template <typename T>
__global__ void maxunpool_bwd_vec_kernel(const T* __restrict__ dout, long long ldo, const uint8_t* __restrict__ arg,
                                         int Carg, int C2, PoolGeom g, T* __restrict__ dg, long long ldg, long long total) {
  constexpr int V = Vec16<T>::N;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int groups = C2 / V;
  const int j0 = (int)(t % groups) * V;
  const long long p = t / groups;
  const long long v0 = child_voxel(g, p, 0);          // children differ from child 0 by a fixed voxel offset
  const uint8_t* a = arg + p * Carg;
  int ja = j0 % Carg;
  uint4 res;
  T* e = reinterpret_cast<T*>(&res);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int ch = a[ja];
    const int dx = ch % g.pw, dy = (ch / g.pw) % g.ph, dz = ch / (g.pw * g.ph);
    const long long v = v0 + ((long long)dz * g.H + dy) * g.W + dx;
    e[i] = dout[v * ldo + j0 + i];
    if (++ja == Carg) ja = 0;
  }
  *reinterpret_cast<uint4*>(dg + p * ldg + j0) = res;
}

template <typename T>
static inline bool vec_ok(const void* a, long long lda, const void* b, long long ldb) {
  constexpr int V = Vec16<T>::N;
  return lda % V == 0 && ldb % V == 0 && ((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0;
}

static int make_geom(PoolGeom& g, int B, int D, int H, int W, int pd, int ph, int pw) {
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || pd <= 0 || ph <= 0 || pw <= 0) {
    set_error("pool: bad geometry B=%d D=%d H=%d W=%d p=(%d,%d,%d)", B, D, H, W, pd, ph, pw);
    return NEXTOU_ERR_INVALID;
  }
  if (D % pd || H % ph || W % pw || pd * ph * pw > 255) {
    set_error("pool: volume (%d,%d,%d) not divisible by pool (%d,%d,%d)", D, H, W, pd, ph, pw);
    return NEXTOU_ERR_INVALID;
  }
  g = PoolGeom{B, D, H, W, pd, ph, pw, D / pd, H / ph, W / pw};
  return 0;
}

static inline unsigned blocks_for(long long total) { return (unsigned)((total + 255) / 256); }

}  // namespace nextou

using namespace nextou;

extern "C" int nextou_maxpool3d_fwd(const void* x, int dtype, long long ldx, int C, int B, int D, int H, int W, int pd,
                                    int ph, int pw, void* out, long long ldo, uint8_t* arg, void* stream) {
  PoolGeom g;
  int rc = make_geom(g, B, D, H, W, pd, ph, pw);
  if (rc) return rc;
  NEXTOU_REQUIRE(x && out && arg && C > 0, "maxpool3d_fwd: bad args");
  const long long total = (long long)B * g.Dp * g.Hp * g.Wp * C;
  DISPATCH_T(dtype, maxpool_fwd_kernel<T><<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)x, ldx, C, g, (T*)out, ldo, arg, total);)
  return check_launch("maxpool_fwd_kernel");
}

extern "C" int nextou_maxpool3d_bwd(const void* dout, int dtype, long long ldo, const uint8_t* arg, int C, int B, int D,
                                    int H, int W, int pd, int ph, int pw, void* dx, long long ldx, void* stream) {
  PoolGeom g;
  int rc = make_geom(g, B, D, H, W, pd, ph, pw);
  if (rc) return rc;
  NEXTOU_REQUIRE(dout && dx && arg && C > 0, "maxpool3d_bwd: bad args");
  const long long total = (long long)B * D * H * W * C;
  DISPATCH_T(dtype, maxpool_bwd_kernel<T><<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)dout, ldo, arg, C, g, (T*)dx, ldx, total);)
  return check_launch("maxpool_bwd_kernel");
}

extern "C" int nextou_avgpool3d_fwd(const void* x, int dtype, long long ldx, int C, int B, int D, int H, int W, int pd,
                                    int ph, int pw, void* out, long long ldo, void* stream) {
  PoolGeom g;
  int rc = make_geom(g, B, D, H, W, pd, ph, pw);
  if (rc) return rc;
  NEXTOU_REQUIRE(x && out && C > 0, "avgpool3d_fwd: bad args");
  const long long total = (long long)B * g.Dp * g.Hp * g.Wp * C;
  DISPATCH_T(dtype, avgpool_fwd_kernel<T><<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)x, ldx, C, g, (T*)out, ldo, total);)
  return check_launch("avgpool_fwd_kernel");
}

extern "C" int nextou_avgpool3d_bwd(const void* dout, int dtype, long long ldo, int C, int B, int D, int H, int W,
                                    int pd, int ph, int pw, void* dx, long long ldx, void* stream) {
  PoolGeom g;
  int rc = make_geom(g, B, D, H, W, pd, ph, pw);
  if (rc) return rc;
  NEXTOU_REQUIRE(dout && dx && C > 0, "avgpool3d_bwd: bad args");
  const long long total = (long long)B * D * H * W * C;
  DISPATCH_T(dtype, avgpool_bwd_kernel<T><<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)dout, ldo, C, g, (T*)dx, ldx, total);)
  return check_launch("avgpool_bwd_kernel");
}

extern "C" int nextou_maxunpool3d_fwd(const void* gsrc, int dtype, long long ldg, const uint8_t* arg, int Carg, int C2,
                                      int B, int D, int H, int W, int pd, int ph, int pw, void* out, long long ldo,
                                      void* stream) {
  PoolGeom g;
  int rc = make_geom(g, B, D, H, W, pd, ph, pw);
  if (rc) return rc;
  NEXTOU_REQUIRE(gsrc && out && arg && Carg > 0 && C2 > 0, "maxunpool3d_fwd: bad args");
  const long long total = (long long)B * D * H * W * C2;
  DISPATCH_T(dtype, {
    if (C2 % Vec16<T>::N == 0 && vec_ok<T>(gsrc, ldg, out, ldo)) {
      const long long tv = total / Vec16<T>::N;
      maxunpool_fwd_vec_kernel<T><<<blocks_for(tv), 256, 0, (cudaStream_t)stream>>>((const T*)gsrc, ldg, arg, Carg, C2, g,
                                                                                     (T*)out, ldo, tv);
    } else {
      maxunpool_fwd_kernel<T><<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)gsrc, ldg, arg, Carg, C2, g,
                                                                                    (T*)out, ldo, total);
    }
  })
  return check_launch("maxunpool_fwd_kernel");
}

extern "C" int nextou_maxunpool3d_bwd(const void* dout, int dtype, long long ldo, const uint8_t* arg, int Carg, int C2,
                                      int B, int D, int H, int W, int pd, int ph, int pw, void* dg, long long ldg,
                                      void* stream) {
  PoolGeom g;
  int rc = make_geom(g, B, D, H, W, pd, ph, pw);
  if (rc) return rc;
  NEXTOU_REQUIRE(dout && dg && arg && Carg > 0 && C2 > 0, "maxunpool3d_bwd: bad args");
  const long long total = (long long)B * g.Dp * g.Hp * g.Wp * C2;
  DISPATCH_T(dtype, {
    if (C2 % Vec16<T>::N == 0 && vec_ok<T>(dg, ldg, dg, ldg)) {
      const long long tv = total / Vec16<T>::N;
      maxunpool_bwd_vec_kernel<T><<<blocks_for(tv), 256, 0, (cudaStream_t)stream>>>((const T*)dout, ldo, arg, Carg, C2, g,
                                                                                     (T*)dg, ldg, tv);
    } else {
      maxunpool_bwd_kernel<T><<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)dout, ldo, arg, Carg, C2, g,
                                                                                    (T*)dg, ldg, total);
    }
  })
  return check_launch("maxunpool_bwd_kernel");
}

// ------------------------------------------------------------------------------------------------------
// Row copy / add on token-major matrices whose views differ in pitch or column offset:
//   out[r][0..cols) = a[r][0..cols) (+ b[r][0..cols))
// The two places where the U-Net topology needs a data movement that no GEMM epilogue can absorb: the skip half of
// torch.cat((up, skip), 1) (NexToU_Encoder_Decoder.py:322) written behind the up-sampled half of the concatenation buffer,
// and the sum of the two gradients of a tensor that is consumed twice (skip connection / residual shortcut).  `cols` counts
// 16-byte vectors' worth of columns: callers round it up to the channel padding (padding lanes are don't-care).
// ------------------------------------------------------------------------------------------------------
namespace nextou {
template <bool ADD>
__global__ void __launch_bounds__(256) rows_copy_add_kernel(const uint4* __restrict__ a, long long lda, const uint4* __restrict__ b,
                                                            long long ldb, uint4* __restrict__ out, long long ldo,
                                                            long long rows, int vec_per_row, int is_bf16) {
  const long long total = rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const int v = (int)(i - r * vec_per_row);
    uint4 x = a[r * lda + v];
    if (ADD) {
      const uint4 y = b[r * ldb + v];
      if (is_bf16) {
        __nv_bfloat162* xp = reinterpret_cast<__nv_bfloat162*>(&x);
        const __nv_bfloat162* yp = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
        for (int e = 0; e < 4; ++e) xp[e] = __hadd2(xp[e], yp[e]);      // one rounding per element, like at::add on bf16
      } else {
        float* xp = reinterpret_cast<float*>(&x);
        const float* yp = reinterpret_cast<const float*>(&y);
#pragma unroll
        for (int e = 0; e < 4; ++e) xp[e] += yp[e];
      }
    }
    out[r * ldo + v] = x;
  }
}
}  // namespace nextou

// a, b (may be NULL: plain copy), out: [rows] rows of `cols` elements at pitches lda / ldb / ldo (elements).  Bases and
// pitches must be 16-byte aligned and cols a multiple of 16 bytes (the channel-padded token layout guarantees both).
extern "C" int nextou_rows_copy_add(const void* a, long long lda, const void* b, long long ldb, void* out, long long ldo,
                                    long long rows, int cols, int dtype, void* stream) {
  NEXTOU_REQUIRE(a && out && rows > 0 && cols > 0, "rows_copy_add: bad arguments");
  NEXTOU_REQUIRE(dtype == NEXTOU_BF16 || dtype == NEXTOU_F32, "rows_copy_add: bad dtype");
  const int per = dtype == NEXTOU_BF16 ? 8 : 4;            // elements per 16-byte vector
  NEXTOU_REQUIRE(cols % per == 0 && lda % per == 0 && ldo % per == 0 && (b == nullptr || ldb % per == 0) && lda >= cols &&
                     ldo >= cols && (b == nullptr || ldb >= cols),
                 "rows_copy_add: pitches / cols must be multiples of 16 bytes");
  NEXTOU_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)b & 15) == 0, "rows_copy_add: 16-byte alignment");
  const int vpr = cols / per;
  const long long total = rows * vpr;
  long long blocks = (total + 255) / 256;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  cudaStream_t st = (cudaStream_t)stream;
  if (b != nullptr)
    rows_copy_add_kernel<true><<<(unsigned)blocks, 256, 0, st>>>((const uint4*)a, lda / per, (const uint4*)b, ldb / per, (uint4*)out,
                                                                ldo / per, rows, vpr, dtype == NEXTOU_BF16);
  else
    rows_copy_add_kernel<false><<<(unsigned)blocks, 256, 0, st>>>((const uint4*)a, lda / per, nullptr, 0, (uint4*)out, ldo / per,
                                                                 rows, vpr, dtype == NEXTOU_BF16);
  return check_launch("rows_copy_add_kernel");
}

// ------------------------------------------------------------------------------------------------------
// Weight packing: one launch turns an fp32 / bf16 master weight w[R][Cc/groups][taps] (nn.Conv / nn.ConvTranspose layout,
// taps = prod(kernel)) into BOTH bf16 operand packs the tcgen05 kernels take:
//   A[r][t][c]  (row pitch taps*lda_c)  = w[r][c][t]          forward operand   ([Cout][taps][Cin pad])
//   Bt[c][t'][r] (row pitch taps*ldb_c) = w[r][c][t]          data-gradient operand ([Cin][taps][Cout pad]), t' = flipped t
// zero-filled where c >= Cc / r >= R (channel padding) and, for grouped layers, outside the diagonal blocks.
// Replaces the reshape / permute / pad / cast / block_diag chains of ATen ops per layer and step.
// ------------------------------------------------------------------------------------------------------
namespace nextou {
template <typename T>
__global__ void pack_weight_kernel(const NextouPackJob j, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    pack_element(j, reinterpret_cast<const T*>(j.w), i);
}
}  // namespace nextou

extern "C" int nextou_pack_weight(const void* w, int dtype, int R, int Cc, int taps, int groups, int flip_b, void* A, int lda_c,
                                  void* Bt, int ldb_c, void* stream) {
  return nextou_pack_weight_gap(w, dtype, R, Cc, taps, groups, flip_b, 0, 0, A, lda_c, Bt, ldb_c, stream);
}

// Same with a zero channel gap [gap_lo, gap_hi) in the input-channel axis of both packs (lda_c >= Cc + gap width; Bt gets
// Cc + gap width rows): the weight of the first convolution of a decoder stage, whose input is the [up | gap | skip]
// concatenation buffer (NexToU_Encoder_Decoder.py:322), without materialising a zero-padded copy of the weight.
extern "C" int nextou_pack_weight_gap(const void* w, int dtype, int R, int Cc, int taps, int groups, int flip_b, int gap_lo,
                                      int gap_hi, void* A, int lda_c, void* Bt, int ldb_c, void* stream) {
  NEXTOU_REQUIRE(w && A, "pack_weight: null pointer");
  NEXTOU_REQUIRE(gap_lo >= 0 && gap_hi >= gap_lo && gap_lo <= Cc && (gap_hi == gap_lo || groups == 1), "pack_weight: bad channel gap");
  const int Cp = Cc + gap_hi - gap_lo;
  NEXTOU_REQUIRE(R > 0 && Cc > 0 && taps > 0 && groups > 0 && R % groups == 0 && Cc % groups == 0 && lda_c >= Cp &&
                     (Bt == nullptr || ldb_c >= R), "pack_weight: bad shape");
  NextouPackJob j = {};
  j.w = w; j.A = A; j.Bt = Bt; j.R = R; j.Cc = Cc; j.taps = taps; j.groups = groups; j.flip_b = flip_b;
  j.lda_c = lda_c; j.ldb_c = ldb_c; j.gap_lo = gap_lo; j.gap_hi = gap_hi;
  const long long n = (long long)R * taps * lda_c + (Bt ? (long long)Cp * taps * ldb_c : 0);
  long long blocks = (n + 255) / 256;
  if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
  DISPATCH_T(dtype, pack_weight_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(j, n);)
  return check_launch("pack_weight_kernel");
}
