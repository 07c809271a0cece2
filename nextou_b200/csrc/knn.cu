// kNN graph build for the Pool-/Swin-GNN blocks: row normalisation + fused distance-tile / running top-k.
// Replaces DenseDilatedKnnGraph.forward and the dense_knn_matrix helpers
// (reference network_architecture/torch_edge.py:12-163).  The N x M distance matrix never
// leaves the SM: a CTA owns BM query rows, walks the candidates in BN-wide chunks, keeps the
// fp32 distance tile in shared memory and folds it into a per-row sorted top-32 list with a
// warp-shuffle bitonic sort + merge.
//
// Arithmetic contract (shared with oracle/c/nextou_oracle.c, which it must match BIT FOR BIT):
//   * sums over channels in the normalise step: lane l accumulates c = l, l+32, ... with fmaf,
//     then a 16/8/4/2/1 xor-butterfly of plain adds;
//   * dot(x_i, y_j): one accumulator, acc = fmaf(x[c], y[c], acc) for c = 0..C-1 in order;
//   * dist = ((sqx_i + (-2*acc)) + sqy_j) + relpos[i][j], each op rounded to fp32 (TE:21-23,86);
//   * order: ascending (dist, j) — ties go to the lowest candidate index.
#include "common.cuh"
#include <limits.h>

namespace nextou {

// ------------------------------------------------------------------------------------------
// Step 1: normalise rows, transpose to channel-major
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float butterfly_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

__global__ void __launch_bounds__(256) knn_normalize_kernel(const void* __restrict__ x, int dtype, long long ldx,
                                                            long long bstride, const int32_t* __restrict__ row_map,
                                                            int N, int C, int normalize, float* __restrict__ xn,
                                                            int ldn, float* __restrict__ sq) {
  extern __shared__ float tile[];  // [32][CS]
  const int CS = C | 1;            // odd row stride -> conflict-free transposed reads
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = warp * 4 + i;
    const int n = n0 + r;
    float* trow = tile + r * CS;
    if (n < N) {
      const long long src = row_map ? (long long)row_map[(long long)b * N + n] * ldx : (long long)b * bstride + (long long)n * ldx;
      float ss = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float v = ld_as_f32(x, dtype, src + c);
        trow[c] = v;
        ss = fmaf(v, v, ss);
      }
      ss = butterfly_sum(ss);
      const float denom = fmaxf(sqrtf(ss), 1e-12f);
      float t = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float v = normalize ? __fdiv_rn(trow[c], denom) : trow[c];
        trow[c] = v;
        t = fmaf(v, v, t);
      }
      t = butterfly_sum(t);
      if (lane == 0) sq[(long long)b * N + n] = t;
    } else {
      for (int c = lane; c < C; c += 32) trow[c] = 0.f;
    }
  }
  __syncthreads();
  const int n = n0 + lane;
  if (n < ldn) {
    for (int c = warp; c < C; c += 8) xn[((long long)b * C + c) * ldn + n] = tile[lane * CS + c];
  }
}

// ------------------------------------------------------------------------------------------
// Step 2: distance tile + running top-k
// ------------------------------------------------------------------------------------------
struct Cand {
  float d;
  int i;
};
__device__ __forceinline__ bool cand_less(float ad, int ai, float bd, int bi) {
  return (ad < bd) || (ad == bd && ai < bi);
}
// one compare-exchange stage of a 32-lane bitonic network
__device__ __forceinline__ void cmpex(float& d, int& i, int stride, bool keep_min) {
  const float od = __shfl_xor_sync(0xffffffffu, d, stride);
  const int oi = __shfl_xor_sync(0xffffffffu, i, stride);
  const bool other_less = cand_less(od, oi, d, i);
  const bool take = keep_min ? other_less : !other_less;
  if (take) {
    d = od;
    i = oi;
  }
}
__device__ __forceinline__ void bitonic_sort32(float& d, int& i, int lane) {
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
    for (int s = k2 >> 1; s > 0; s >>= 1) {
      const bool up = ((lane & k2) == 0);
      const bool lower = ((lane & s) == 0);
      cmpex(d, i, s, lower == up);
    }
  }
}
__device__ __forceinline__ void bitonic_merge32(float& d, int& i, int lane) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) cmpex(d, i, s, (lane & s) == 0);
}

constexpr int KNN_KC = 16;

constexpr int KNN_CAP = 48;   // candidates of one row and chunk that may tie with / beat the row threshold (else: slow path)

constexpr int KNN_STAGES = 3;  // cp.async ring of channel slabs: three stages need only one CTA barrier per slab

template <int BM, int BN, int TM, int TN>
struct KnnCfg {
  static constexpr int TX = BN / TN, TY = BM / TM, NT = TX * TY;
  static constexpr int DS = BN + 4;  // distance tile row stride (floats)
  // the operand ring is dead while a chunk's top-K runs: the candidate arrays (+ thresholds, counters) live in it
  static constexpr size_t ring_floats = (size_t)KNN_STAGES * KNN_KC * (BM + BN);
  static constexpr size_t cand_floats = 2 * (size_t)BM * KNN_CAP + 3 * (size_t)BM;
  static constexpr size_t smem_bytes =
      sizeof(float) * ((ring_floats > cand_floats ? ring_floats : cand_floats) + (size_t)BM * DS + 2 * (size_t)BM * 32);
};

// two fp32 FMAs per instruction (FFMA2): each half is an IEEE fma.rn, i.e. bit-identical to fmaf
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(unsigned long long& c, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}

// The running top-32 list of one row, merged with its BN chunk candidates by ONE FULL WARP (sorted across the lanes): a few
// qualifying candidates are inserted one by one (ballot -> position, shfl_up -> shift), many take the bitonic sort + merge.
// Slow path of knn_topk_kernel (rows whose threshold ties with more than KNN_CAP candidates) — the order (distance, index) is
// total, so both paths produce the same list.
template <int BN, int DS>
__device__ __forceinline__ void warp_row_topk(const float* __restrict__ Ds, int row, int j0, int M, int K, int lane, float& td,
                                              int& ti) {
#pragma unroll 1
  for (int s = 0; s < (BN + 31) / 32; ++s) {
    const int col = s * 32 + lane;
    const int gj = j0 + col;
    float d = __int_as_float(0x7f800000);
    int ci = INT_MAX;
    if (col < BN && gj < M) {
      d = Ds[row * DS + col];
      ci = gj;
    }
    const float thr_d = __shfl_sync(0xffffffffu, td, K - 1);
    const int thr_i = __shfl_sync(0xffffffffu, ti, K - 1);
    unsigned q = __ballot_sync(0xffffffffu, cand_less(d, ci, thr_d, thr_i));
    if (q == 0u) continue;
    if (__popc(q) > 10) {
      bitonic_sort32(d, ci, lane);
      const float rd = __shfl_sync(0xffffffffu, d, 31 - lane);
      const int ri = __shfl_sync(0xffffffffu, ci, 31 - lane);
      if (cand_less(rd, ri, td, ti)) {
        td = rd;
        ti = ri;
      }
      bitonic_merge32(td, ti, lane);
    } else {
      while (q) {
        const int src = __ffs(q) - 1;
        q &= q - 1;
        const float cd = __shfl_sync(0xffffffffu, d, src);
        const int cx = __shfl_sync(0xffffffffu, ci, src);
        const unsigned m = __ballot_sync(0xffffffffu, cand_less(cd, cx, td, ti));
        if (m == 0u) continue;
        const int pos = __ffs(m) - 1;          // the list is sorted: lanes >= pos hold larger entries
        const float ud = __shfl_up_sync(0xffffffffu, td, 1);
        const int ui = __shfl_up_sync(0xffffffffu, ti, 1);
        if (lane > pos) {
          td = ud;
          ti = ui;
        } else if (lane == pos) {
          td = cd;
          ti = cx;
        }
      }
    }
  }
}

// LK = 8 / 16 / 24 / 32: register list length of the per-row threshold search (>= k * dilation)
template <int BM, int BN, int TM, int TN, int LK>
__global__ void __launch_bounds__(KnnCfg<BM, BN, TM, TN>::NT, (TM * TN <= 32 ? 2 : 1))
    knn_topk_kernel(const float* __restrict__ xn, const float* __restrict__ sqx, int ldn,
                    const float* __restrict__ yn, const float* __restrict__ sqy, int ldm,
                    const float* __restrict__ relpos, int N, int M, int C, int k, int dilation,
                    int64_t* __restrict__ out, int32_t* __restrict__ out32, float* __restrict__ part_d,
                    int* __restrict__ part_i) {
  using Cfg = KnnCfg<BM, BN, TM, TN>;
  constexpr int KC = KNN_KC, TX = Cfg::TX, TY = Cfg::TY, NT = Cfg::NT, DS = Cfg::DS;
  extern __shared__ __align__(16) float smem[];
  constexpr int ST = KNN_STAGES;
  constexpr size_t RING = Cfg::ring_floats > Cfg::cand_floats ? Cfg::ring_floats : Cfg::cand_floats;
  float* Xs = smem;                    // [ST][KC][BM]
  float* Ys = Xs + ST * KC * BM;       // [ST][KC][BN]
  float* Ds = smem + RING;             // [BM][DS]
  float* Ld = Ds + BM * DS;            // [BM][32]
  int* Li = reinterpret_cast<int*>(Ld + BM * 32);
  float* Cd = smem;                                     // (aliases the ring) [BM][KNN_CAP] candidates of the chunk: distance,
  int* Ci = reinterpret_cast<int*>(Cd + BM * KNN_CAP);  //                                   index
  float* Thr = reinterpret_cast<float*>(Ci + BM * KNN_CAP);   // [BM] row threshold = K-th smallest distance so far
  int* Cnt = reinterpret_cast<int*>(Thr + BM);                // [BM] candidates <= threshold
  int* New = Cnt + BM;                                        // [BM] ... of which from this chunk (0: the running list stands)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid % TX, ty = tid / TX;
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * BM;
  const int K = k * dilation;
  const float* xb = xn + (long long)b * C * ldn;
  const float* yb = yn + (long long)b * C * ldm;
  const int nslab = (C + KC - 1) / KC;

  for (int t = tid; t < BM * 32; t += NT) {
    Ld[t] = __int_as_float(0x7f800000);
    Li[t] = INT_MAX;
  }

  // candidate range of this CTA: gridDim.z > 1 splits the BN-wide chunks over CTAs, each of which leaves a sorted
  // partial top-32 list per row in (part_d, part_i) for knn_merge_kernel
  const int nchunks = (M + BN - 1) / BN;
  const int cpz = (nchunks + gridDim.z - 1) / gridDim.z;
  const int jbeg = blockIdx.z * cpz * BN;
  const int jend = min(M, (int)(blockIdx.z + 1) * cpz * BN);
  const bool split = gridDim.z > 1;

  if (split && jbeg >= jend) {   // defensive: a CTA without candidates still owes the merge its (empty) lists
    __syncthreads();
    for (int t = tid; t < BM * 32; t += NT) {
      const int gi = i0 + t / 32;
      if (gi < N) {
        const long long o = (((long long)b * N + gi) * gridDim.z + blockIdx.z) * 32 + (t & 31);
        part_d[o] = Ld[t];
        part_i[o] = Li[t];
      }
    }
    return;
  }

  // cp.async slots of this thread: the (channel, 16-byte column group) pairs are the same for every slab
  constexpr int XCH = KC * BM / 4, YCH = KC * BN / 4;
  constexpr int NXS = (XCH + NT - 1) / NT, NYS = (YCH + NT - 1) / NT;
  int xs_off[NXS], xg_off[NXS], xkc[NXS], ys_off[NYS], yg_off[NYS], ykc[NYS];
#pragma unroll
  for (int n = 0; n < NXS; ++n) {
    const int g = tid + n * NT;
    const int kc = g / (BM / 4), q = g % (BM / 4);
    const bool ok = g < XCH && i0 + q * 4 < ldn;
    xkc[n] = ok ? kc : (1 << 30);            // channel index inside the slab (huge = never load)
    xs_off[n] = kc * BM + q * 4;
    xg_off[n] = kc * ldn + i0 + q * 4;
    if (g >= XCH) xs_off[n] = -1;
  }
#pragma unroll
  for (int n = 0; n < NYS; ++n) {
    const int g = tid + n * NT;
    const int kc = g / (BN / 4), q = g % (BN / 4);
    ykc[n] = kc;
    ys_off[n] = g < YCH ? kc * BN + q * 4 : -1;
    yg_off[n] = kc * ldm + q * 4;
  }

  for (int j0 = jbeg; j0 < jend; j0 += BN) {
    unsigned long long acc2[TM][TN / 2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN / 2; ++j) acc2[i][j] = 0ull;

    auto load_slab = [&](int s) {
      const int c0 = s * KC;
      float* xs = Xs + (s % ST) * KC * BM;
      float* ys = Ys + (s % ST) * KC * BN;
      const float* xsrc = xb + (long long)c0 * ldn;
      const float* ysrc = yb + (long long)c0 * ldm + j0;
#pragma unroll
      for (int n = 0; n < NXS; ++n) {
        if (xs_off[n] < 0) continue;
        const bool ok = c0 + xkc[n] < C;
        cp_async16(xs + xs_off[n], ok ? (const void*)(xsrc + xg_off[n]) : (const void*)xb, ok);
      }
#pragma unroll
      for (int n = 0; n < NYS; ++n) {
        if (ys_off[n] < 0) continue;
        const bool ok = (c0 + ykc[n] < C) && (j0 + (ys_off[n] - ykc[n] * BN) < ldm);
        cp_async16(ys + ys_off[n], ok ? (const void*)(ysrc + yg_off[n]) : (const void*)yb, ok);
      }
    };

    // three-stage ring, one barrier per slab: the barrier of slab s also proves that every thread is done with slab s - 1,
    // whose buffer the load issued right after it (slab s + 2) overwrites
    load_slab(0);
    cp_async_commit();
    if (nslab > 1) load_slab(1);
    cp_async_commit();
    for (int s = 0; s < nslab; ++s) {
      cp_async_wait<1>();
      __syncthreads();
      if (s + 2 < nslab) load_slab(s + 2);
      cp_async_commit();
      const float* xs = Xs + (s % ST) * KC * BM;
      const float* ys = Ys + (s % ST) * KC * BN;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        float a[TM];
        unsigned long long b2[TN / 2];
        if constexpr (TM >= 4) {
#pragma unroll
          for (int g = 0; g < TM / 4; ++g) {
            const float4 v = *reinterpret_cast<const float4*>(xs + kc * BM + g * 4 * TY + ty * 4);
            a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < TM; ++i) a[i] = xs[kc * BM + ty * TM + i];
        }
#pragma unroll
        for (int g = 0; g < TN / 4; ++g) {
          const float4 v = *reinterpret_cast<const float4*>(ys + kc * BN + g * 4 * TX + tx * 4);
          b2[g * 2 + 0] = pack_f32x2(v.x, v.y);
          b2[g * 2 + 1] = pack_f32x2(v.z, v.w);
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const unsigned long long aa = pack_f32x2(a[i], a[i]);
#pragma unroll
          for (int j = 0; j < TN / 2; ++j) ffma2(acc2[i][j], aa, b2[j]);
        }
      }
    }
    cp_async_wait<0>();
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN / 2; ++j) unpack_f32x2(acc2[i][j], acc[i][2 * j], acc[i][2 * j + 1]);

    // epilogue: distances -> shared tile (|y|^2 and the relative-position bias as 16-byte loads when M allows)
    const bool vec4 = (M & 3) == 0;
    float sy[TN];
#pragma unroll
    for (int g = 0; g < TN / 4; ++g) {
      const int gj = j0 + g * 4 * TX + tx * 4;
      if (vec4 && gj < M) {
        const float4 v = *reinterpret_cast<const float4*>(sqy + (long long)b * M + gj);
        sy[g * 4 + 0] = v.x; sy[g * 4 + 1] = v.y; sy[g * 4 + 2] = v.z; sy[g * 4 + 3] = v.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) sy[g * 4 + e] = (gj + e < M) ? sqy[(long long)b * M + gj + e] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int row = TM >= 4 ? (i / 4) * 4 * TY + ty * 4 + (i % 4) : ty * TM + i;
      const int gi = i0 + row;
      const float sx = (gi < N) ? sqx[(long long)b * N + gi] : 0.f;
      const float* rprow = relpos ? relpos + (long long)gi * M : nullptr;
#pragma unroll
      for (int g = 0; g < TN / 4; ++g) {
        const int col = g * 4 * TX + tx * 4;
        const int gj0 = j0 + col;
        float rp[4] = {0.f, 0.f, 0.f, 0.f};
        if (rprow && gi < N) {
          if (vec4 && gj0 < M) {
            const float4 v = *reinterpret_cast<const float4*>(rprow + gj0);
            rp[0] = v.x; rp[1] = v.y; rp[2] = v.z; rp[3] = v.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (gj0 + e < M) rp[e] = rprow[gj0 + e];
          }
        }
        float dv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float d = __int_as_float(0x7f800000);
          if (gi < N && gj0 + e < M) {
            d = __fadd_rn(__fadd_rn(sx, __fmul_rn(-2.f, acc[i][g * 4 + e])), sy[g * 4 + e]);
            if (relpos) d = __fadd_rn(d, rp[e]);
          }
          dv[e] = d;
        }
        *reinterpret_cast<float4*>(Ds + row * DS + col) = make_float4(dv[0], dv[1], dv[2], dv[3]);
      }
    }
    __syncthreads();

    // Running top-K of every row (sorted by the strict total order (distance, index), hence unique), in three steps that keep
    // all lanes busy (the warp-per-row sorted-list insertion this replaces took 56 % of the kernel time):
    // (1) TPR threads per row, each over a segment of the chunk (thread 0 of the row also over the running list): the K
    //     smallest distance VALUES in a sorted register list (min / max insertion; slots below LK - K are -inf sentinels);
    //     the lists of a row are combined pairwise through shared memory -> the row threshold T = K-th smallest value.
    //     A thread walks its segment rotated by (row / 8) % 4 columns: DS = 172 maps 8 consecutive rows to banks 4 apart.
    const bool first = (j0 == jbeg);
    const float INF = __int_as_float(0x7f800000);
    constexpr int NFULL = NT / 32;
    // a CTA that walks MANY chunks (one global graph of 21 952 tokens = 172 chunks): after the first few almost no row changes
    // any more, and the warp-per-row path (one ballot per 32 candidates, nothing else when none beats the K-th best) is cheaper
    // than the block-wide steps below
    const bool warp_path = (j0 - jbeg) >= 8 * BN;
    if (warp_path) {
      if (warp < NFULL) {
        for (int row = warp; row < BM; row += NFULL) {
          if (i0 + row >= N) continue;
          float td = Ld[row * 32 + lane];
          int ti = Li[row * 32 + lane];
          warp_row_topk<BN, DS>(Ds, row, j0, M, K, lane, td, ti);
          Ld[row * 32 + lane] = td;
          Li[row * 32 + lane] = ti;
        }
      }
    } else {
    {
      constexpr int TPR = (NT / BM >= 4) ? 4 : (NT / BM >= 2 ? 2 : 1);
      constexpr int SEG = (BN + TPR - 1) / TPR;
      static_assert(2 * 32 <= 2 * KNN_CAP, "the published lists live in the candidate arrays");
      float* Ls = Cd;                      // [2][32][BM] published lists (Cd and Ci are contiguous; row-minor: no bank conflicts)
      const bool active = tid < BM * TPR;
      const int row = tid % BM, part = tid / BM;
      float lst[LK];
#pragma unroll
      for (int p = 0; p < LK; ++p) lst[p] = (p < LK - K) ? -INF : INF;
      // sorted insert without a serial chain: slot p takes min(old[p], max(old[p - 1], v)) — every slot depends only on
      // the OLD list, so the LK updates are independent (the bubble form v -> min / max -> next slot is LK deep)
      auto push = [&](float v) {
        if (v < lst[LK - 1]) {
#pragma unroll
          for (int p = LK - 1; p > 0; --p) lst[p] = fminf(lst[p], fmaxf(lst[p - 1], v));
          lst[0] = fminf(lst[0], v);
        }
      };
      auto publish = [&](int slot) {
#pragma unroll
        for (int p = 0; p < LK; ++p) Ls[(slot * 32 + p) * BM + row] = lst[p];
      };
      auto absorb = [&](int slot) {
        for (int p = LK - K; p < LK; ++p) {       // ascending: stop at the first value that cannot enter
          const float v = Ls[(slot * 32 + p) * BM + row];
          if (!(v < lst[LK - 1])) break;
          push(v);
        }
      };
      if (active) {
        // later chunks: the running list (sorted, K entries) seeds thread 0's list directly; the other threads of the row only
        // look at values below the running K-th best `cap` (a value >= cap cannot lower the K-th smallest value, and ties with
        // it are collected by step (2) from the distance tile itself)
        float cap = INF;
        if (!first) {
          cap = Ld[row * 32 + K - 1];
          if (part == 0) {
#pragma unroll
            for (int p = 0; p < LK; ++p) lst[p] = (p >= LK - K) ? Ld[row * 32 + p - (LK - K)] : -INF;
          }
        }
        const int c0 = part * SEG;
        const int len = min(SEG, BN - c0);
        const int skew = (row >> 3) & 3;
        const float* drow = Ds + row * DS + c0;
#pragma unroll 4
        for (int c = 0; c < len; ++c) {
          int col = c + skew;
          col -= (col >= len) ? len : 0;
          const float v = drow[col];
          if (v < cap) push(v);
        }
      }
      if constexpr (TPR == 4) {
        if (active && (part & 1)) publish(part >> 1);
        __syncthreads();
        if (active && !(part & 1)) absorb(part >> 1);
        __syncthreads();
      }
      if constexpr (TPR >= 2) {
        if (active && part == TPR / 2) publish(0);
        __syncthreads();
        if (active && part == 0) absorb(0);
      }
      if (active && part == 0) {
        Thr[row] = lst[LK - 1];
        Cnt[row] = 0;
        New[row] = 0;
      }
    }
    __syncthreads();
    // (2) all threads: the entries <= T (the K best and whatever ties with the K-th) -> the row's candidate array
    for (int e = tid; e < BM * BN; e += NT) {
      const int row = e / BN, col = e - row * BN;
      const float d = Ds[row * DS + col];
      if (d <= Thr[row] && j0 + col < M && i0 + row < N) {
        New[row] = 1;
        const int pos = atomicAdd(&Cnt[row], 1);
        if (pos < KNN_CAP) {
          Cd[row * KNN_CAP + pos] = d;
          Ci[row * KNN_CAP + pos] = j0 + col;
        }
      }
    }
    if (!first) {
      __syncthreads();                       // New[] is complete
      for (int e = tid; e < BM * 32; e += NT) {
        const int row = e >> 5;
        if (New[row] == 0) continue;         // nothing in this chunk reaches the row's list
        const float d = Ld[e];
        const int ci = Li[e];
        if ((e & 31) < K && ci != INT_MAX && d <= Thr[row]) {
          const int pos = atomicAdd(&Cnt[row], 1);
          if (pos < KNN_CAP) {
            Cd[row * KNN_CAP + pos] = d;
            Ci[row * KNN_CAP + pos] = ci;
          }
        }
      }
    }
    __syncthreads();
    // (3) rank of every candidate among its row's candidates = its slot in the new running list
    for (int e = tid; e < BM * 32; e += NT) {
      const int row = e >> 5;
      const int cnt = Cnt[row];
      if (cnt > KNN_CAP || New[row] == 0) continue;   // slow path below / list unchanged
      const float* cd = Cd + row * KNN_CAP;
      const int* cx = Ci + row * KNN_CAP;
      for (int a = e & 31; a < cnt; a += 32) {
        const float d = cd[a];
        const int ci = cx[a];
        int rank = 0;
        for (int bq = 0; bq < cnt; ++bq) rank += cand_less(cd[bq], cx[bq], d, ci) ? 1 : 0;
        if (rank < K) {
          Ld[row * 32 + rank] = d;
          Li[row * 32 + rank] = ci;
        }
      }
    }
    // rows with more than KNN_CAP candidates at or below the threshold (massive distance ties): one full warp per row
    if (warp < NFULL) {
      for (int row = warp; row < BM; row += NFULL) {
        if (Cnt[row] <= KNN_CAP || New[row] == 0) continue;
        float td = first ? INF : Ld[row * 32 + lane];
        int ti = first ? INT_MAX : Li[row * 32 + lane];
        warp_row_topk<BN, DS>(Ds, row, j0, M, K, lane, td, ti);
        Ld[row * 32 + lane] = td;
        Li[row * 32 + lane] = ti;
      }
    }
    }   // !warp_path
    __syncthreads();
    if (j0 + BN >= jend) {
      for (int e = tid; e < BM * 32; e += NT) {
        const int row = e >> 5, q = e & 31;
        const int gi = i0 + row;
        if (gi >= N) continue;
        if (split) {
          const long long o = (((long long)b * N + gi) * gridDim.z + blockIdx.z) * 32 + q;
          part_d[o] = q < K ? Ld[e] : INF;
          part_i[o] = q < K ? Li[e] : INT_MAX;
        } else if (q < K && (q % dilation) == 0) {
          const long long o = ((long long)b * N + gi) * k + q / dilation;
          out[o] = Li[e];
          if (out32) out32[o] = Li[e];
        }
      }
    }
  }
}

// merges the `parts` sorted partial lists of each row (one warp per row) and writes the k*dilation best
__global__ void __launch_bounds__(256) knn_merge_kernel(const float* __restrict__ part_d, const int* __restrict__ part_i,
                                                        long long rows, int parts, int k, int dilation,
                                                        int64_t* __restrict__ out, int32_t* __restrict__ out32) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long base = row * parts * 32;
  float td = part_d[base + lane];
  int ti = part_i[base + lane];
  for (int z = 1; z < parts; ++z) {
    const float rd = part_d[base + z * 32 + 31 - lane];   // the other list, reversed -> bitonic sequence of 64
    const int ri = part_i[base + z * 32 + 31 - lane];
    if (cand_less(rd, ri, td, ti)) {
      td = rd;
      ti = ri;
    }
    bitonic_merge32(td, ti, lane);
  }
  if (lane < k * dilation && (lane % dilation) == 0) {
    const long long o = row * k + lane / dilation;
    out[o] = ti;
    if (out32) out32[o] = ti;
  }
}

template <int BM, int BN, int TM, int TN>
static int launch_topk(const float* xn, const float* sqx, int ldn, const float* yn, const float* sqy, int ldm,
                       const float* relpos, int B, int N, int M, int C, int k, int dilation, int64_t* out,
                       int32_t* out32, cudaStream_t st, int msplit = 1, void* workspace = nullptr) {
  using Cfg = KnnCfg<BM, BN, TM, TN>;
  const int K = k * dilation;
  auto kern = K <= 8 ? knn_topk_kernel<BM, BN, TM, TN, 8> : K <= 16 ? knn_topk_kernel<BM, BN, TM, TN, 16>
              : K <= 24 ? knn_topk_kernel<BM, BN, TM, TN, 24> : knn_topk_kernel<BM, BN, TM, TN, 32>;
  int rc = ensure_smem(kern, Cfg::smem_bytes);
  if (rc) return rc;
  dim3 grid((N + BM - 1) / BM, B, msplit);
  const long long rows = (long long)B * N;
  float* part_d = reinterpret_cast<float*>(workspace);
  int* part_i = reinterpret_cast<int*>(part_d + (msplit > 1 ? rows * msplit * 32 : 0));
  kern<<<grid, Cfg::NT, Cfg::smem_bytes, st>>>(xn, sqx, ldn, yn, sqy, ldm, relpos, N, M, C, k, dilation, out, out32, part_d,
                                               part_i);
  rc = check_launch("knn_topk_kernel");
  if (rc || msplit == 1) return rc;
  knn_merge_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(part_d, part_i, rows, msplit, k, dilation, out, out32);
  return check_launch("knn_merge_kernel");
}

// candidate-split plan of the big cross-graph sites: 128-row tiles, one or more 168-wide chunks per CTA
static int knn_msplit(int B, int N, int M) {
  if (M % 168 != 0 || N <= 168 || M < 2 * 168) return 1;
  const long long row_tiles = (long long)((N + 127) / 128) * B;
  const int nchunks = M / 168;
  long long ms = (4LL * num_sms() + row_tiles - 1) / row_tiles;
  if (ms > nchunks) ms = nchunks;
  if (ms < 1) ms = 1;
  const long long cpz = (nchunks + ms - 1) / ms;   // chunks per CTA; drop the CTAs that would be left without a chunk
  return (int)((nchunks + cpz - 1) / cpz);
}

}  // namespace nextou

using namespace nextou;

extern "C" int nextou_knn_normalize(const void* x, int x_dtype, long long ldx, long long x_batch_stride,
                                    const int32_t* row_map, int B, int N, int C, int normalize, float* xn, int ldn,
                                    float* sq, void* stream) {
  NEXTOU_REQUIRE(x && xn && sq, "knn_normalize: null pointer");
  NEXTOU_REQUIRE(B > 0 && N > 0 && C > 0, "knn_normalize: bad shape B=%d N=%d C=%d", B, N, C);
  NEXTOU_REQUIRE(B <= 65535, "knn_normalize: B=%d > 65535", B);
  NEXTOU_REQUIRE(ldn >= N && ldn % 4 == 0, "knn_normalize: ldn=%d must be >= N=%d and a multiple of 4", ldn, N);
  NEXTOU_REQUIRE(x_dtype == NEXTOU_F32 || x_dtype == NEXTOU_BF16, "knn_normalize: bad dtype %d", x_dtype);
  const size_t smem = sizeof(float) * 32 * (size_t)(C | 1);
  NEXTOU_REQUIRE(smem <= 200 * 1024, "knn_normalize: C=%d too large", C);
  int rc = ensure_smem(knn_normalize_kernel, smem);
  if (rc) return rc;
  dim3 grid((ldn + 31) / 32, B);
  knn_normalize_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, x_dtype, ldx, x_batch_stride, row_map, N, C, normalize,
                                                                 xn, ldn, sq);
  return check_launch("knn_normalize_kernel");
}

extern "C" size_t nextou_knn_topk_workspace_bytes(int B, int N, int M) {
  const int ms = knn_msplit(B, N, M);
  return ms > 1 ? (size_t)B * N * ms * 32 * 8 : 0;
}

extern "C" int nextou_knn_topk(const float* xn, const float* sqx, int ldn, const float* yn, const float* sqy, int ldm,
                               const float* relpos, int B, int N, int M, int C, int k, int dilation, int64_t* out_idx,
                               int32_t* out_idx32, void* stream) {
  return nextou_knn_topk_ws(xn, sqx, ldn, yn, sqy, ldm, relpos, B, N, M, C, k, dilation, out_idx, out_idx32, nullptr, 0, stream);
}

extern "C" int nextou_knn_topk_ws(const float* xn, const float* sqx, int ldn, const float* yn, const float* sqy, int ldm,
                                  const float* relpos, int B, int N, int M, int C, int k, int dilation, int64_t* out_idx,
                                  int32_t* out_idx32, void* workspace, size_t workspace_bytes, void* stream) {
  NEXTOU_REQUIRE(xn && sqx && yn && sqy && out_idx, "knn_topk: null pointer");
  NEXTOU_REQUIRE(B > 0 && N > 0 && M > 0 && C > 0, "knn_topk: bad shape B=%d N=%d M=%d C=%d", B, N, M, C);
  NEXTOU_REQUIRE(B <= 65535, "knn_topk: B=%d > 65535", B);
  NEXTOU_REQUIRE(k >= 1 && dilation >= 1 && k * dilation <= 32, "knn_topk: k*dilation=%d outside [1,32]", k * dilation);
  NEXTOU_REQUIRE(k * dilation <= M, "knn_topk: k*dilation=%d > M=%d (topk would fail, TE:87)", k * dilation, M);
  NEXTOU_REQUIRE(ldn >= N && ldn % 4 == 0 && ldm >= M && ldm % 4 == 0, "knn_topk: ldn/ldm must be padded to 4");
  cudaStream_t st = (cudaStream_t)stream;
  if (M % 168 == 0) {   // every NexToU site: windows of 4x7x6 = 168 tokens and their multiples -> exact-fit candidate chunks
    if (N <= 168)   // windows: three 56-row tiles per window (measured faster than one 168 x 168 CTA with 8 x 8 register tiles)
      return launch_topk<56, 168, 4, 8>(xn, sqx, ldn, yn, sqy, ldm, relpos, B, N, M, C, k, dilation, out_idx, out_idx32, st);
    // cross-graph sites with many candidates: the per-(i, j) channel chain must stay sequential for bit-exactness, so the
    // parallelism comes from splitting the CANDIDATES over CTAs (sorted partial lists, merged by knn_merge_kernel)
    const int ms = knn_msplit(B, N, M);
    if (ms > 1 && workspace != nullptr && workspace_bytes >= (size_t)B * N * ms * 32 * 8)
      return launch_topk<128, 168, 8, 8>(xn, sqx, ldn, yn, sqy, ldm, relpos, B, N, M, C, k, dilation, out_idx, out_idx32, st, ms,
                                         workspace);
    return launch_topk<64, 168, 4, 8>(xn, sqx, ldn, yn, sqy, ldm, relpos, B, N, M, C, k, dilation, out_idx, out_idx32, st);
  }
  const long long tiles128 = (long long)((N + 127) / 128) * B;
  if (tiles128 >= 2LL * num_sms())
    return launch_topk<128, 128, 8, 8>(xn, sqx, ldn, yn, sqy, ldm, relpos, B, N, M, C, k, dilation, out_idx, out_idx32, st);
  return launch_topk<64, 128, 4, 8>(xn, sqx, ldn, yn, sqy, ldm, relpos, B, N, M, C, k, dilation, out_idx, out_idx32, st);
}
