/*
 * TEST INFRASTRUCTURE — CPU restatement (plain C) of the integer/index-exact parts of the NexToU
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * Product code (nextou_b200/) never links or calls it.
 *
 * Follows the reference (paths relative to /root/reference):
 *   oracle_knn_normalize   F.normalize(x, p=2, dim=1)                   network_architecture/torch_edge.py:154-160
 *   oracle_knn_topk        pairwise distance + relative_pos + topk       torch_edge.py:12-55, 58-110, 133
 *   oracle_bti_critical    binary_topological_interaction_module         loss/bti_loss.py:76-117
 *                          (and topological_interaction_module            loss/ti_loss.py:76-117)
 *
 * The floating-point summation ORDER is part of this file's contract (the CUDA kernels reproduce it
 * bit for bit); the reference itself leaves it to MKL/ATen, so agreement with the real reference is
 * checked tie-aware in tests/test_oracle_pin.py, while CUDA-vs-oracle is checked bit-exact.
 *   - channel sums: 32 "lanes", lane l accumulates c = l, l+32, ... with fmaf, then a 16/8/4/2/1
 *     xor-butterfly of adds;
 *   - dot products: acc = fmaf(x[c], y[c], acc), c ascending, single accumulator;
 *   - dist = ((sqx + (-2*acc)) + sqy) + relpos  (torch_edge.py:21-23 / 52-55, then :86 / :107);
 *   - neighbour order: ascending (dist, index) — ties to the lowest index (torch.topk leaves ties
 *     implementation-defined).
 *
 * Build: see oracle/c/Makefile (gcc -O2 -ffp-contract=off -mavx2 -mfma -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float lane_butterfly_sum_sq(const float* v, int C, int stride) {
  float s[32], t[32];
  for (int l = 0; l < 32; ++l) {
    float a = 0.f;
    for (int c = l; c < C; c += 32) a = fmaf(v[(long)c * stride], v[(long)c * stride], a);
    s[l] = a;
  }
  for (int off = 16; off > 0; off >>= 1) {
    for (int l = 0; l < 32; ++l) t[l] = s[l] + s[l ^ off];
    memcpy(s, t, sizeof(s));
  }
  return s[0];
}

/* x: token-major [B][N][C] fp32.  xn: channel-major [B][C][N].  sq: [B][N]. */
void oracle_knn_normalize(const float* x, int B, int N, int C, int normalize, float* xn, float* sq) {
#pragma omp parallel for schedule(static)
  for (long bn = 0; bn < (long)B * N; ++bn) {
    const int b = (int)(bn / N), n = (int)(bn % N);
    const float* row = x + bn * C;
    const float ss = lane_butterfly_sum_sq(row, C, 1);
    float denom = sqrtf(ss);
    if (!(denom > 1e-12f)) denom = 1e-12f; /* clamp_min(eps) */
    float* dst = xn + (long)b * C * N + n;
    for (int c = 0; c < C; ++c) dst[(long)c * N] = normalize ? row[c] / denom : row[c];
    sq[bn] = lane_butterfly_sum_sq(dst, C, N);
  }
}

/* xn: [B][C][N], yn: [B][C][M] (pass yn = xn, sqy = sqx, M = N for the self graph); relpos: [N][M] or NULL.
 * out: int64 [B][N][k]: the k*dilation nearest in ascending (dist, j) order, every dilation-th kept. */
int oracle_knn_topk(const float* xn, const float* sqx, const float* yn, const float* sqy, const float* relpos,
                    int B, int N, int M, int C, int k, int dilation, int64_t* out) {
  const int K = k * dilation;
  if (K > M || K < 1) return -1;
#pragma omp parallel
  {
    float* acc = (float*)malloc(sizeof(float) * M);
    float* bd = (float*)malloc(sizeof(float) * K);
    int* bi = (int*)malloc(sizeof(int) * K);
#pragma omp for schedule(dynamic, 16)
    for (long bn = 0; bn < (long)B * N; ++bn) {
      const int b = (int)(bn / N), i = (int)(bn % N);
      const float* xb = xn + (long)b * C * N;
      const float* yb = yn + (long)b * C * M;
      for (int j = 0; j < M; ++j) acc[j] = 0.f;
      for (int c = 0; c < C; ++c) {
        const float xv = xb[(long)c * N + i];
        const float* yr = yb + (long)c * M;
        for (int j = 0; j < M; ++j) acc[j] = fmaf(xv, yr[j], acc[j]);
      }
      const float sx = sqx[bn];
      int cnt = 0;
      for (int j = 0; j < M; ++j) {
        float d = (sx + (-2.f * acc[j])) + sqy[(long)b * M + j];
        if (relpos) d = d + relpos[(long)i * M + j];
        /* insert (d, j) into the sorted list; equal d keeps the earlier (lower) index first */
        if (cnt == K && !(d < bd[K - 1])) continue;
        int p = cnt < K ? cnt : K - 1;
        while (p > 0 && d < bd[p - 1]) {
          bd[p] = bd[p - 1];
          bi[p] = bi[p - 1];
          --p;
        }
        bd[p] = d;
        bi[p] = j;
        if (cnt < K) ++cnt;
      }
      for (int t = 0; t < k; ++t) out[bn * k + t] = bi[t * dilation];
    }
    free(acc);
    free(bd);
    free(bi);
  }
  return 0;
}

/* Critical-voxel map of the (binary) topological interaction module (bti_loss.py:76-117).
 *  labels : uint8 [B][D][H][W] argmax class per voxel (for 2-D pass D = 1 and dim = 2)
 *  maskA/maskC : per interaction, bit c set <=> class c belongs to set A / set C (bti_loss.py:90-98;
 *                TI_Loss uses singleton sets, ti_loss.py:89-98)
 *  inclusion[t] != 0 : C := not (C or A)   (bti_loss.py:91-95)
 *  connectivity : 26|6 (3-D) or 8|4 (2-D); min_thick only widens the box kernel (bti_loss.py:52-71)
 *  out : uint8 critical map, 1 where any interaction is violated (bti_loss.py:107-115) */
int oracle_bti_critical(const uint8_t* labels, int B, int D, int H, int W, int dim, const uint32_t* maskA,
                        const uint32_t* maskC, const uint8_t* inclusion, int n_inter, int connectivity,
                        int min_thick, uint8_t* out) {
  const int box = (dim == 3 && connectivity == 26) || (dim == 2 && connectivity == 8);
  const int cross = (dim == 3 && connectivity == 6) || (dim == 2 && connectivity == 4);
  if (!box && !cross) return -1;
  const int r = box ? min_thick : 1;
  const int rd = dim == 3 ? r : 0;
  const long V = (long)D * H * W;
  memset(out, 0, (size_t)B * V);
  if (n_inter < 0) return -2;   /* no upper bound: the reference loops over a python list (bti_loss.py:84) */
#pragma omp parallel for schedule(static)
  for (long bz = 0; bz < (long)B * D; ++bz) {
    const int b = (int)(bz / D), z = (int)(bz % D);
    const uint8_t* lab = labels + (long)b * V;
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const uint32_t me = 1u << lab[((long)z * H + y) * W + x];
        /* classes present in the neighbourhood (zero padding: outside contributes nothing; note that
         * for inclusion the complement mask is built BEFORE the conv, so padding is still 0) */
        uint32_t nb = 0;
        for (int dz = -rd; dz <= rd; ++dz)
          for (int dy = -r; dy <= r; ++dy)
            for (int dx = -r; dx <= r; ++dx) {
              if (cross && (abs(dz) + abs(dy) + abs(dx) > 1)) continue;
              const int zz = z + dz, yy = y + dy, xx = x + dx;
              if (zz < 0 || zz >= D || yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
              nb |= 1u << lab[((long)zz * H + yy) * W + xx];
            }
        int crit = 0;
        for (int t = 0; t < n_inter && !crit; ++t) {
          const uint32_t A = maskA[t];
          uint32_t Cm = maskC[t];
          if (inclusion[t]) Cm = ~(Cm | A);
          const int inA = (me & A) != 0, inC = (me & Cm) != 0;
          const int nearA = (nb & A) != 0, nearC = (nb & Cm) != 0;
          crit = (nearC && inA) || (nearA && inC);
        }
        out[(long)b * V + ((long)z * H + y) * W + x] = (uint8_t)crit;
      }
  }
  return 0;
}
