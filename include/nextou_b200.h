/*
 * nextou_b200 — C-ABI of the B200-native NexToU hot path (libnextou_b200.so).
 *
 * The reference (PengchengShi1220/NexToU) has no native code and no FFI: every GPU
 * instruction is issued by stock PyTorch ATen ops.  This header is therefore the NEW seam
 * under the reference's Python classes; each entry point names the reference call site
 * (file:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every function returns 0 on success or a negative nextou_status code; the message
 *     of the last failure on the calling thread is returned by nextou_last_error().
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller allocates
 *     every input, output and workspace (a *_workspace_bytes query exists where needed).
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised.
 *   - dtype codes: NEXTOU_F32 = 0, NEXTOU_BF16 = 1.
 *   - "token-major" = row per token / voxel, channels contiguous (NDHWC); `ld*` are row
 *     strides in ELEMENTS.
 */
#ifndef NEXTOU_B200_H
#define NEXTOU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEXTOU_ABI_VERSION 1

enum nextou_status {
  NEXTOU_OK = 0,
  NEXTOU_ERR_INVALID = -1,     /* bad argument (shape, alignment, unsupported k, ...) */
  NEXTOU_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed                */
  NEXTOU_ERR_UNSUPPORTED = -3, /* valid request this build cannot serve              */
  NEXTOU_ERR_WORKSPACE = -4    /* workspace too small                                */
};

enum nextou_dtype { NEXTOU_F32 = 0, NEXTOU_BF16 = 1 };

int nextou_abi_version(void);
const char* nextou_last_error(void);
/* number of kernels launched by this library on this process so far (bench.py's gpu_launches) */
long long nextou_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * kNN graph build.  Replaces DenseDilatedKnnGraph.forward (network_architecture/torch_edge.py:151-163):
 * F.normalize (TE:154-160) -> pairwise distance (TE:12-55) -> += relative_pos (TE:79,86,107)
 * -> topk(-dist, k*dilation) (TE:80,87,108) -> [..., ::dilation] (TE:133).
 * ------------------------------------------------------------------------------------------ */

/* Step 1: L2-normalise token rows (F.normalize p=2 dim=channels eps=1e-12, TE:154-160) and
 * transpose to channel-major fp32:  xn[b][c][n] = x[b][n][c] / max(||x[b][n]||, 1e-12),
 * sq[b][n] = sum_c xn^2.   x: token-major, dtype f32|bf16, row stride ldx, batch stride
 * x_batch_stride (elements).  If row_map != NULL (int32[B*N]) the source row of (b,n) is
 * row_map[b*N+n] (absolute row index into x; used to fold torch.roll + window_partition,
 * NexToU_Encoder_Decoder.py:634-660,784 into the load).  xn: [B][C][ldn], ldn >= N, ldn % 4 == 0.
 * normalize = 0 skips the division (xn = x, sq = sum_c x^2): dense_knn_matrix called on raw features (TE:58-90). */
int nextou_knn_normalize(const void* x, int x_dtype, long long ldx, long long x_batch_stride,
                         const int32_t* row_map, int B, int N, int C, int normalize, float* xn, int ldn,
                         float* sq, void* stream);

/* Step 2: fused distance tile + running top-k; the N x M distance matrix is never written.
 *   dist[b][i][j] = (sqx[b][i] + (-2 * dot(xn[b][:,i], yn[b][:,j]))) + sqy[b][j]  (+ relpos[i][j])
 * out_idx[b][i][0..k) = indices j of the (k*dilation) smallest dist in ascending (dist, j)
 * order, keeping every dilation-th (TE:133).  relpos is (N, M) fp32 row-major shared by all b,
 * or NULL.  For the self-graph (TE:58-90) pass yn = xn, sqy = sqx, M = N.
 * out_idx is int64 [B][N][k] (the dtype torch.topk returns); if out_idx32 != NULL an int32 copy
 * is written as well (consumed by nextou_mrconv_*).  Requires 1 <= k*dilation <= 32, <= M. */
int nextou_knn_topk(const float* xn, const float* sqx, int ldn, const float* yn, const float* sqy,
                    int ldm, const float* relpos, int B, int N, int M, int C, int k, int dilation,
                    int64_t* out_idx, int32_t* out_idx32, void* stream);
/* Same with a caller-allocated workspace: when nextou_knn_topk_workspace_bytes(B, N, M) > 0 the candidates of the big
 * cross-graph sites (Pool-GNN, N = 10752 / 1344 queries x M = 1344 candidates) are split over CTAs, each leaving a sorted
 * partial list per query row in the workspace, merged by a second kernel (identical result: the order (distance, index)
 * is total).  workspace == NULL falls back to the unsplit kernel. */
size_t nextou_knn_topk_workspace_bytes(int B, int N, int M);
int nextou_knn_topk_ws(const float* xn, const float* sqx, int ldn, const float* yn, const float* sqy, int ldm,
                       const float* relpos, int B, int N, int M, int C, int k, int dilation, int64_t* out_idx,
                       int32_t* out_idx32, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Max-relative message passing.  Replaces MRConv.forward lines 401-409
 * (network_architecture/NexToU_Encoder_Decoder.py) and batched_index_select (torch_nn.py:94-115):
 *   out[row][2c] = x[row][c];  out[row][2c+1] = max_j ( y[nbr(row,j)][c] - x[row][c] )
 * x, y, out token-major (dtype f32|bf16, same for all three).  idx: int32 [R][k], candidate index LOCAL
 * to its graph (graph g = r / N owns candidates g*M .. g*M+M-1).  q_row_map / y_row_map (int32 or NULL)
 * translate "graph-major" query / candidate numbers to physical rows (shifted windows, ED:634-660,784).
 * arg: uint8 [rows][C] neighbour slot attaining the max (first on ties), consumed by the backward.
 * ------------------------------------------------------------------------------------------ */
int nextou_mrconv_gather_fwd(const void* x, long long ldx, const void* y, long long ldy, int dtype, int C,
                             const int32_t* idx, int k, const int32_t* q_row_map, const int32_t* y_row_map,
                             long long R, int N, int M, void* out, long long ldo, uint8_t* arg, void* stream);
/* dx[row][c] += dout[row][2c] - dout[row][2c+1];  dy[nbr(row,arg)][c] += dout[row][2c+1].
 * dx / dy are fp32 accumulation buffers the caller zero-fills (dy may alias dx for the self graph). */
int nextou_mrconv_gather_bwd(const void* dout, long long ldo, int dtype, int C, const int32_t* idx, int k,
                             const uint8_t* arg, const int32_t* q_row_map, const int32_t* y_row_map, long long R,
                             int N, int M, float* dx, long long lddx, float* dy, long long lddy, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pool-GNN down / up-sampling on token-major volumes [B*D*H*W][C] (2-D: D = 1, pd = 1).  Replaces
 * nn.MaxPool3d(return_indices) / nn.MaxUnpool3d / F.avg_pool3d in PoolDyGraphConv.forward
 * (NexToU_Encoder_Decoder.py:511-512, 524-528, 536-549).  Pools are non-overlapping; `arg` holds the
 * uint8 child index (dz*ph*pw + dy*pw + dx) of the maximum instead of torch's int64 flat index.
 * Unpool: output channel j takes the arg of channel j % Carg (indices_cat = cat(indices, indices), ED:536).
 * ------------------------------------------------------------------------------------------ */
int nextou_maxpool3d_fwd(const void* x, int dtype, long long ldx, int C, int B, int D, int H, int W, int pd, int ph,
                         int pw, void* out, long long ldo, uint8_t* arg, void* stream);
int nextou_maxpool3d_bwd(const void* dout, int dtype, long long ldo, const uint8_t* arg, int C, int B, int D, int H,
                         int W, int pd, int ph, int pw, void* dx, long long ldx, void* stream);
int nextou_avgpool3d_fwd(const void* x, int dtype, long long ldx, int C, int B, int D, int H, int W, int pd, int ph,
                         int pw, void* out, long long ldo, void* stream);
int nextou_avgpool3d_bwd(const void* dout, int dtype, long long ldo, int C, int B, int D, int H, int W, int pd, int ph,
                         int pw, void* dx, long long ldx, void* stream);
int nextou_maxunpool3d_fwd(const void* g, int dtype, long long ldg, const uint8_t* arg, int Carg, int C2, int B, int D,
                           int H, int W, int pd, int ph, int pw, void* out, long long ldo, void* stream);
int nextou_maxunpool3d_bwd(const void* dout, int dtype, long long ldo, const uint8_t* arg, int Carg, int C2, int B,
                           int D, int H, int W, int pd, int ph, int pw, void* dg, long long ldg, void* stream);

/* ------------------------------------------------------------------------------------------
 * (Binary) topological interaction loss.  Replaces BTI_Loss.forward (loss/bti_loss.py:120-145),
 * binary_topological_interaction_module (bti_loss.py:76-117) and the TI twins (loss/ti_loss.py).
 * Logits are addressed as logits[b*stride_b + c*stride_c + v*stride_v] (NCDHW-contiguous and
 * channels-last both fit); target dtype codes: 0 f32, 1 bf16, 2 i64, 3 u8, 4 i32; <= 32 classes.
 * ------------------------------------------------------------------------------------------ */
/* labels[b][v] = argmax_c logits (first maximum; bti_loss.py:132-134); if ce != NULL also the per-voxel
 * fp64 cross entropy  logsumexp(logits) - logits[target]  (bti_loss.py:141; 0 for out-of-range targets). */
int nextou_bti_argmax_ce(const void* logits, int dtype, long long stride_b, long long stride_c, long long stride_v,
                         int B, int NC, long long V, const void* target, int target_code, uint8_t* labels, double* ce,
                         void* stream);
/* crit[b][z][y][x] = 1 iff any interaction t is violated:  (dilate(C_t) & A_t) | (dilate(A_t) & C_t), with
 * A_t / C_t the voxels whose label bit is in mask_a[t] / mask_c[t] (inclusion: C := ~(C|A), bti_loss.py:91-95),
 * dilation by the 3^d box of radius min_thick (connectivity 26 / 8) or the 6- / 4-cross, zero padded.
 * The three interaction arrays are HOST pointers (n_inter <= 32 entries). */
int nextou_bti_critical_map(const uint8_t* labels, int B, int D, int H, int W, int dim, const uint32_t* mask_a_host,
                            const uint32_t* mask_c_host, const uint8_t* inclusion_host, int n_inter, int connectivity,
                            int min_thick, uint8_t* crit, void* stream);
size_t nextou_bti_masked_sum_workspace_bytes(int B);
/* out[0] = mean_b sum_v ce[b][v] * crit[b][v]  (bti_loss.py:142-143), fp64, fixed summation order. */
int nextou_bti_masked_sum(const double* ce, const uint8_t* crit, int B, long long V, double* workspace, double* out,
                          void* stream);
/* dlogits = crit * grad_out[0] / B * (softmax(logits) - onehot(target)); grad_out is a DEVICE fp64 scalar. */
int nextou_bti_ce_bwd(const void* logits, int dtype, long long stride_b, long long stride_c, long long stride_v, int B,
                      int NC, long long V, const void* target, int target_code, const uint8_t* crit,
                      const double* grad_out, void* dlogits, long long dstride_b, long long dstride_c,
                      long long dstride_v, void* stream);

/* ---- fused per-scale training loss:  w_ce * CE + w_dice * SoftDice + w_ti * (B)TI -------------------------------
 * Replaces the three softmax evaluations and ~20 element-wise passes of DC_and_CE_and_BTI_Loss.forward
 * (loss/compound_bti_loss.py:33-61; Dice and CE are upstream nnU-Net's MemoryEfficientSoftDiceLoss /
 * RobustCrossEntropyLoss) with two passes over the logits.  NC in {2..8, 14, 16, 19}.
 * nextou_dsloss_plan: number of CTAs per batch item (= rows of `partial`).
 * nextou_dsloss_stats: labels[B][V] = argmax (first maximum), ce[B][V] = -log softmax[target] (fp64 storage, fp32 math),
 *   sums[B][3*NC+1] = sum_v p_c | sum_v p_c*y_c | sum_v y_c | sum_v ce, reduced in fixed order from
 *   partial[B][nblk][3*NC+1].
 * nextou_dsloss_bwd: dlogit_c = p_c (g_c - sum_k p_k g_k) + (scal[0] + scal[1]*crit) (p_c - y_c) with
 *   g_c = coef_a[b][c]*y_c + coef_b[b][c] (the Dice derivative); crit may be NULL. */
int nextou_dsloss_plan(long long V, int B, int* nblk_out);
int nextou_dsloss_stats(const void* logits, int dtype, long long stride_b, long long stride_c, long long stride_v, int B,
                        int NC, long long V, const void* target, int target_code, uint8_t* labels, double* ce,
                        double* partial, double* sums, void* stream);
int nextou_dsloss_bwd(const void* logits, int dtype, long long stride_b, long long stride_c, long long stride_v, int B,
                      int NC, long long V, const void* target, int target_code, const uint8_t* crit, const float* coef_a,
                      const float* coef_b, const float* scal, void* dlogits, long long dstride_b, long long dstride_c,
                      long long dstride_v, void* stream);
/* The per-class algebra between the two passes in one tiny launch each (instead of ~30 ATen kernels on [B, NC] doubles):
 * total = w_ce * mean CE - w_dice * mean Dice + w_ti * ti and the Dice derivative coefficients coef[2][B][NC] (fp64);
 * then, in backward, coef32 = coef * gout and the CE / TI scales nextou_dsloss_bwd takes. */
int nextou_dsloss_finish(const double* sums, const double* pooled, int B, int NC, long long V, double w_ce, double w_dice, double w_ti,
                         const double* ti, int batch_dice, int do_bg, double smooth, double grad_world, double* total,
                         double* coef, void* stream);
int nextou_dsloss_scale(const double* coef, const double* gout, int B, int NC, long long V, double w_ce, double w_ti, float* coef32,
                        float* scal, void* stream);

/* Same contract, restricted to kh, kw in {1, 3} (all non-down-sampling NexToU convolutions): halo-reuse variant
 * (csrc/conv_tcgen05.cu) — one haloed activation box per depth tap feeds all in-plane taps through row-shifted
 * UMMA descriptors, cutting the L2 -> SM activation traffic 6.4x. */
int nextou_conv3d_ndhwc_halo_fwd(const void* x, long long ldx, int B, int D, int H, int W, int Cin, const void* wpack,
                                 int Cout, int kd, int kh, int kw, const float* bias, void* out, long long ldo,
                                 int out_dtype, void* stream);
/* Weight gradient of the same convolutions (and, with kd = kh = kw = 1, of the 1x1 layers):
 *   dW[co][tap][ci] += sum_v dy[v][co] * x[v + tap - pad][ci]     (fp32, the caller zero-fills dW[Cout][taps][cin_stride])
 * dy / x: bf16 token-major volumes (pitches % 8 == 0).  Both operands are MN-major tcgen05 operands (the voxel axis is
 * K), the voxel axis is split over CTAs and reduced with fp32 atomics. */
int nextou_conv3d_ndhwc_wgrad(const void* dy, long long ldy, const void* x, long long ldx, int B, int D, int H, int W,
                              int Cin, int Cout, int kd, int kh, int kw, float* dW, int cin_stride, void* stream);

/* Same contract, kh, kw in {1, 3}: halo-reuse variant (one haloed X box per 8x8 voxel brick feeds all in-plane taps). */
int nextou_conv3d_ndhwc_halo_wgrad(const void* dy, long long ldy, const void* x, long long ldx, int B, int D, int H,
                                   int W, int Cin, int Cout, int kd, int kh, int kw, float* dW, int cin_stride,
                                   void* stream);

/* ------------------------------------------------------------------------------------------
 * Batch / instance normalisation (+ LeakyReLU) on a dense token-major matrix x[instances][rows][C]
 * (C = physical row pitch).  Replaces nn.BatchNorm{2,3}d in train mode (nnUNetTrainer_NexToU.py:54-55),
 * nn.InstanceNorm{2,3}d(affine) (torch_nn.py:42-46) and the LeakyReLU(0.01) behind them (torch_nn.py:16-17).
 * Batch norm = 1 instance spanning every token of the batch; instance norm = 1 instance per batch item.
 * slope = 1 disables the activation.  If instances > 1, rows*C must be a multiple of 4.
 * ------------------------------------------------------------------------------------------ */
/* number of CTA partial rows the stats / bwd kernels write: partial must hold instances*nblk*2*C floats */
int nextou_norm_plan(int C, long long rows, int instances, int* nblk_out);
/* mean / invstd [instances][C] (biased variance, eps inside the sqrt); running_* (may be NULL) are updated
 * in place with `momentum` and the UNBIASED variance, like nn.BatchNorm. */
int nextou_norm_stats(const void* x, int dtype, int C, long long rows, int instances, float eps, float* partial,
                      float* mean, float* invstd, float* running_mean, float* running_var, float momentum,
                      void* stream);
/* Same; additionally increments *num_batches_tracked (int64, device; may be NULL) like nn.BatchNorm does in training. */
int nextou_norm_stats_tracked(const void* x, int dtype, int C, int c_valid, long long rows, int instances, float eps,
                              float* partial, float* mean, float* invstd, float* running_mean, float* running_var,
                              float momentum, long long* num_batches_tracked, void* stream);
/* nextou_norm_apply with the per-channel vectors (gamma, beta) holding only c_valid <= C entries (C = physical pitch of the
 * channel-padded rows; the padding lanes use gamma = 1, beta = 0).  Likewise c_valid in _stats_tracked (running_*) and
 * _bwd_colsum: no padded copies of the module's parameters / buffers are needed. */
int nextou_norm_apply_cv(const void* x, int dtype, int C, int c_valid, long long rows, int instances, const float* mean,
                         const float* invstd, const float* gamma, const float* beta, float slope, void* y, void* stream);
/* Same with the block's residual shortcut fused in: y = lrelu(norm(x)) + residual (residual: same [instances][rows][C]
 * layout, may be NULL) — `x + BN(fc2(...))` of FFN / SwinGrapher / PoolGrapher (NexToU_Encoder_Decoder.py:389, 817, 932). */
int nextou_norm_apply_res(const void* x, int dtype, int C, int c_valid, long long rows, int instances, const float* mean,
                          const float* invstd, const float* gamma, const float* beta, float slope, const void* residual,
                          void* y, void* stream);
/* y = lrelu((x - mean) * invstd * gamma + beta, slope); gamma / beta [C] fp32 or NULL */
int nextou_norm_apply(const void* x, int dtype, int C, long long rows, int instances, const float* mean,
                      const float* invstd, const float* gamma, const float* beta, float slope, void* y,
                      void* stream);
/* backward through activation + normalisation: sums[inst][0][C] = sum dy' (= d beta), sums[inst][1][C] =
 * sum dy'*xhat (= d gamma), dx = gamma*invstd*(dy' - sums0/rows - xhat*sums1/rows), dy' = dy*lrelu'(pre). */
int nextou_norm_bwd(const void* x, const void* dy, int dtype, int C, long long rows, int instances,
                    const float* mean, const float* invstd, const float* gamma, const float* beta, float slope,
                    float* partial, float* sums, void* dx, void* stream);
/* Same, and additionally dx_colsum[inst][C] = sum_rows dx (as stored) when dx_colsum != NULL: the bias gradient of the
 * convolution / linear layer that feeds this normalisation, for free while dx is written (`partial` is reused). */
int nextou_norm_bwd_colsum(const void* x, const void* dy, int dtype, int C, int c_valid, long long rows, int instances,
                           const float* mean, const float* invstd, const float* gamma, const float* beta, float slope,
                           float* partial, float* sums, void* dx, float* dx_colsum, void* stream);
/* The two steps of nextou_norm_bwd_colsum, separately, for SyncBatchNorm (upstream nnU-Net converts every BatchNorm with
 * SyncBatchNorm.convert_sync_batchnorm under DDP): step 1 leaves this rank's sums, the caller all-reduces them over the
 * ranks, step 2 takes the reduced sums and the GLOBAL row count n_total. */
int nextou_norm_bwd_reduce(const void* x, const void* dy, int dtype, int C, int c_valid, long long rows, int instances,
                           const float* mean, const float* invstd, const float* gamma, const float* beta, float slope,
                           float* partial, float* sums, void* stream);
int nextou_norm_bwd_apply(const void* x, const void* dy, int dtype, int C, int c_valid, long long rows, int instances,
                          long long n_total, const float* mean, const float* invstd, const float* gamma, const float* beta,
                          float slope, const float* sums, float* partial, void* dx, float* dx_colsum, void* stream);
/* First halves of nextou_norm_stats / nextou_norm_bwd_reduce alone: this rank's per-CTA partial rows [*nblk_out][2C] (sized by
 * nextou_norm_plan), the input of the cross-rank finalize below. */
int nextou_norm_partial_stats(const void* x, int dtype, int C, long long rows, float* partial, int* nblk_out, void* stream);
int nextou_norm_bwd_partial(const void* x, const void* dy, int dtype, int C, int c_valid, long long rows, const float* mean,
                            const float* invstd, const float* gamma, const float* beta, float slope, float* partial,
                            int* nblk_out, void* stream);
/* SyncBatchNorm statistics over NVLink peer memory (csrc/syncnorm.cu).  Upstream nnU-Net converts every BatchNorm with
 * SyncBatchNorm.convert_sync_batchnorm under DDP: 78 layers x (forward + backward) = 156 latency-bound exchanges per step.  One
 * kernel per exchange: local fp64 column sums -> stores into the slot of every peer (mapped peer addresses) -> per-CTA flags ->
 * sum in rank order -> mean / invstd / running statistics (forward) or global sums (backward).  No host synchronisation, no
 * collective library call.  peer_base: DEVICE array [world] with the base address of every rank's symmetric buffer as mapped
 * on this rank; offset: byte offset of this call site's slot (nextou_sync_slot_bytes, same on all ranks, zeroed once); epoch:
 * local device array of nextou_sync_slot_ctas zeroed uint64 counters owned by the call site. */
int nextou_sync_slot_ctas(int C, int backward);
long long nextou_sync_slot_bytes(int world, int C, int backward);
int nextou_sync_norm_finalize(const float* partial, int nblk, int C, int c_valid, long long rows, float eps, void* const* peer_base,
                              long long offset, int rank, int world, unsigned long long* epoch, float* mean, float* invstd,
                              float* running_mean, float* running_var, float momentum, long long* num_batches_tracked,
                              double* n_total, void* stream);
int nextou_sync_norm_bwd_finalize(const float* partial, int nblk, int C, void* const* peer_base, long long offset, int rank, int world,
                                  unsigned long long* epoch, float* sums_local, float* sums_global, void* stream);
/* column sums of a dense [rows][C] matrix: sums[0][C] = sum_r x, sums[1][C] = sum_r x^2 (fp32); `partial` as for
 * nextou_norm_stats with instances = 1.  Bias gradients of the 1x1 / spatial convolutions (d bias = colsum(dY)). */
int nextou_colsum(const void* x, int dtype, int C, long long rows, float* partial, float* sums, void* stream);
/* eval-mode batch norm: y = lrelu(x*scale[c] + shift[c]) with the running statistics folded by the caller */
int nextou_affine_act(const void* x, int dtype, int C, long long rows, const float* scale, const float* shift,
                      float slope, void* y, void* stream);

/* ------------------------------------------------------------------------------------------
 * tcgen05 / TMEM / TMA GEMM engine (csrc/gemm_tcgen05.cu): bf16 operands, fp32 accumulation in tensor memory.
 * ------------------------------------------------------------------------------------------ */
/* C[M][ldc] = A[M][K] * B[N][K]^T (+ bias[N]).  The 1x1 convolutions of the graphers / FFNs / segmentation heads
 * (NexToU_Encoder_Decoder.py:305, 373-381, 710-720, 833-842) are exactly this with A = token rows, B = conv weight.
 * A, B: bf16, K contiguous, row pitches lda / ldb (elements, multiples of 8), 16-byte aligned.  C: bf16 or fp32,
 * ldc % 8 == 0; columns [N, ldc) are written as zeros (channel padding of the token-major format). */
int nextou_gemm_bf16_tn(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M,
                        int N, int K, const float* bias, int out_dtype, void* stream);
/* Stride-1 'same' convolution (odd kernel, zero padding (k-1)/2: the StackedConvBlocks convs, ED:125-141, 281-300)
 * as implicit GEMM.  x: bf16 NDHWC [B][D][H][W][ldx]; wpack: bf16 [Cout][taps*cin_pad], cin_pad = ceil(Cin/64)*64,
 * taps ordered (kd, kh, kw), zero padded; out: [B*D*H*W][ldo], columns [Cout, ldo) zero-filled.  2-D: D = kd = 1.
 * The data-gradient of such a convolution is the same call on dY with the flipped / transposed weight pack. */
int nextou_conv3d_ndhwc_fwd(const void* x, long long ldx, int B, int D, int H, int W, int Cin, const void* wpack,
                            int Cout, int kd, int kh, int kw, const float* bias, void* out, long long ldo,
                            int out_dtype, void* stream);


/* Strided convolution forward — the down-sampling first convolution of encoder stages 1..5 (StackedConvBlocks call sites
 * NexToU_Encoder_Decoder.py:125-141; conv(k, stride, pad=(k-1)//2)):  out[o] = bias + sum_k x[o*s + k - p] . W[k].
 * x: bf16 NDHWC [B][Di][Hi][Wi][ldx]; wpack as for nextou_conv3d_ndhwc_fwd; out: [B*Do*Ho*Wo][ldo] with
 * Do = (Di + 2 pd - kd) / sd + 1 (likewise H, W).  The strided gather is done by the TMA unit (element strides). */
int nextou_conv3d_ndhwc_strided_fwd(const void* x, long long ldx, int B, int Di, int Hi, int Wi, int Cin,
                                    const void* wpack, int Cout, int kd, int kh, int kw, int sd, int sh, int sw, int pd,
                                    int ph, int pw, const float* bias, void* out, long long ldo, int out_dtype,
                                    void* stream);
/* Data gradient of that convolution, and — with pd = ph = pw = 0 and kernel == stride — the FORWARD of the decoder's
 * transposed convolutions (ConvTranspose(k = s = stride), NexToU_Encoder_Decoder.py:273-276, 321):
 *   dx[u][ci] = bias[ci] + sum_{(v,k): v*s + k - p == u} dy[v][:] . wpack_t[ci][k][:]
 * dy: bf16 NDHWC [B][Do][Ho][Wo][ldy] (Cout channels); wpack_t: bf16 [Cin][taps*cout_pad], taps in (kd,kh,kw) order (not
 * flipped), cout_pad = ceil(Cout/64)*64; dx: [B*Di*Hi*Wi][ldx].  One launch per output parity class (u mod s). */
int nextou_conv3d_ndhwc_strided_dgrad(const void* dy, long long ldy, int B, int Do, int Ho, int Wo, int Cout,
                                      const void* wpack_t, int Cin, int kd, int kh, int kw, int sd, int sh, int sw, int pd,
                                      int ph, int pw, const float* bias, void* dx, long long ldx, int Di, int Hi, int Wi,
                                      int out_dtype, void* stream);
/* Same, writing only columns [0, store_cols) of every dx row (Cin <= store_cols <= ldx, store_cols % 8 == 0; columns
 * [Cin, store_cols) zero-filled): the up-sampled half of torch.cat((up, skip), 1) (NexToU_Encoder_Decoder.py:322) is
 * written straight into the concatenation buffer. */
int nextou_conv3d_ndhwc_strided_dgrad_cols(const void* dy, long long ldy, int B, int Do, int Ho, int Wo, int Cout,
                                           const void* wpack_t, int Cin, int kd, int kh, int kw, int sd, int sh, int sw,
                                           int pd, int ph, int pw, const float* bias, void* dx, long long ldx,
                                           int store_cols, int Di, int Hi, int Wi, int out_dtype, void* stream);
/* Weight gradient with a strided read:  dW[m][tap][n] += sum_i dy[i][m] * x[i*s + tap - pad][n]  (fp32, caller zero-fills
 * dW[Cout][taps][cin_stride]).  dy: bf16 tokens of the dense grid [B][D][H][W][ldy] (Cout = M channels); x: bf16 tokens of
 * the strided-read volume [B][Dx][Hx][Wx][ldx] (Cin = N channels).  Strided convolution: dy = output gradient, x = input.
 * Transposed convolution (kernel == stride, pad 0): dy := the layer input, x := the output gradient. */
int nextou_conv3d_ndhwc_strided_wgrad(const void* dy, long long ldy, const void* x, long long ldx, int B, int D, int H,
                                      int W, int Dx, int Hx, int Wx, int Cin, int Cout, int kd, int kh, int kw, int sd,
                                      int sh, int sw, int pd, int ph, int pw, float* dW, int cin_stride, void* stream);

/* out[r][0..cols) = a[r][0..cols) (+ b[r][0..cols), b may be NULL) on row views of different pitch / column offset (elements;
 * bases, pitches and cols multiples of 16 bytes: the channel-padded token layout).  The skip half of torch.cat((up, skip), 1)
 * (NexToU_Encoder_Decoder.py:322) and the gradient sum of a tensor consumed twice (skip connection, residual shortcut). */
int nextou_rows_copy_add(const void* a, long long lda, const void* b, long long ldb, void* out, long long ldo, long long rows,
                         int cols, int dtype, void* stream);

/* First convolution of the network (Cin = image modalities <= 4, Cout <= 64; NexToU_Encoder_Decoder.py:125-141 stage 0): K =
 * taps x Cin is far below one tensor-core slab, so these are streaming CUDA-core kernels.  x, out / dy: bf16 token-major;
 * w: the fp32 master weight (Cout, Cin, kd, kh, kw) (rounded to bf16 on load, like the operand packs of the tensor-core path);
 * dW[Cout][taps][cin_stride] fp32, zero-filled by the caller.  `supported` tells whether a layer is covered (taps x Cin <= 12). */
int nextou_conv3d_small_cin_supported(int Cin, int Cout, int kd, int kh, int kw, long long ldo);
int nextou_conv3d_small_cin_fwd(const void* x, long long ldx, int B, int D, int H, int W, int Cin, const float* w, int Cout,
                                int kd, int kh, int kw, const float* bias, void* out, long long ldo, void* stream);
int nextou_conv3d_small_cin_wgrad(const void* dy, long long ldy, const void* x, long long ldx, int B, int D, int H, int W,
                                  int Cin, int Cout, int kd, int kh, int kw, float* dW, int cin_stride, void* stream);

/* Weight gradient of a down-sampling convolution (3 x 3 in-plane kernel, in-plane stride 2, padding 1; depth kernel 1 | 3, depth
 * stride 1 | 2) with halo reuse over the four (h, w)-parity planes of the input (tensor maps with doubled strides over the same
 * memory: no copy).  Covers ceil16(Cin) * 9 <= 512 (the full-resolution encoder layer); `supported` tells. */
int nextou_conv3d_ndhwc_planes_wgrad_supported(int Cin, int kd, int kh, int kw, int sd, int sh, int sw, int pd, int ph, int pw);
int nextou_conv3d_ndhwc_planes_wgrad(const void* dy, long long ldy, const void* x, long long ldx, int B, int Do, int Ho, int Wo,
                                     int Di, int Hi, int Wi, int Cin, int Cout, int kd, int sd, int pd, float* dW, int cin_stride,
                                     void* stream);

/* Forward of a kernel == stride transposed convolution (decoder up-sampling, NexToU_Encoder_Decoder.py:273-276, 321) as ONE
 * persistent GEMM over the input voxels whose epilogue scatters the kd*kh*kw class segments of a row to their output voxels
 * (the input is read once).  wpack_t: the Bt pack of nextou_pack_weight on the (Cin, Cout, *k) weight, [Cout][taps][cin_pad64];
 * out rows: columns [0, store_cols) written.  `supported`: ceil16(store_cols) <= 256. */
int nextou_convtranspose_scatter_fwd_supported(int Cout, int store_cols, int kd, int kh, int kw);
int nextou_convtranspose_scatter_fwd(const void* x, long long ldx, int B, int D, int H, int W, int Cin, const void* wpack_t, int Cout,
                                     int kd, int kh, int kw, const float* bias, void* out, long long ldo, int store_cols,
                                     void* stream);

/* Weight packing (one launch per layer and step): master weight w[R][Cc/groups][taps] (fp32 | bf16; nn.Conv layout
 * (Cout, Cin/groups, *k) or nn.ConvTranspose layout (Cin, Cout, *k)) ->
 *   A [R][taps][lda_c]  bf16 = w[r][c][t]       (forward operand;       lda_c >= Cc, zero padded)
 *   Bt[Cc][taps][ldb_c] bf16 = w[r][c][flip(t)] (data-gradient operand; ldb_c >= R,  zero padded; may be NULL)
 * groups > 1 expands a grouped 1x1 convolution (torch_nn.py:85) to its block-diagonal dense operand. */
int nextou_pack_weight(const void* w, int dtype, int R, int Cc, int taps, int groups, int flip_b, void* A, int lda_c,
                       void* Bt, int ldb_c, void* stream);
/* Same with a zero channel gap [gap_lo, gap_hi) inserted into the INPUT-channel axis of both packs (lda_c >= Cc + gap width,
 * Bt has Cc + gap width rows): the first convolution of a decoder stage reads the concatenation buffer [up | gap | skip]
 * (NexToU_Encoder_Decoder.py:321-322) whose up-sampled half is padded to a multiple of 8 channels. */
int nextou_pack_weight_gap(const void* w, int dtype, int R, int Cc, int taps, int groups, int flip_b, int gap_lo, int gap_hi,
                           void* A, int lda_c, void* Bt, int ldb_c, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer side of the training step: upstream nnUNetTrainer.train_step runs
 * torch.nn.utils.clip_grad_norm_(network.parameters(), 12) and torch.optim.SGD(momentum 0.99, nesterov, weight_decay 3e-5)
 * .step() after backward; the next forward re-derives the bf16 operand packs of every weight.  Here: all parameter tensors
 * in three launches + all packs in a fourth.  The tables live in device memory and are built once by the caller.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  float* param;      /* fp32 master weights, updated in place */
  float* grad;       /* fp32 gradient (optionally overwritten with the clipped gradient) */
  float* momentum;   /* fp32 momentum buffer (zero-initialised by the caller) or NULL when momentum == 0 */
  long long numel;
} NextouOptTensor;
typedef struct {
  int tensor;        /* index into the tensor / pack-job table */
  int count;         /* elements of this chunk (<= nextou_opt_chunk_elems()) */
  long long start;   /* first element */
} NextouOptChunk;
typedef struct {
  const void* w;     /* master weight [R][Cc/groups][taps] */
  void* A;           /* bf16 [R][taps][lda_c] */
  void* Bt;          /* bf16 [Cc + gap][taps][ldb_c] or NULL */
  int R, Cc, taps, groups, flip_b, lda_c, ldb_c, gap_lo, gap_hi;
} NextouPackJob;
int nextou_opt_chunk_elems(void);
/* state[0] = 2-norm of all gradients, state[1] = min(1, max_norm / (state[0] + 1e-6)) (1 when max_norm <= 0);
 * partial: n_partial doubles of workspace.  No host synchronisation. */
int nextou_opt_grad_norm(const NextouOptTensor* tensors, const NextouOptChunk* chunks, int n_chunks, float max_norm,
                         double* partial, int n_partial, float* state, void* stream);
/* g' = state[1]*g + weight_decay*p;  buf = momentum*buf + (1-dampening)*g';  p -= lr[0] * (nesterov ? g' + momentum*buf : buf)
 * (torch.optim.SGD.step on clipped gradients; state == NULL: no clipping; lr is read on the device). */
int nextou_opt_sgd_step(const NextouOptTensor* tensors, const NextouOptChunk* chunks, int n_chunks, const float* lr,
                        const float* state, float momentum, float dampening, float weight_decay, int nesterov,
                        int write_clipped_grad, void* stream);
/* every pack job (fp32 master weights) in one launch; chunks index the concatenated A | Bt element range of a job */
int nextou_opt_pack_weights(const NextouPackJob* jobs, const NextouOptChunk* chunks, int n_chunks, void* stream);

/* Inference forms: out = lrelu(acc * scale[N] + shift[N], slope).  An eval-mode BatchNorm (+ the
 * LeakyReLU behind it) is folded into the epilogue of the layer that feeds it: scale = gamma * rsqrt(running_var + eps),
 * shift = beta + (conv bias - running_mean) * scale — no statistics pass, no normalisation pass, no collective.
 * scale == NULL means 1; slope == 1 disables the activation; the plain entry points are these with scale = NULL, slope = 1. */
int nextou_gemm_bf16_tn_affine(const void* A, long long lda, const void* B, long long ldb, void* C, long long ldc, int M,
                               int N, int K, const float* scale, const float* shift, float slope, int out_dtype,
                               void* stream);
int nextou_conv3d_ndhwc_halo_fwd_affine(const void* x, long long ldx, int B, int D, int H, int W, int Cin, const void* wpack,
                                        int Cout, int kd, int kh, int kw, const float* scale, const float* shift, float slope,
                                        void* out, long long ldo, int out_dtype, void* stream);
int nextou_conv3d_ndhwc_strided_fwd_affine(const void* x, long long ldx, int B, int Di, int Hi, int Wi, int Cin,
                                           const void* wpack, int Cout, int kd, int kh, int kw, int sd, int sh, int sw,
                                           int pd, int ph, int pw, const float* scale, const float* shift, float slope,
                                           void* out, long long ldo, int out_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEXTOU_B200_H */
