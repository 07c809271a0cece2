// Cross-rank batch statistics of SyncBatchNorm over NVLink peer memory.
//
// Upstream nnU-Net converts every BatchNorm with SyncBatchNorm.convert_sync_batchnorm before wrapping the network in DDP
// (SURVEY.md 8e): 78 layers x (forward + backward) = 156 exchanges of 2C+1 values per training step, all on the critical
// path.  As NCCL all-reduces (plus the ~10 small ATen kernels that build / unpack their buffers) they cost 6.4 ms of a
// 38.7 ms step at 2 GPUs (round-1 SCALE).  Here an exchange is ONE kernel: every rank reduces its per-CTA partial rows to
// per-channel fp64 sums, STORES them into its slot of every peer's buffer through the peers' mapped addresses (NVLink /
// NVSwitch peer memory, torch symmetric memory gives the pointers), raises a per-CTA flag on each peer, waits for the
// flags of all peers, adds the slots in rank order (every rank gets bit-identical statistics) and finishes the layer's
// statistics (mean / invstd / running statistics forward; global sums backward) in place.  Nothing but the device clock is
// involved: no host synchronisation, no collective library call; CUDA graphs capture it like any other kernel.
//
// Slot layout in every rank's symmetric buffer (doubles):  [2 epochs parity][world source ranks][cols]  then the flags
// [world source ranks][n_cta] (uint64).  A CTA owns 32 columns: it only waits for the matching CTA of every peer.
#include "common.cuh"

namespace nextou {

constexpr int SX_SLICES = 32;      // block = 32 columns x 32 row slices, like the single-GPU finalize

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

struct SyncSlot {
  void* const* peer_base;    // device array [world]: base address of every rank's symmetric buffer as mapped on THIS rank
  long long offset;          // byte offset of this call site's slot inside the buffers (same on every rank)
  int rank, world, cols, n_cta;
  unsigned long long* epoch; // local, [n_cta]: exchanges done so far by each CTA of this call site
};

__device__ __forceinline__ double* slot_data(const SyncSlot& s, int on_rank, int parity, int src) {
  return reinterpret_cast<double*>(reinterpret_cast<char*>(s.peer_base[on_rank]) + s.offset) + ((long long)parity * s.world + src) * s.cols;
}
__device__ __forceinline__ unsigned long long* slot_flag(const SyncSlot& s, int on_rank, int src, int cta) {
  double* end = reinterpret_cast<double*>(reinterpret_cast<char*>(s.peer_base[on_rank]) + s.offset) + 2LL * s.world * s.cols;
  return reinterpret_cast<unsigned long long*>(end) + (long long)src * s.n_cta + cta;
}

// One exchange of this CTA's 32 slot columns.  Lane l of slice 0 carries the local value `mine` (already reduced over the
// partial rows); returns the sum over all ranks of that column (valid in the 32 threads of slice 0).
__device__ __forceinline__ double exchange_columns(const SyncSlot& s, double mine) {
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) {
    s_epoch = s.epoch[blockIdx.x] + 1;
    s.epoch[blockIdx.x] = s_epoch;
  }
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  const int parity = (int)(epoch & 1);
  // push my 32 sums into my slot on every rank (mine included), then raise my flag there
  if (slice == 0)
    for (int r = 0; r < s.world; ++r) slot_data(s, r, parity, s.rank)[col] = mine;
  __syncthreads();
  // st.release.sys orders the 32 data stores above (made visible to this thread by the CTA barrier: release is cumulative)
  // before the flag, on every peer; no separate system-wide fence
  if (threadIdx.x < s.world) st_release_sys(slot_flag(s, threadIdx.x, s.rank, blockIdx.x), epoch);
  // wait for the matching CTA of every rank
  if (threadIdx.x < s.world) {
    const unsigned long long* f = slot_flag(s, s.rank, threadIdx.x, blockIdx.x);
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < epoch)
      if (clock64() - t0 > 20000000000LL) __trap();       // ~10 s: a rank that never arrives must not hang the GPU forever
  }
  __syncthreads();                                          // ld.acquire.sys above + the barrier order the slot reads below
  double g = 0.0;
  if (slice == 0)
    for (int r = 0; r < s.world; ++r) g += ld_volatile_f64(slot_data(s, s.rank, parity, r) + col);   // rank order: identical everywhere
  return g;
}

// fixed-order fp64 sum of column `pcol` (or nothing if pcol < 0) of the local partial rows [nblk][ld]; valid in slice 0
__device__ __forceinline__ double local_column_sum(const float* __restrict__ partial, int nblk, int ld, int pcol, double* sh) {
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  double acc = 0.0;
  if (pcol >= 0)
    for (int b = slice; b < nblk; b += SX_SLICES) acc += (double)partial[(long long)b * ld + pcol];
  sh[slice * 32 + lane] = acc;
  __syncthreads();
  double tot = 0.0;
  if (slice == 0)
    for (int i = 0; i < SX_SLICES; ++i) tot += sh[i * 32 + lane];
  return tot;
}

constexpr int SX_CH_PER_CTA = 15;   // forward: lanes (2j, 2j+1) = (sum x, sum x^2) of channel j, lane 31 = the row count

// forward: mean / invstd / running statistics over the rows of all ranks.  grid = ceil(C / 15)
__global__ void __launch_bounds__(32 * SX_SLICES)
    sync_norm_finalize_kernel(const float* __restrict__ partial, int nblk, int C, int c_valid, double rows, float eps, SyncSlot s,
                              float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                              float* __restrict__ running_var, float momentum, long long* __restrict__ tracked,
                              double* __restrict__ n_total) {
  __shared__ double sh[SX_SLICES * 32];
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * SX_CH_PER_CTA + (lane >> 1);
  const bool data = lane < 2 * SX_CH_PER_CTA && c < C;
  double mine = local_column_sum(partial, nblk, 2 * C, data ? (lane & 1) * C + c : -1, sh);
  if (lane == 31) mine = rows;
  const double g = exchange_columns(s, mine);
  if (threadIdx.x >= 32) return;
  const double n = __shfl_sync(0xffffffffu, g, 31);
  const double s2 = __shfl_down_sync(0xffffffffu, g, 1);
  if (data && (lane & 1) == 0) {
    const double m = g / n;
    double var = s2 / n - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr && c < c_valid) {
      const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
      running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * m);
      running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
    }
  }
  if (blockIdx.x == 0 && lane == 31) {
    if (n_total != nullptr) *n_total = n;
    if (tracked != nullptr) *tracked += 1;
  }
}

// backward: cols = 2C (sum dy', sum dy' xhat).  sums_local keeps this rank's sums (d beta, d gamma), sums_global the exchange
__global__ void __launch_bounds__(32 * SX_SLICES)
    sync_norm_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int cols, SyncSlot s, float* __restrict__ sums_local,
                                  float* __restrict__ sums_global) {
  __shared__ double sh[SX_SLICES * 32];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const double local = local_column_sum(partial, nblk, cols, col < cols ? col : -1, sh);
  const double g = exchange_columns(s, local);
  if (threadIdx.x < 32 && col < cols) {
    sums_local[col] = (float)local;
    sums_global[col] = (float)g;
  }
}

}  // namespace nextou

using namespace nextou;

// CTAs of an exchange: forward packs 15 channels (+ the row count) into 32 slot columns, backward 32 plain columns of 2C
extern "C" int nextou_sync_slot_ctas(int C, int backward) { return backward ? (2 * C + 31) / 32 : (C + SX_CH_PER_CTA - 1) / SX_CH_PER_CTA; }
// bytes one call site needs in every rank's symmetric buffer (data + flags), rounded to 128; zero-initialised once
extern "C" long long nextou_sync_slot_bytes(int world, int C, int backward) {
  const long long n_cta = nextou_sync_slot_ctas(C, backward);
  const long long b = 8LL * (2LL * world * n_cta * 32 + (long long)world * n_cta);
  return (b + 127) / 128 * 128;
}

static int make_slot(SyncSlot& s, void* const* peer_base, long long offset, int rank, int world, int n_cta, unsigned long long* epoch,
                     const char* who) {
  NEXTOU_REQUIRE(peer_base && epoch && world >= 1 && world <= 32 && rank >= 0 && rank < world && offset >= 0 && offset % 8 == 0,
                 "%s: bad exchange slot", who);
  s.peer_base = peer_base; s.offset = offset; s.rank = rank; s.world = world; s.cols = n_cta * 32; s.n_cta = n_cta; s.epoch = epoch;
  return 0;
}

// Forward statistics of a SyncBatchNorm layer from this rank's CTA partial rows [nblk][2C] (nextou_norm_partial_stats): mean /
// invstd over the rows of ALL ranks, running statistics (unbiased variance, global count), *n_total = global row count
// (device).  peer_base: device array [world] of the ranks' symmetric buffer bases as mapped on this rank; offset: this call
// site's slot (nextou_sync_slot_bytes(world, C, 0) bytes); epoch: local device array of nextou_sync_slot_ctas(C, 0) counters.
// Slot and counters are zeroed once by the caller and then owned by the call site.
extern "C" int nextou_sync_norm_finalize(const float* partial, int nblk, int C, int c_valid, long long rows, float eps,
                                         void* const* peer_base, long long offset, int rank, int world, unsigned long long* epoch,
                                         float* mean, float* invstd, float* running_mean, float* running_var, float momentum,
                                         long long* num_batches_tracked, double* n_total, void* stream) {
  NEXTOU_REQUIRE(partial && mean && invstd && nblk > 0 && C > 0 && rows > 0, "sync_norm_finalize: bad arguments");
  SyncSlot s;
  int rc = make_slot(s, peer_base, offset, rank, world, nextou_sync_slot_ctas(C, 0), epoch, "sync_norm_finalize");
  if (rc) return rc;
  sync_norm_finalize_kernel<<<s.n_cta, 32 * SX_SLICES, 0, (cudaStream_t)stream>>>(partial, nblk, C, c_valid, (double)rows, eps, s, mean,
                                                                                 invstd, running_mean, running_var, momentum,
                                                                                 num_batches_tracked, n_total);
  return check_launch("sync_norm_finalize_kernel");
}

// Backward sums of a SyncBatchNorm layer from this rank's partial rows [nblk][2C]: sums_local (this rank: d beta | d gamma) and
// sums_global (all ranks: what nextou_norm_bwd_apply takes together with the global row count).
extern "C" int nextou_sync_norm_bwd_finalize(const float* partial, int nblk, int C, void* const* peer_base, long long offset, int rank,
                                             int world, unsigned long long* epoch, float* sums_local, float* sums_global,
                                             void* stream) {
  NEXTOU_REQUIRE(partial && sums_local && sums_global && nblk > 0 && C > 0, "sync_norm_bwd_finalize: bad arguments");
  SyncSlot s;
  int rc = make_slot(s, peer_base, offset, rank, world, nextou_sync_slot_ctas(C, 1), epoch, "sync_norm_bwd_finalize");
  if (rc) return rc;
  sync_norm_bwd_finalize_kernel<<<s.n_cta, 32 * SX_SLICES, 0, (cudaStream_t)stream>>>(partial, nblk, 2 * C, s, sums_local, sums_global);
  return check_launch("sync_norm_bwd_finalize_kernel");
}
