// Optimizer side of the training step: what upstream nnU-Net's train_step does after backward —
//   torch.nn.utils.clip_grad_norm_(parameters, 12)  +  torch.optim.SGD(momentum 0.99, nesterov, weight_decay 3e-5).step()
// — for ALL parameter tensors in three launches, and the bf16 tcgen05 operand packs of the updated weights in a fourth
// (replacing the 95 per-layer pack launches of the next forward pass):
//   opt_sqnorm_kernel   : per-CTA partial sums of grad^2 over the flat gradient segments
//   opt_clip_coef_kernel: total norm -> clip coefficient min(1, max_norm / (norm + 1e-6))   (device scalar, no host sync)
//   opt_sgd_kernel      : multi-tensor update  g' = coef*g + wd*p;  buf = mu*buf + (1-damp)*g';  p -= lr * (nesterov ? g' + mu*buf : buf)
//   opt_pack_kernel     : every registered weight -> forward operand A[r][t][c] and data-gradient operand Bt[c][t'][r] (bf16)
// Everything the kernels need (pointers, sizes, pack geometry) sits in device tables built once by the caller; the learning
// rate is read from device memory, so a captured CUDA graph follows the LR schedule.
#include "common.cuh"

namespace nextou {

constexpr int OPT_THREADS = 256;
constexpr int OPT_CHUNK = 16384;      // elements per CTA chunk of the multi-tensor update

__global__ void __launch_bounds__(OPT_THREADS)
    opt_sqnorm_kernel(const NextouOptTensor* __restrict__ tensors, const NextouOptChunk* __restrict__ chunks, int n_chunks,
                      double* __restrict__ partial) {
  __shared__ double sh[OPT_THREADS / 32];
  double acc = 0.0;
  for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
    const NextouOptChunk ch = chunks[ci];
    const float* g = tensors[ch.tensor].grad + ch.start;
    float a = 0.f;
    const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
    int i = threadIdx.x * 4;
    if (vec)
      for (; i + 4 <= ch.count; i += OPT_THREADS * 4) {
        const float4 v = *reinterpret_cast<const float4*>(g + i);
        a = fmaf(v.x, v.x, a); a = fmaf(v.y, v.y, a); a = fmaf(v.z, v.z, a); a = fmaf(v.w, v.w, a);
      }
    else
      for (; i + 4 <= ch.count; i += OPT_THREADS * 4)
#pragma unroll
        for (int e = 0; e < 4; ++e) a = fmaf(g[i + e], g[i + e], a);
    // ragged tail of the chunk (at most 3 elements, owned by one thread)
    if (i < ch.count && i + 4 > ch.count)
      for (int e = i; e < ch.count; ++e) a = fmaf(g[e], g[e], a);
    acc += (double)a;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < OPT_THREADS / 32; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}

// state[0] = total gradient norm (what clip_grad_norm_ returns), state[1] = clip coefficient
__global__ void opt_clip_coef_kernel(const double* __restrict__ partial, int n, float max_norm, float* __restrict__ state) {
  __shared__ double sh[32];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sh[w];
    const float norm = (float)sqrt(tot);
    state[0] = norm;
    float coef = 1.f;
    if (max_norm > 0.f) {
      coef = max_norm / (norm + 1e-6f);        // torch.nn.utils.clip_grad_norm_: clamp(max_norm / (total_norm + 1e-6), max=1)
      if (coef > 1.f) coef = 1.f;
    }
    state[1] = coef;
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
    opt_sgd_kernel(const NextouOptTensor* __restrict__ tensors, const NextouOptChunk* __restrict__ chunks, int n_chunks,
                   const float* __restrict__ lr_ptr, const float* __restrict__ state, float momentum, float dampening,
                   float weight_decay, int nesterov, int write_grad) {
  const float lr = *lr_ptr;
  const float coef = state ? state[1] : 1.f;
  for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
    const NextouOptChunk ch = chunks[ci];
    const NextouOptTensor t = tensors[ch.tensor];
    float* p = t.param + ch.start;
    float* g = t.grad + ch.start;
    float* m = t.momentum ? t.momentum + ch.start : nullptr;
    for (int i = threadIdx.x; i < ch.count; i += OPT_THREADS) {
      const float w = p[i];
      const float gc = g[i] * coef;                     // the clipped gradient (what param.grad holds after clip_grad_norm_)
      float d = fmaf(weight_decay, w, gc);
      if (m != nullptr) {
        const float b = fmaf(momentum, m[i], (1.f - dampening) * d);
        m[i] = b;
        d = nesterov ? fmaf(momentum, b, d) : b;
      }
      p[i] = fmaf(-lr, d, w);
      if (write_grad) g[i] = gc;
    }
  }
}

// one chunk = OPT_CHUNK consecutive output elements of one pack job (A elements first, then Bt elements); the element
// mapping is pack_element() of common.cuh, shared with the per-layer nextou_pack_weight
__global__ void __launch_bounds__(OPT_THREADS)
    opt_pack_kernel(const NextouPackJob* __restrict__ jobs, const NextouOptChunk* __restrict__ chunks, int n_chunks) {
  for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
    const NextouOptChunk ch = chunks[ci];
    const NextouPackJob j = jobs[ch.tensor];
    for (long long i = ch.start + threadIdx.x; i < ch.start + ch.count; i += OPT_THREADS)
      pack_element(j, reinterpret_cast<const float*>(j.w), i);
  }
}

}  // namespace nextou

using namespace nextou;

extern "C" int nextou_opt_chunk_elems(void) { return OPT_CHUNK; }

extern "C" int nextou_opt_grad_norm(const NextouOptTensor* tensors, const NextouOptChunk* chunks, int n_chunks, float max_norm,
                                    double* partial, int n_partial, float* state, void* stream) {
  NEXTOU_REQUIRE(tensors && chunks && partial && state && n_chunks > 0 && n_partial > 0, "opt_grad_norm: bad arguments");
  int blocks = n_chunks < n_partial ? n_chunks : n_partial;
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  cudaStream_t st = (cudaStream_t)stream;
  opt_sqnorm_kernel<<<blocks, OPT_THREADS, 0, st>>>(tensors, chunks, n_chunks, partial);
  int rc = check_launch("opt_sqnorm_kernel");
  if (rc) return rc;
  opt_clip_coef_kernel<<<1, 256, 0, st>>>(partial, blocks, max_norm, state);
  return check_launch("opt_clip_coef_kernel");
}

extern "C" int nextou_opt_sgd_step(const NextouOptTensor* tensors, const NextouOptChunk* chunks, int n_chunks, const float* lr,
                                   const float* state, float momentum, float dampening, float weight_decay, int nesterov,
                                   int write_clipped_grad, void* stream) {
  NEXTOU_REQUIRE(tensors && chunks && lr && n_chunks > 0, "opt_sgd_step: bad arguments");
  int blocks = n_chunks < 8 * num_sms() ? n_chunks : 8 * num_sms();
  opt_sgd_kernel<<<blocks, OPT_THREADS, 0, (cudaStream_t)stream>>>(tensors, chunks, n_chunks, lr, state, momentum, dampening,
                                                                  weight_decay, nesterov, write_clipped_grad);
  return check_launch("opt_sgd_kernel");
}

extern "C" int nextou_opt_pack_weights(const NextouPackJob* jobs, const NextouOptChunk* chunks, int n_chunks, void* stream) {
  NEXTOU_REQUIRE(jobs && chunks && n_chunks > 0, "opt_pack_weights: bad arguments");
  int blocks = n_chunks < 8 * num_sms() ? n_chunks : 8 * num_sms();
  opt_pack_kernel<<<blocks, OPT_THREADS, 0, (cudaStream_t)stream>>>(jobs, chunks, n_chunks);
  return check_launch("opt_pack_kernel");
}
