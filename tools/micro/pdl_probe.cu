// Probe: cost of a kernel boundary inside a CUDA graph with and without programmatic dependent launch (PDL).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/pdl_probe tools/micro/pdl_probe.cu && tools/micro/pdl_probe
#include <cstdio>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void step_kernel(float* buf, int n, int work) {
  if (PDL) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = buf[i];
    for (int k = 0; k < work; ++k) v = v * 1.0001f + 0.5f;
    buf[(i + 1) % n] = v;   // depends on the previous kernel's writes (different element: needs the ordering)
  }
}

template <bool PDL>
static void launch(float* buf, int n, int work, int blocks, cudaStream_t st) {
  if (!PDL) {
    step_kernel<false><<<blocks, 256, 0, st>>>(buf, n, work);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, step_kernel<true>, buf, n, work);
}

template <bool PDL>
static float run(int blocks, int work, int chain) {
  const int n = blocks * 256;
  float* buf;
  cudaMalloc(&buf, n * sizeof(float));
  cudaMemset(buf, 0, n * sizeof(float));
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < chain; ++i) launch<PDL>(buf, n, work, blocks, st);
  cudaError_t e = cudaStreamEndCapture(st, &g);
  if (e != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(e)); return -1; }
  e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return -1; }
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEventRecord(a, st);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(b, st);
  cudaStreamSynchronize(st);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  cudaFree(buf);
  return ms * 1000.f / (10 * chain);
}

int main() {
  const int chain = 400;
  for (int blocks : {8, 148, 592, 2368})
    for (int work : {0, 2000}) {
      const float a = run<false>(blocks, work, chain), b = run<true>(blocks, work, chain);
      printf("blocks %5d work %5d: plain %.2f us/kernel, PDL %.2f us/kernel\n", blocks, work, a, b);
    }
  return 0;
}
