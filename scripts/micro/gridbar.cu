// Microbenchmark: cost of a grid-wide barrier inside one persistent kernel vs a kernel boundary inside a CUDA graph.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gridbar gridbar.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
  }
  __syncthreads();
}
__global__ void persistent(unsigned* ctr, int iters, float* sink) {
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    acc += sink[(blockIdx.x * blockDim.x + threadIdx.x) & 1023];           // one dependent load per phase
    grid_barrier(ctr, (unsigned)(i + 1) * gridDim.x);
  }
  if (acc == 123.f) sink[0] = acc;
}
__global__ void tiny(float* sink) {
  float v = sink[(blockIdx.x * blockDim.x + threadIdx.x) & 1023];
  if (v == 123.f) sink[0] = v;
}
int main() {
  unsigned* ctr; float* sink; CK(cudaMalloc(&ctr, 4)); CK(cudaMalloc(&sink, 4096)); CK(cudaMemset(sink, 0, 4096));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 2000;
  for (int ctas : {74, 148, 296, 592}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaMemsetAsync(ctr, 0, 4, st));
      cudaEventRecord(a, st);
      persistent<<<ctas, 256, 0, st>>>(ctr, iters, sink);
      cudaEventRecord(b, st); CK(cudaStreamSynchronize(st));
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep) printf("grid barrier, %3d CTAs x 256 thr: %.3f us per phase\n", ctas, ms * 1e3 / iters);
    }
  }
  for (int ctas : {1, 148, 592}) {
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < 200; ++i) tiny<<<ctas, 256, 0, st>>>(sink);
    CK(cudaStreamEndCapture(st, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a, st); for (int k = 0; k < 10; ++k) cudaGraphLaunch(ge, st); cudaEventRecord(b, st); CK(cudaStreamSynchronize(st));
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep == 2) printf("graph of tiny kernels, %3d CTAs: %.3f us per kernel boundary\n", ctas, ms * 1e3 / 2000);
    }
  }
  return 0;
}
