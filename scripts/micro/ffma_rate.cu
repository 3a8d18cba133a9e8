// FP32 FMA issue rate per SM on B200 for the register pattern of the chained-MLP inner loop (8 x 8 accumulator tile per
// lane, operands in registers): scalar FFMA vs packed FFMA2, 8 / 16 / 32 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu && ./ffma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int MODE>
__global__ void k(float* out, const float* in, int iters, long long* cyc) {
  float a[8], w[8];
  for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + 32 * i]; w[i] = in[threadIdx.x + 32 * i + 7]; }
  float acc[8][8];
  unsigned long long acc2[8][4];
  for (int m = 0; m < 8; ++m)
    for (int i = 0; i < 8; ++i) acc[m][i] = 0.f;
  for (int m = 0; m < 8; ++m)
    for (int i = 0; i < 4; ++i) acc2[m][i] = 0ull;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[m][i] = fmaf(a[m], w[i], acc[m][i]);
    } else {
      unsigned long long w2[4] = {pack2(w[0], w[1]), pack2(w[2], w[3]), pack2(w[4], w[5]), pack2(w[6], w[7])};
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const unsigned long long a2 = pack2(a[m], a[m]);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc2[m][i] = fma2(a2, w2[i], acc2[m][i]);
      }
    }
    // keep the operands changing a little so nothing is hoisted
    a[it & 7] += 1e-9f;
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int m = 0; m < 8; ++m)
    for (int i = 0; i < 8; ++i) s += acc[m][i];
  for (int m = 0; m < 8; ++m)
    for (int i = 0; i < 4; ++i) s += (float)(acc2[m][i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float *out, *in;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
  cudaMalloc(&cyc, 8);
  const int iters = 4000;
  for (int mode = 0; mode < 2; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(out, in, iters, cyc);
        else k<1><<<148, threads>>>(out, in, iters, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double fma = 64.0 * iters * threads;          // per SM
      printf("%s  %4d threads/SM: %9lld cycles  -> %6.1f FMA/clk/SM\n", mode == 0 ? "FFMA " : "FFMA2", threads, c, fma / c);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
