// Microbenchmark: issue-to-completion time of back-to-back tcgen05.mma kind::tf32 instructions on one SM,
// A from shared memory (SS) vs A from tensor memory (TS), for several tile shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/micro/mma_rate scripts/micro/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "../../distributed_sac_b200/csrc/gemm_tc.cuh"
using namespace bsac;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

B200_D void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int M, int N, int ts, int nmma, int same_acc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  float* f = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) f[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tmem_base;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const uint32_t kofs = (uint32_t)(i & 3) * 32;                   // 4 k-steps of 8 tf32 inside a 128-B swizzle row
      const uint64_t ad = tc_smem_desc(sa + kofs, 16, 1024, 2), bd = tc_smem_desc(sb + kofs, 16, 1024, 2);
      const uint32_t d = tb + (same_acc ? 0 : (uint32_t)((i >> 2) & 1) * 256);
      if (ts) mma_ts(d, tb + 256 + (uint32_t)(i & 3) * 8 + (same_acc ? 0 : 0), bd, idesc, i > 0);
      else tc_mma_tf32(d, ad, bd, idesc, i > 0);
    }
    const long long t1 = clock64();
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
  long long* d; CK(cudaMalloc(&d, 64));
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int shapes[][2] = {{128, 64}, {128, 128}, {128, 256}, {64, 64}, {64, 128}};
  for (auto& s : shapes)
    for (int ts = 0; ts < 2; ++ts)
      for (int same = 0; same < 2; ++same) {
        if (ts && same == 0 && s[1] == 256) { }     // D at 0/256 + A at 256.. overlaps for N=256: timing only
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 2; ++rep) {
          rate_kernel<<<1, 128, 64 * 1024>>>(d, s[0], s[1], ts, 256, same);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("M=%d N=%d ts=%d: %s\n", s[0], s[1], ts, cudaGetErrorString(e)); return 1; }
          CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
        }
        printf("M=%3d N=%3d K=8 tf32  A from %s, %s accumulator(s): issue %.1f cyc/MMA, complete %.1f cyc/MMA  (math floor %.0f)\n", s[0], s[1],
               ts ? "TMEM" : "smem", same ? "one" : "two alternating", h[0] / 256.0, h[1] / 256.0, s[0] * s[1] * 8 / 2048.0);
      }
  return 0;
}
