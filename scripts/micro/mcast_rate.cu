// Micro-benchmark: how fast can N CTAs stream the SAME weight matrix (256 KB, L2-resident) into shared memory as 32 KB
// chunks -- unicast (every CTA pulls every chunk, what chain_kernel does) versus cluster multicast (each CTA of a cluster
// of C pulls 1/C of a chunk and multicasts it to all C).  Answers: is the ~920-cycle chunk time of the chained LL kernels an
// L2-side limit (then multicast helps) or a per-SM receive limit (then it does not)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mcast_rate mcast_rate.cu && ./mcast_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int CHUNK_BYTES = 32768, NSTAGE = 3, THREADS = 256;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  const long long t0 = clock64();
  while (!mbar_try(b, ph)) if (clock64() - t0 > 2000000000LL) __trap();
}
__device__ __forceinline__ void remote_arrive(uint32_t local_bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

template <int C>
__global__ void __launch_bounds__(THREADS, 1) stream_kernel(const float* __restrict__ W, int chunks_per_pass, int passes, long long* out_cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + NSTAGE * CHUNK_BYTES);
  uint64_t* empty = full + NSTAGE;
  uint32_t rank = 0;
  if (C > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(s32(full + i), 1); mbar_init(s32(empty + i), C); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (C > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  else __syncthreads();
  const int total = chunks_per_pass * passes;
  int issued = 0;
  auto issue = [&]() {
    if (tid == 0 && issued < total) {
      const int slot = issued % NSTAGE, use = issued / NSTAGE;
      if (use > 0) mbar_wait(s32(empty + slot), (uint32_t)((use - 1) & 1));     // every CTA of the cluster is done with the slot
      const char* src = reinterpret_cast<const char*>(W) + (size_t)(issued % chunks_per_pass) * CHUNK_BYTES;
      const uint32_t bar = s32(full + slot), dst = s32(sm + slot * CHUNK_BYTES);
      mbar_expect(bar, CHUNK_BYTES);
      if (C == 1) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(CHUNK_BYTES), "r"(bar) : "memory");
      } else {
        const uint32_t part = CHUNK_BYTES / C;
        const uint16_t mask = (uint16_t)((1u << C) - 1);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                     ::"r"(dst + rank * part), "l"(src + rank * part), "r"(part), "r"(bar), "h"(mask) : "memory");
      }
      ++issued;
    }
  };
  issue(); issue();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int g = 0; g < total; ++g) {
    const int slot = g % NSTAGE;
    mbar_wait(s32(full + slot), (uint32_t)((g / NSTAGE) & 1));
    __syncthreads();
    if (tid == 0) issue();
    // token consumption: every thread reads 16 B of the chunk (keeps the data dependence honest without being the bottleneck)
    const float4 v = *reinterpret_cast<const float4*>(sm + slot * CHUNK_BYTES + tid * 16);
    acc += v.x + v.y + v.z + v.w;
    __syncthreads();                    // the CTA is done with the slot ...
    if (tid == 0) {                     // ... tell every CTA of the cluster
      if (C == 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(empty + slot)) : "memory");
      else for (int r = 0; r < C; ++r) remote_arrive(s32(empty + slot), (uint32_t)r);
    }
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
  if (acc == 12345.678f) *sink = acc;
  if (C > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}


// Unicast only: per-SM receive rate versus bytes in flight (stages x chunk bytes; stages-1 chunks are kept in flight).
__global__ void __launch_bounds__(THREADS, 1) depth_kernel(const float* __restrict__ W, int nstage, int chunk_bytes, int total, long long* out_cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)nstage * chunk_bytes);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) mbar_init(s32(full + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int per_pass = 262144 / chunk_bytes;
  int issued = 0;
  auto issue = [&]() {
    if (tid == 0 && issued < total) {
      const int slot = issued % nstage;
      const char* src = reinterpret_cast<const char*>(W) + (size_t)(issued % per_pass) * chunk_bytes;
      const uint32_t bar = s32(full + slot), dst = s32(sm + (size_t)slot * chunk_bytes);
      mbar_expect(bar, chunk_bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(src), "r"(chunk_bytes), "r"(bar) : "memory");
      ++issued;
    }
  };
  for (int i = 0; i < nstage - 1; ++i) issue();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int g = 0; g < total; ++g) {
    const int slot = g % nstage;
    mbar_wait(s32(full + slot), (uint32_t)((g / nstage) & 1));
    __syncthreads();
    if (tid == 0) issue();
    const float4 v = *reinterpret_cast<const float4*>(sm + (size_t)slot * chunk_bytes + tid * 16);
    acc += v.x + v.y + v.z + v.w;
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
  if (acc == 12345.678f) *sink = acc;
}

static int run_depth(const float* W, int ctas, int nstage, int chunk_bytes, long long* d_cyc, float* d_sink) {
  const size_t smem = (size_t)nstage * chunk_bytes + 256;
  CK(cudaFuncSetAttribute(depth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int total = 16 * (262144 / chunk_bytes);
  for (int it = 0; it < 3; ++it) depth_kernel<<<ctas, THREADS, smem>>>(W, nstage, chunk_bytes, total, d_cyc, d_sink);
  CK(cudaDeviceSynchronize());
  long long cyc = 0;
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  const double per32 = (double)cyc / total * (32768.0 / chunk_bytes);
  printf("  %3d CTAs, %2d stages x %5d B (%3d KB in flight): %7.1f cycles per 32 KB  (%5.1f B/clk/SM, %6.0f B/clk chip-wide)\n", ctas, nstage, chunk_bytes,
         (nstage - 1) * chunk_bytes / 1024, per32, 32768.0 / per32, ctas * 32768.0 / per32);
  return 0;
}

// Unicast: one 32 KB chunk issued as L pieces by L lanes of warp 0 in ONE warp instruction (same mbarrier).  Is the ~675-cycle
// cost of a bulk copy per REQUEST (then pieces from different lanes overlap) or per SM?
__global__ void __launch_bounds__(THREADS, 1) lanes_kernel(const float* __restrict__ W, int nstage, int chunk_bytes, int L, int total, long long* out_cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)nstage * chunk_bytes);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) mbar_init(s32(full + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int per_pass = 262144 / chunk_bytes;
  int issued = 0;
  auto issue = [&]() {          // called by all lanes of warp 0
    if (issued < total) {
      const int slot = issued % nstage;
      const char* src = reinterpret_cast<const char*>(W) + (size_t)(issued % per_pass) * chunk_bytes;
      const uint32_t bar = s32(full + slot), dst = s32(sm + (size_t)slot * chunk_bytes);
      if (tid == 0) mbar_expect(bar, chunk_bytes);
      __syncwarp();
      const int piece = chunk_bytes / L;
      if (tid < L)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + tid * piece), "l"(src + tid * piece), "r"(piece), "r"(bar) : "memory");
      ++issued;
    }
  };
  if (tid < 32) for (int i = 0; i < nstage - 1; ++i) issue();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int g = 0; g < total; ++g) {
    const int slot = g % nstage;
    mbar_wait(s32(full + slot), (uint32_t)((g / nstage) & 1));
    __syncthreads();
    if (tid < 32) issue();
    const float4 v = *reinterpret_cast<const float4*>(sm + (size_t)slot * chunk_bytes + tid * 16);
    acc += v.x + v.y + v.z + v.w;
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
  if (acc == 12345.678f) *sink = acc;
}

static int run_lanes(const float* W, int ctas, int nstage, int chunk_bytes, int L, long long* d_cyc, float* d_sink) {
  const size_t smem = (size_t)nstage * chunk_bytes + 256;
  CK(cudaFuncSetAttribute(lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int total = 16 * (262144 / chunk_bytes);
  for (int it = 0; it < 3; ++it) lanes_kernel<<<ctas, THREADS, smem>>>(W, nstage, chunk_bytes, L, total, d_cyc, d_sink);
  CK(cudaDeviceSynchronize());
  long long cyc = 0;
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  const double per32 = (double)cyc / total * (32768.0 / chunk_bytes);
  printf("  %3d CTAs, %2d stages x %5d B as %2d pieces: %7.1f cycles per 32 KB  (%5.1f B/clk/SM, %6.0f B/clk chip-wide)\n", ctas, nstage, chunk_bytes, L,
         per32, 32768.0 / per32, ctas * 32768.0 / per32);
  return 0;
}

template <int C>
static int run(const float* W, int ctas, long long* d_cyc, float* d_sink) {
  const size_t smem = NSTAGE * CHUNK_BYTES + 128;
  CK(cudaFuncSetAttribute(stream_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const int chunks = 8, passes = 16;
  for (int it = 0; it < 3; ++it) CK(cudaLaunchKernelEx(&cfg, stream_kernel<C>, W, chunks, passes, d_cyc, d_sink));
  CK(cudaDeviceSynchronize());
  long long cyc = 0;
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  printf("  cluster %d, %3d CTAs: %7.1f cycles per 32 KB chunk per CTA  (%5.1f B/clk/SM received, %6.0f B/clk read from L2 chip-wide)\n", C, ctas,
         (double)cyc / (chunks * passes), CHUNK_BYTES / ((double)cyc / (chunks * passes)), (double)ctas / C * CHUNK_BYTES / ((double)cyc / (chunks * passes)));
  return 0;
}

int main() {
  float* W; long long* d_cyc; float* d_sink;
  CK(cudaMalloc(&W, 8 * CHUNK_BYTES)); CK(cudaMemset(W, 0, 8 * CHUNK_BYTES));
  CK(cudaMalloc(&d_cyc, 8)); CK(cudaMalloc(&d_sink, 4));
  for (int ctas : {8, 128}) {
    printf("unicast, chunk issued as pieces by several lanes (%d CTAs):\n", ctas);
    const int cfgs[][3] = {{3, 32768, 1}, {3, 32768, 2}, {3, 32768, 4}, {3, 32768, 8}, {3, 32768, 16}, {3, 32768, 32}, {2, 65536, 8}, {3, 65536, 8}, {6, 16384, 4}};
    for (auto& c : cfgs) if (run_lanes(W, ctas, c[0], c[1], c[2], d_cyc, d_sink)) return 1;
  }
  for (int ctas : {8, 128}) {
    printf("unicast, bytes in flight (%d CTAs):\n", ctas);
    const int cfgs[][2] = {{3, 32768}, {4, 32768}, {6, 32768}, {3, 16384}, {6, 16384}, {12, 16384}, {12, 8192}, {24, 8192}, {2, 65536}, {3, 65536}};
    for (auto& c : cfgs) if (run_depth(W, ctas, c[0], c[1], d_cyc, d_sink)) return 1;
  }
  for (int ctas : {8, 128}) {
    printf("%d CTAs streaming the same 256 KB:\n", ctas);
    if (run<1>(W, ctas, d_cyc, d_sink)) return 1;
    if (run<2>(W, ctas, d_cyc, d_sink)) return 1;
    if (run<4>(W, ctas, d_cyc, d_sink)) return 1;
    if (run<8>(W, ctas, d_cyc, d_sink)) return 1;
  }
  return 0;
}
