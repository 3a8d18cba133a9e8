// Latency microbenchmark for the grouped FFMA GEMM at LunarLander shapes: a CUDA graph of chained launches,
// time per launch = graph time / launches.   Build (from repo root):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o scripts/micro/simt_lat scripts/micro/simt_lat.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#define GS_PROF 1
#include "../../distributed_sac_b200/csrc/gemm_simt.cuh"
using namespace bsac;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

static GemmProb prob(int mode, float* A, float* B, float* bias, float* mask, float* C, float* C2, int M, int N, int K, int lda, int ldb, int ldc) {
  GemmProb p; memset(&p, 0, sizeof(p));
  p.A = A; p.B = B; p.bias = bias; p.mask = mask; p.C = C; p.C2 = C2; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.ldmask = ldc; p.mode = mode; p.relu = mode == GEMM_FWD;
  return p;
}
int main() {
  const size_t n = 1 << 22;
  float* buf; CK(cudaMalloc(&buf, n * 4 * 8)); CK(cudaMemset(buf, 0, n * 4 * 8));
  float* X = buf, *H0 = buf + n, *H1 = buf + 2 * n, *W = buf + 3 * n, *Bi = buf + 4 * n, *G = buf + 5 * n, *D = buf + 6 * n;
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  struct Case { const char* name; GemmGroup g; dim3 grid; };
  std::vector<Case> cases;
  { Case c; c.name = "fwd L1  K=256 (512+256+256 rows x 256)"; memset(&c.g, 0, sizeof(c.g)); c.g.G = 3;
    c.g.p[0] = prob(GEMM_FWD, H0, W, Bi, nullptr, H1, nullptr, 512, 256, 256, 256, 256, 256);
    c.g.p[1] = prob(GEMM_FWD, H0 + 512 * 256, W + 65536, Bi + 256, nullptr, H1 + 512 * 256, nullptr, 256, 256, 256, 256, 256, 256);
    c.g.p[2] = prob(GEMM_FWD, H0 + 768 * 256, W + 131072, Bi + 512, nullptr, H1 + 768 * 256, nullptr, 256, 256, 256, 256, 256, 256);
    c.grid = dim3(8, 16, 3); cases.push_back(c); }
  { Case c; c.name = "fwd L0  K=8/10"; memset(&c.g, 0, sizeof(c.g)); c.g.G = 3;
    c.g.p[0] = prob(GEMM_FWD, X, W, Bi, nullptr, H0, nullptr, 512, 256, 8, 8, 8, 256);
    c.g.p[1] = prob(GEMM_FWD, X + 8192, W + 65536, Bi + 256, nullptr, H0 + 512 * 256, nullptr, 256, 256, 10, 12, 12, 256);
    c.g.p[2] = prob(GEMM_FWD, X + 16384, W + 131072, Bi + 512, nullptr, H0 + 768 * 256, nullptr, 256, 256, 10, 12, 12, 256);
    c.grid = dim3(8, 16, 3); cases.push_back(c); }
  { Case c; c.name = "wgrad L1 x2 + dgrad L1 x2 (256^3)"; memset(&c.g, 0, sizeof(c.g)); c.g.G = 4;
    c.g.p[0] = prob(GEMM_WGRAD, D, H0, nullptr, nullptr, G, G + 300000, 256, 256, 256, 256, 256, 256);
    c.g.p[1] = prob(GEMM_WGRAD, D + 65536, H0 + 65536, nullptr, nullptr, G + 65536, G + 300256, 256, 256, 256, 256, 256, 256);
    c.g.p[2] = prob(GEMM_DGRAD, D, W, nullptr, H0, H1, nullptr, 256, 256, 256, 256, 256, 256);
    c.g.p[3] = prob(GEMM_DGRAD, D + 65536, W + 65536, nullptr, H0 + 65536, H1 + 65536, nullptr, 256, 256, 256, 256, 256, 256);
    c.grid = dim3(8, 8, 4); cases.push_back(c); }
  { Case c; c.name = "wgrad L0 x2 (256 x 10, K=256)"; memset(&c.g, 0, sizeof(c.g)); c.g.G = 2;
    c.g.p[0] = prob(GEMM_WGRAD, D, X, nullptr, nullptr, G, G + 300000, 256, 10, 256, 256, 12, 12);
    c.g.p[1] = prob(GEMM_WGRAD, D + 65536, X + 8192, nullptr, nullptr, G + 65536, G + 300256, 256, 10, 256, 256, 12, 12);
    c.grid = dim3(1, 8, 2); cases.push_back(c); }
  const int L = 40;
  for (auto& c : cases) {
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < L; ++i) gemm_simt_kernel<<<c.grid, GS_THREADS, 0, st>>>(c.g);
    CK(cudaStreamEndCapture(st, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(a, st); cudaGraphLaunch(ge, st); cudaEventRecord(b, st); CK(cudaStreamSynchronize(st));
      float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("%-44s grid %3d CTAs: %.2f us per launch\n", c.name, c.grid.x * c.grid.y * c.grid.z, best * 1e3 / L);
    {  // intra-kernel marks of the last graph replay: entry -> tiles in smem -> k loop done -> stores issued -> next entry
      static unsigned long long hp[4096]; int hn = 0;
      CK(cudaMemcpyFromSymbol(hp, g_gs_prof, sizeof(hp))); CK(cudaMemcpyFromSymbol(&hn, g_gs_prof_n, 4));
      const int base = hn - L; double d[4] = {0, 0, 0, 0}; int cnt = 0;
      for (int i = base + 5; i + 1 < hn && i < 1023; ++i, ++cnt) {
        d[0] += hp[i * 4 + 1] - hp[i * 4 + 0]; d[1] += hp[i * 4 + 2] - hp[i * 4 + 1]; d[2] += hp[i * 4 + 3] - hp[i * 4 + 2];
        d[3] += hp[(i + 1) * 4 + 0] - hp[i * 4 + 3];
      }
      printf("      CTA0: load %.2f us | k loop %.2f | epilogue %.2f | end -> next kernel entry %.2f\n", d[0] / cnt / 1e3, d[1] / cnt / 1e3,
             d[2] / cnt / 1e3, d[3] / cnt / 1e3);
      int zero = 0; CK(cudaMemcpyToSymbol(g_gs_prof_n, &zero, 4));
    }
  }
  return 0;
}
