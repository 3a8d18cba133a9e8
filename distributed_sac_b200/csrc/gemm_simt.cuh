// Grouped fp32 FFMA GEMM for the MLP layers of the SAC step (exact-fp32 path).
//
// One launch runs G independent problems x R replicas (blockIdx.z = rep * G + g), so the
// twin critics / target critics / the actor's 2B-row batch are a single grid.  Three
// problem kinds share one tile engine and differ only in how operand tiles are fetched:
//
//   FWD   C[m][n] = act( sum_k A[m][k] * W[n][k] + bias[n] )        nn.Linear forward
//                   (reference: LunarLander_Distributed_SAC/src/model.py:41-44,119-125)
//   DGRAD C[m][n] = ( sum_k dY[m][k] * W[k][n] ) * [mask[m][n] > 0]   input gradient, ReLU' fused
//   WGRAD C[m][n] = sum_k dY[k][m] * X[k][n] ;  C2[m] = sum_k dY[k][m]   weight + bias gradient
//
// Reductions run over the full K inside one CTA in a fixed order (no split-K, no atomics),
// so results are bit-reproducible run to run and replica to replica.
#pragma once
#include "common.cuh"

namespace bsac {

enum { GEMM_FWD = 0, GEMM_DGRAD = 1, GEMM_WGRAD = 2 };

struct GemmProb {
  const float* A;
  const float* B;
  const float* bias;   // FWD only (may be null)
  const float* mask;   // DGRAD only (may be null): activation whose >0 gates the gradient
  float* C;
  float* C2;           // WGRAD only (may be null): bias gradient
  long long rsA, rsB, rsBias, rsMask, rsC, rsC2;   // per-replica strides (floats)
  int M, N, K;
  int lda, ldb, ldc, ldmask;
  int mode;
  int relu;
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(const GemmProb* __restrict__ probs, int G) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int LA = (BM * BK) / NT;   // A elements per thread per k-tile
  constexpr int LB = (BN * BK) / NT;
  static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0, "tile/threads mismatch");
  static_assert(TM % 2 == 0 && TN % 2 == 0, "micro tile");

  const int g = blockIdx.z % G, rep = blockIdx.z / G;
  const GemmProb P = probs[g];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= P.M || n0 >= P.N) return;

  const float* __restrict__ A = P.A + (long long)rep * P.rsA;
  const float* __restrict__ Bm = P.B + (long long)rep * P.rsB;
  const bool a_kc = (P.mode != GEMM_WGRAD);   // A element (m,k) at A[m*lda+k] else A[k*lda+m]
  const bool b_kc = (P.mode == GEMM_FWD);     // B element (n,k) at B[n*ldb+k] else B[k*ldb+n]
  const int M = P.M, N = P.N, K = P.K, lda = P.lda, ldb = P.ldb;

  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;   // WGRAD bias gradient (threads tid < BM of the n-tile-0 CTAs)
  const bool do_bsum = (P.mode == GEMM_WGRAD) && (P.C2 != nullptr) && (blockIdx.x == 0);

  float ra[LA], rb[LB];

  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int e = tid + i * NT, mm, kk;
      if (a_kc) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < M && k < K) v = a_kc ? __ldg(A + (long long)m * lda + k) : __ldg(A + (long long)k * lda + m);
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int e = tid + i * NT, nn, kk;
      if (b_kc) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < N && k < K) v = b_kc ? __ldg(Bm + (long long)n * ldb + k) : __ldg(Bm + (long long)k * ldb + n);
      rb[i] = v;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int e = tid + i * NT, mm, kk;
      if (a_kc) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      As[kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int e = tid + i * NT, nn, kk;
      if (b_kc) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      Bs[kk][nn] = rb[i];
    }
  };

  fetch(0);
  stash();
  __syncthreads();
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = (k0 + BK) < K;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (do_bsum && tid < BM) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) bsum += As[kk][tid];
    }
    __syncthreads();
    if (more) {
      stash();
      __syncthreads();
    }
  }

  float* __restrict__ C = P.C + (long long)rep * P.rsC;
  const float* __restrict__ bias = P.bias ? P.bias + (long long)rep * P.rsBias : nullptr;
  const float* __restrict__ mask = P.mask ? P.mask + (long long)rep * P.rsMask : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (P.mode == GEMM_FWD) {
        if (bias) v += bias[n];
        if (P.relu) v = fmaxf(v, 0.f);
      } else if (P.mode == GEMM_DGRAD) {
        if (mask && !(mask[(long long)m * P.ldmask + n] > 0.f)) v = 0.f;
      }
      C[(long long)m * P.ldc + n] = v;
    }
  }
  if (do_bsum && tid < BM && (m0 + tid) < M) (P.C2 + (long long)rep * P.rsC2)[m0 + tid] = bsum;
}

}  // namespace bsac
