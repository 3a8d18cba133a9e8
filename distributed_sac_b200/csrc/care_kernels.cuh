// CARE(M) state-encoder stages that are not plain GEMMs (MT10_Distributed_CARE/src/state_encoder.py,
// context_encoder.py with use_modified_care = true):
//
//   care_tables_kernel       per TASK (not per sample): z = E[t] (frozen RoBERTa row), attention
//                            att = softmax(trunk(z)) and context code ctx = mlp_context(z).  The reference
//                            evaluates these 768-wide MLPs on all B rows (state_encoder.py:85-93); they only
//                            depend on the task id, so T rows suffice.
//   care_mix_kernel          encoded state of a row: [ctx[t] | sum_k att[t][k] Zk[k][row] / sum_k att[t][k]]
//                            (state_encoder.py:88-94), written straight into the consumer MLP's input buffer(s).
//   care_mix_bwd_kernel      backward of the mix for the critic update: dZk, d(att) per row.
//   care_tab_reduce_kernel   per-task sums of d(att), d(ctx) (fixed order), softmax backward.
//   care_tab_wgrad_kernel    backward of trunk / mlp_context on the T task rows -> weight and bias gradients.
// The mixture-of-encoders layers themselves (einsum 'kio,bi->kbo', state_encoder.py:155-174) run as K grouped
// problems of the GEMM engines; the mixture weights are stored [k][out][in] here (the reference holds [k][in][out]).
#pragma once
#include "common.cuh"
#include "sac_kernels.cuh"

namespace bsac {

constexpr int kCareMaxLayers = 9;     // <= 8 hidden + output
constexpr int kCareMaxT = 64;

struct CareNet {                      // a small MLP given by offsets into one replica's parameter arena
  int n;                              // number of Linear layers
  int dims[kCareMaxLayers + 1];       // dims[0] = input width
  long long w[kCareMaxLayers], b[kCareMaxLayers];
  int act_off[kCareMaxLayers];        // where layer j's output lives inside a task row of the table
};

struct CareTabArgs {
  const float* params; long long rsP;
  long long inst_delta[2];            // parameter offset of the encoder instance (0 = critic's, target_delta = target's)
  float* tab[2]; long long rsTab;     // table block per instance: [T][row_w]
  long long emb_off;
  CareNet trunk, ctx;
  int T, K, row_w, off_att;
  int original;                       // CARE(O): ctx = the shared, trainable context encoder on relu(E[t]); the trunk reads its output
};

// grid (T, n_inst, R), block 512: 16 warps, one output neuron per warp at a time, up to 24 independent
// loads in flight per lane (the 768-long dot products are pure latency otherwise)
__global__ void __launch_bounds__(512) care_tables_kernel(CareTabArgs P) {
  KStamp ks_;
  __shared__ float xe[2048];
  __shared__ float bufA[512], bufB[512];
  const int t = blockIdx.x, inst = blockIdx.y, rep = blockIdx.z;
  const float* __restrict__ par = P.params + rep * P.rsP;
  const float* __restrict__ E = par + P.emb_off + (long long)t * P.ctx.dims[0];
  float* __restrict__ row = P.tab[inst] + rep * P.rsTab + (long long)t * P.row_w;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int cin = P.ctx.dims[0];       // embedding row; == trunk input for CARE(M), the header input for CARE(O)
  for (int i = threadIdx.x; i < cin; i += 512) xe[i] = E[i];
  __syncthreads();
  __shared__ float zc[512];                             // CARE(O): context code z = cenc(relu(E[t])), the trunk's input
  if (P.original) {
    for (int i = threadIdx.x; i < cin; i += 512) xe[i] = fmaxf(xe[i], 0.f);    // Embedding -> ReLU -> header (context_encoder.py:73-77)
    __syncthreads();
  }
  for (int pass = 0; pass < 2; ++pass) {
    const bool is_ctx = (pass == 0);                    // context net first: in CARE(O) the trunk consumes its output
    const CareNet& N = is_ctx ? P.ctx : P.trunk;
    const long long delta = (is_ctx && P.original) ? 0 : P.inst_delta[inst];    // the context encoder has no target copy
    const float* cur = (!is_ctx && P.original) ? zc : xe;
    float* nxt = bufA;
    for (int j = 0; j < N.n; ++j) {
      const int nin = N.dims[j], nout = N.dims[j + 1];
      const float* __restrict__ W = par + delta + N.w[j];
      const float* __restrict__ bb = par + delta + N.b[j];
      if (nin <= 64) {
        // narrow layer: a warp's (up to four) output neurons are evaluated together -- all their weights and biases are
        // requested before the first FMA, so the layer costs one memory round trip instead of one per neuron
        float a[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f}, w0[4], w1[4];
        const float x0 = lane < nin ? cur[lane] : 0.f, x1 = lane + 32 < nin ? cur[lane + 32] : 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int o = warp + 16 * g;
          const float* __restrict__ wr = W + (long long)o * nin;
          w0[g] = (o < nout && lane < nin) ? __ldg(wr + lane) : 0.f;
          w1[g] = (o < nout && lane + 32 < nin) ? __ldg(wr + lane + 32) : 0.f;
          bv[g] = o < nout ? __ldg(bb + o) : 0.f;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int o = warp + 16 * g;
          if (o < nout) {                                  // warp-uniform
            a[g] = fmaf(x0, w0[g], a[g]);
            a[g] = fmaf(x1, w1[g], a[g]);
            float v = warp_sum(a[g]) + bv[g];
            if (j < N.n - 1) v = fmaxf(v, 0.f);
            if (lane == 0) { nxt[o] = v; row[N.act_off[j] + o] = v; }
          }
        }
        for (int o = warp + 64; o < nout; o += 16) {     // wider-than-64-output layers: remaining neurons one at a time
          const float* __restrict__ wr = W + (long long)o * nin;
          float acc = fmaf(x0, lane < nin ? __ldg(wr + lane) : 0.f, 0.f);
          acc = fmaf(x1, lane + 32 < nin ? __ldg(wr + lane + 32) : 0.f, acc);
          float v = warp_sum(acc) + bb[o];
          if (j < N.n - 1) v = fmaxf(v, 0.f);
          if (lane == 0) { nxt[o] = v; row[N.act_off[j] + o] = v; }
        }
      } else
      for (int o = warp; o < nout; o += 16) {
        const float* __restrict__ wr = W + (long long)o * nin;
        const float bo = __ldg(bb + o);                    // bias requested with the weights, not after the reduction
        float a = 0.f;
        int i = lane;
        for (; i + 23 * 32 < nin; i += 24 * 32) {        // 768-wide rows: the whole row in flight (one round trip, not three)
          float wv[24];
#pragma unroll
          for (int u = 0; u < 24; ++u) wv[u] = __ldg(wr + i + u * 32);
#pragma unroll
          for (int u = 0; u < 24; ++u) a = fmaf(cur[i + u * 32], wv[u], a);
        }
        for (; i + 7 * 32 < nin; i += 8 * 32) {
          float wv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) wv[u] = __ldg(wr + i + u * 32);
#pragma unroll
          for (int u = 0; u < 8; ++u) a = fmaf(cur[i + u * 32], wv[u], a);
        }
        for (; i < nin; i += 32) a = fmaf(cur[i], __ldg(wr + i), a);
        a = warp_sum(a) + bo;
        if (j < N.n - 1) a = fmaxf(a, 0.f);
        if (lane == 0) { nxt[o] = a; row[N.act_off[j] + o] = a; }
      }
      __syncthreads();
      cur = nxt;
      nxt = (nxt == bufA) ? bufB : bufA;
    }
    if (is_ctx && P.original) {
      for (int i = threadIdx.x; i < N.dims[N.n]; i += 512) zc[i] = cur[i];
      __syncthreads();
    }
    if (!is_ctx && threadIdx.x == 0) {                  // softmax over the K logits (F.softmax, dim=-1)
      const float* lg = cur;
      float mx = lg[0];
      for (int k = 1; k < P.K; ++k) mx = fmaxf(mx, lg[k]);
      float s = 0.f, e[32];
      for (int k = 0; k < P.K; ++k) { e[k] = expf(lg[k] - mx); s += e[k]; }
      for (int k = 0; k < P.K; ++k) row[P.off_att + k] = e[k] / s;
    }
    __syncthreads();
  }
}

struct CareMixArgs {
  const float* Z; long long rsZ, kstride; int ldz;      // Zk[k] = Z + k*kstride, [rows][ldz]
  const float* tab; long long rsTab; int row_w, off_att, off_ctx;
  const int* tid; long long rsR;
  int rows, B, K, mo, co;
  float* dst1; long long rsD1; int ld1;                 // every row r -> dst1[r]
  float* dst2; long long rsD2; int ld2; int row_off2;   // rows r >= row_off2 -> dst2[r - row_off2]   (nullable)
};

// one warp per row; grid (ceil(rows/8), R), block 256
__global__ void __launch_bounds__(256) care_mix_kernel(CareMixArgs P) {
  KStamp ks_;
  const int rep = blockIdx.y, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int r = blockIdx.x * 8 + warp;
  if (r >= P.rows) return;
  const int t = (P.tid + rep * P.rsR)[r % P.B];        // the CURRENT observation's task, also for the s' rows
  const float* __restrict__ trow = P.tab + rep * P.rsTab + (long long)t * P.row_w;
  float den = 0.f;
  for (int k = 0; k < P.K; ++k) den += trow[P.off_att + k];       // alpha.sum(dim=1)
  float* d1 = P.dst1 + rep * P.rsD1 + (long long)r * P.ld1;
  float* d2 = (P.dst2 != nullptr && r >= P.row_off2) ? P.dst2 + rep * P.rsD2 + (long long)(r - P.row_off2) * P.ld2 : nullptr;
  for (int j = lane; j < P.co; j += 32) {
    const float v = trow[P.off_ctx + j];
    d1[j] = v;
    if (d2) d2[j] = v;
  }
  const float* __restrict__ Z = P.Z + rep * P.rsZ + (long long)r * P.ldz;
  for (int j = lane; j < P.mo; j += 32) {
    float num = 0.f;
    for (int k = 0; k < P.K; ++k) num += Z[k * P.kstride + j] * trow[P.off_att + k];    // (z_encs * alpha).sum(dim=1)
    const float v = num / den;
    d1[P.co + j] = v;
    if (d2) d2[P.co + j] = v;
  }
}

struct CareMixBwdArgs {
  const float* dx; long long rsDxNet, rsDxRep; int lddx;      // [2][B][xw] gradients wrt the critic input rows
  const float* Z; long long rsZ, kstride; int ldz; int z_row_off;   // Zk rows of the s half start at z_row_off
  const float* tab; long long rsTab; int row_w, off_att;
  const int* tid; long long rsR;
  float* dZ; long long rsDZ, dkstride; int lddz;               // [K][B][lddz]
  float* datt; long long rsDatt;                               // [B][K]
  int B, K, mo, co;
};

// one warp per row; grid (ceil(B/8), R)
__global__ void __launch_bounds__(256) care_mix_bwd_kernel(CareMixBwdArgs P) {
  KStamp ks_;
  const int rep = blockIdx.y, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int i = blockIdx.x * 8 + warp;
  if (i >= P.B) return;
  const int t = (P.tid + rep * P.rsR)[i];
  const float* __restrict__ trow = P.tab + rep * P.rsTab + (long long)t * P.row_w;
  const float* __restrict__ d0 = P.dx + rep * P.rsDxRep + (long long)i * P.lddx + P.co;   // d z_enc part
  const float* __restrict__ Z = P.Z + rep * P.rsZ + (long long)(P.z_row_off + i) * P.ldz;
  float att[32], den = 0.f;
  for (int k = 0; k < P.K; ++k) { att[k] = trow[P.off_att + k]; den += att[k]; }
  float dden = 0.f, dak[32];
  for (int k = 0; k < P.K; ++k) dak[k] = 0.f;
  float* dZ = P.dZ + rep * P.rsDZ + (long long)i * P.lddz;
  for (int j = lane; j < P.mo; j += 32) {
    const float dz = d0[j] + d0[P.rsDxNet + j];
    float num = 0.f;
    for (int k = 0; k < P.K; ++k) num += Z[k * P.kstride + j] * att[k];
    const float dnum = dz / den;                       // z_enc = num / den
    dden += -(dz * num) / (den * den);
    for (int k = 0; k < P.K; ++k) {
      dZ[k * P.dkstride + j] = dnum * att[k];
      dak[k] = fmaf(dnum, Z[k * P.kstride + j], dak[k]);
    }
  }
  dden = warp_sum(dden);
  for (int k = 0; k < P.K; ++k) {
    const float s = warp_sum(dak[k]);
    if (lane == 0) (P.datt + rep * P.rsDatt)[(long long)i * P.K + k] = s + dden;
  }
}

struct CareTabReduceArgs {
  const float* datt; long long rsDatt;                  // [B][K]
  const float* dx; long long rsDxNet, rsDxRep; int lddx; // d ctx = dx[net][i][0..co)
  const float* tab; long long rsTab; int row_w, off_att;
  const int* tid; long long rsR;
  float* dtab; long long rsDtab;                         // [T][K + co]: dlogits (after softmax backward) | dctx
  int B, K, co;
};

// grid (T, R), block 1024 = 16 row groups x 64 values; partial sums combined in a fixed order
__global__ void __launch_bounds__(1024) care_tab_reduce_kernel(CareTabReduceArgs P) {
  KStamp ks_;
  __shared__ float part[16][64];
  __shared__ float tot[64];
  const int t = blockIdx.x, rep = blockIdx.y;
  const int v = threadIdx.x % 64, g = threadIdx.x / 64;
  const int nv = P.K + P.co;
  const int* __restrict__ tid = P.tid + rep * P.rsR;
  for (int v0 = 0; v0 < nv; v0 += 64) {
    const int vv = v0 + v;
    float s = 0.f;
    if (vv < nv) {
      const float* __restrict__ src = vv < P.K ? P.datt + rep * P.rsDatt + vv : P.dx + rep * P.rsDxRep + (vv - P.K);
      const long long ld = vv < P.K ? P.K : P.lddx;
      const long long second = vv < P.K ? 0 : P.rsDxNet;
      for (int i = g; i < P.B; i += 16) {
        if (tid[i] != t) continue;
        const float* d = src + (long long)i * ld;
        s += second ? d[0] + d[second] : d[0];
      }
    }
    part[g][v] = s;
    __syncthreads();
    if (g == 0) {
      float a = 0.f;
#pragma unroll
      for (int q = 0; q < 16; ++q) a += part[q][v];
      tot[v] = a;
    }
    __syncthreads();
    float* out = P.dtab + rep * P.rsDtab + (long long)t * nv;
    if (g == 0 && vv < nv && vv >= P.K) out[vv] = tot[v];
    if (v0 == 0 && threadIdx.x == 0) {                 // softmax backward on the K attention gradients (K <= 32 < 64)
      const float* att = P.tab + rep * P.rsTab + (long long)t * P.row_w + P.off_att;
      float dot = 0.f;
      for (int k = 0; k < P.K; ++k) dot += att[k] * tot[k];
      for (int k = 0; k < P.K; ++k) out[k] = att[k] * (tot[k] - dot);
    }
    __syncthreads();
  }
}

struct CareTabWgradArgs {
  const float* params; long long rsP; long long emb_off;
  const float* tab; long long rsTab; int row_w;
  const float* dtab; long long rsDtab;                   // [T][K + co]
  float* grads; long long rsG;
  CareNet trunk, ctx;
  int T, K, co;
  int original, off_ctx;              // CARE(O): trunk layer-0 input = z (table, off_ctx); ctx layer-0 input = relu(E[t])
};

// grid (nblk, R), block 256.  Every CTA re-derives the (tiny) per-task backward chains in shared memory, then
// takes a slice of the weight / bias gradient elements: dW_j[o][i] = sum_t dz_j[t][o] * a_{j-1}[t][i].
__global__ void __launch_bounds__(256) care_tab_wgrad_kernel(CareTabWgradArgs P) {
  KStamp ks_;
  extern __shared__ float sm[];
  const int rep = blockIdx.y, T = P.T;
  const float* __restrict__ par = P.params + rep * P.rsP;
  const float* __restrict__ tab = P.tab + rep * P.rsTab;
  const float* __restrict__ dtab = P.dtab + rep * P.rsDtab;
  // dz storage: for net n, layer j: dz[n][j] is [T][dims[j+1]]
  int dz_off[2][kCareMaxLayers];
  int cur = 0;
  for (int n = 0; n < 2; ++n) {
    const CareNet& N = n == 0 ? P.trunk : P.ctx;
    for (int j = 0; j < N.n; ++j) { dz_off[n][j] = cur; cur += T * N.dims[j + 1]; }
  }
  for (int n = 0; n < 2; ++n) {
    const CareNet& N = n == 0 ? P.trunk : P.ctx;
    const int last = N.n - 1, nlast = N.dims[N.n];
    for (int e = threadIdx.x; e < T * nlast; e += 256) {
      const int t = e / nlast, o = e % nlast;
      sm[dz_off[n][last] + e] = dtab[(long long)t * (P.K + P.co) + (n == 0 ? o : P.K + o)];
    }
    __syncthreads();
    for (int j = last; j >= 1; --j) {                  // dz_{j-1} = (dz_j W_j) * [a_{j-1} > 0]
      const int nin = N.dims[j], nout = N.dims[j + 1];
      const float* __restrict__ W = par + N.w[j];
      for (int e = threadIdx.x; e < T * nin; e += 256) {
        const int t = e / nin, i = e % nin;
        float s = 0.f;
        for (int o = 0; o < nout; ++o) s = fmaf(sm[dz_off[n][j] + t * nout + o], __ldg(W + (long long)o * nin + i), s);
        const float a = tab[(long long)t * P.row_w + N.act_off[j - 1] + i];
        sm[dz_off[n][j - 1] + e] = a > 0.f ? s : 0.f;
      }
      __syncthreads();
    }
  }
  // gradient elements of every (net, layer): [weights | bias], grid-strided per layer
  const long long gstart = (long long)blockIdx.x * 256 + threadIdx.x, gstride = (long long)gridDim.x * 256;
  for (int n = 0; n < 2; ++n) {
    const CareNet& N = n == 0 ? P.trunk : P.ctx;
    for (int j = 0; j < N.n; ++j) {
      const int nin = N.dims[j], nout = N.dims[j + 1];
      const long long nw = (long long)nin * nout, tot = nw + nout;
      for (long long q = gstart; q < tot; q += gstride) {
        float s = 0.f;
        if (q < nw) {
          const int o = (int)(q / nin), i = (int)(q % nin);
          for (int t = 0; t < T; ++t) {
            float a;
            if (j > 0) a = tab[(long long)t * P.row_w + N.act_off[j - 1] + i];
            else if (P.original && n == 0) a = tab[(long long)t * P.row_w + P.off_ctx + i];          // trunk(z.detach())
            else {
              a = par[P.emb_off + (long long)t * nin + i];
              if (P.original) a = fmaxf(a, 0.f);                                                  // header(relu(E))
            }
            s = fmaf(sm[dz_off[n][j] + t * nout + o], a, s);
          }
          (P.grads + rep * P.rsG)[N.w[j] + q] = s;
        } else {
          const int o = (int)(q - nw);
          for (int t = 0; t < T; ++t) s += sm[dz_off[n][j] + t * nout + o];
          (P.grads + rep * P.rsG)[N.b[j] + o] = s;
        }
      }
    }
  }
}

}  // namespace bsac
