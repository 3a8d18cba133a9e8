// tcgen05 (5th-gen tensor core) grouped GEMM for the hidden layers of the SAC step, at
// fp32-class accuracy via the 3xTF32 split:  x = hi + lo (hi = top 19 bits, lo = x - hi),
// D += A_hi*B_lo + A_lo*B_hi + A_hi*B_hi, accumulated in fp32 in TMEM.
//
// One CTA = one 128 x 64 (or 128 x 128) output tile of one problem (blockIdx.z = problem x replica):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes (SWIZZLE_128B for K-major operands,
//               SWIZZLE_128B_ATOM_32B for MN-major ones) of the fp32 operands,
//               straight from the row-major activation / weight / gradient buffers; out-of-range
//               rows and columns are zero-filled by the TMA unit (ragged 400-wide layers, tails)
//   warps 2-5   splitter: emit lo = x - trunc19(x) beside each landed tile (same swizzled offsets, so the
//               layout is untouched); the raw fp32 tile itself serves as hi because the tensor core reads
//               only the top 19 bits of a tf32 operand word (measured); later the epilogue warps
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (kind::tf32, M=128, K=8).  A tf32 MMA costs
//               ~105 cycles whatever its shape up to N=128 (measured), so the 64-wide tile issues TWO per k step:
//               A_hi * [B_hi ; B_lo] (N=128: the lo tile lies right behind the hi tile in shared memory) into the
//               adjacent TMEM column blocks [main_i | cross_i], and A_lo * B_hi (N=64) into cross_i; the 128-wide
//               tile issues three (N=128 each).  tcgen05.commit releases the smem stage / signals the epilogue
//   accumulate  the tensor core adds into the fp32 accumulator with truncation (measured: relative
//               bias ~ #MMAs x 2^-24), so the tiny cross terms A_hi*B_lo + A_lo*B_hi get their own
//               accumulators and the A_hi*B_hi sum is spread over up to four (three) accumulators (contiguous
//               K ranges); the epilogue adds them with round-to-nearest.  Keeps the result within ~1e-6 of fp64.
//   epilogue    tcgen05.ld 32x32b (thread = accumulator row) -> bias+ReLU | ReLU' mask | plain ->
//               global; the weight-gradient kind also emits the bias gradient (column sums of dY),
//               accumulated by the splitter while it walks the A tiles, in a fixed order.
// The three GEMM kinds only differ in operand majors (instruction-descriptor bits 15/16 and the
// smem-descriptor LBO/SBO), exactly as in gemm_simt.cuh:
//   FWD   A = X  [M][K] K-major,  B = W [N][K] K-major
//   DGRAD A = dY [M][K] K-major,  B = W [K][N] MN-major
//   WGRAD A = dY [K][M] MN-major, B = X [K][N] MN-major
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_simt.cuh"

namespace bsac {

constexpr int TC_BM = 128, TC_BK = 32;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;                  // 16 KiB
constexpr int TC_SPLIT_THREADS = 256;                          // warps 2-9: splitter, then epilogue (two warps per TMEM lane quarter)
constexpr int TC_THREADS = 64 + TC_SPLIT_THREADS;
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_BSUM_BYTES = 16 * TC_BM * 4;                  // [16 partials][128 m] fp32

// Tile widths: 128x64 (many CTAs: latency-critical small problems), 128x128 (the 400-wide shapes: an N=128 MMA amortises
// the A-operand shared-memory reads over twice the columns) and 128x160 (400 = 3 x 160 - 80: launch groups whose 128-wide
// tiling needs a second wave of CTAs but whose 160-wide tiling fits one, e.g. 5120 rows x 400 columns = 160 vs 120 tiles).
template <int BN, bool PAIRED = (BN == 64)>
struct TcCfg {
  static constexpr int kBN = BN;
  static constexpr int kBBytes = BN * TC_BK * 4;                           // 8 / 16 / 20 KiB
  static constexpr int kStageBytes = 2 * (TC_A_BYTES + kBBytes);           // hi + lo of both operands: 48 / 64 / 72 KiB
  static constexpr int kStages = BN == 64 ? 4 : 3;
  // PAIRED: accumulators [main_i | cross_i] (2*BN columns each) fed by ONE N=2*BN MMA per k step whose B operand spans the
  // hi tile and the lo tile behind it (A_hi * [B_hi ; B_lo]) plus one N=BN MMA (A_lo * B_hi -> cross_i): 2 instead of 3
  // tcgen05.mma per k step.  An MMA costs ~105 cycles whatever its N <= 128 and ~171 at N = 256 (measured,
  // scripts/micro/mma_rate.cu): 210 instead of 315 cycles per k step at BN = 64, 276 instead of 315 at BN = 128.
  // Unpaired: kNMain rotating main accumulators + 1 shared cross accumulator.
  static constexpr bool kPaired = PAIRED;
  static constexpr int kNMain = PAIRED ? (512 / (2 * BN)) : (BN == 128 ? 3 : 2);   // TMEM: 4*128 | 2*256 | (3+1)*128 | (2+1)*160 columns
  static constexpr int kSmemBytes = kStages * kStageBytes + TC_BSUM_BYTES + 256 + 1024;   // + barriers + align slack
  static_assert(PAIRED ? (2 * BN * kNMain <= TC_TMEM_COLS && 2 * BN <= 256) : (BN * (kNMain + 1) <= TC_TMEM_COLS), "TMEM columns");
  static_assert(kSmemBytes <= 232448, "shared memory");
  static_assert(BN % 32 == 0 && kBBytes % (16 * TC_SPLIT_THREADS) == 0, "tile width");
};
constexpr int TC_BN = 64;            // default tile width (descriptor-eligibility bounds, tests)
constexpr int TC_SMEM_BYTES = TcCfg<64>::kSmemBytes;

struct alignas(128) TcProb {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmC;     // output [M][N] row-major, box 32 x 32, SWIZZLE_128B (valid when c_tma != 0)
  const float* bias;
  const float* mask;
  float* C;
  float* C2;
  int M, N, K;
  int ldc, ldmask;
  int mode, relu;
  int a_mn, b_mn;
  int c_tma;           // 1: the epilogue stores its 32 x 32 blocks with cp.async.bulk.tensor (ldc % 4 == 0, 16-B aligned C)
  long long* dbg;      // optional clock64() timeline of CTA (0,0) (b200sac_tc_gemm_timeline)
};

#define TC_STAMP(i) do { if (dbg) dbg[(i)] = clock64(); } while (0)

// ---- PTX wrappers ---------------------------------------------------------------------------
B200_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

B200_D void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
B200_D void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
B200_D void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
B200_D bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become a trapped launch, never a hung GPU.
B200_D void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
B200_D void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
B200_D void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}
B200_D void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
B200_D void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
B200_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
B200_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
B200_D void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
B200_D void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
B200_D void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
B200_D void tc_ld32_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
B200_D void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

B200_D void stg_f32(float* p, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
B200_D float ldg_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// smem matrix descriptor, Blackwell version field = 1 (cute::UMMA::SmemDescriptor).
// layout_type 2 = SWIZZLE_128B (16-B atoms; K-major operands), 1 = SWIZZLE_128B_BASE32B (32-B atoms;
// the only layout tcgen05 accepts for MN-major tf32 operands).
B200_D uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// number of main accumulators for nk k-chunks: ~2 chunks (8 MMAs) each, at most `cap`
B200_D int tc_nmain(int nk, int cap) {
  int n = nk / 2;
  return n < 1 ? 1 : (n > cap ? cap : n);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, majors, N>>3, M>>4
B200_D uint32_t tc_instr_desc(int a_mn, int b_mn, int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

template <int BN, bool PAIRED>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const TcProb* __restrict__ probs) {
  using Cfg = TcCfg<BN, PAIRED>;
  constexpr int TC_BN = BN, TC_B_BYTES = Cfg::kBBytes, TC_STAGE_BYTES = Cfg::kStageBytes, TC_STAGES = Cfg::kStages,
                TC_NMAIN = Cfg::kNMain;
  KStamp ks_;
  extern __shared__ uint8_t smem_raw[];
  const TcProb* P = probs + blockIdx.z;
  const int M = P->M, N = P->N, K = P->K;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
  if (m0 >= M || n0 >= N) return;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;            // SWIZZLE_128B tiles need 1024-B alignment
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bsum_s = base + TC_STAGES * TC_STAGE_BYTES;
  float* bsum_g = reinterpret_cast<float*>(gbase + TC_STAGES * TC_STAGE_BYTES);
  const uint32_t bars = bsum_s + TC_BSUM_BYTES;
  uint8_t* bars_g = gbase + TC_STAGES * TC_STAGE_BYTES + TC_BSUM_BYTES;
  // barrier slots (8 B each): full[4] ready[4] empty[4] accum[1]; then the TMEM base address word
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto ready_bar = [&](int s) { return bars + 8u * (TC_STAGES + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (2 * TC_STAGES + s); };
  const uint32_t accum_bar = bars + 8u * (3 * TC_STAGES);
  const uint32_t tmem_slot = bars + 8u * (3 * TC_STAGES + 1);
  volatile uint32_t* tmem_slot_g = reinterpret_cast<volatile uint32_t*>(bars_g + 8 * (3 * TC_STAGES + 1));

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nk = (K + TC_BK - 1) / TC_BK;
  long long* dbg = (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x % 32) == 0) ? P->dbg : nullptr;
  if (threadIdx.x == 0) TC_STAMP(0);
  const int a_mn = P->a_mn, b_mn = P->b_mn;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P->tmA) : "memory");    // descriptor fetch overlaps the barrier / TMEM setup
    asm volatile("prefetch.tensormap [%0];" ::"l"(&P->tmB) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(ready_bar(s), TC_SPLIT_THREADS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;
  if (threadIdx.x == 0) TC_STAMP(1);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // MN-major operands need one 32-wide box per 32 rows/columns (up to 8 boxes per stage).  A single thread issues a
    // cp.async.bulk.tensor only every ~250 cycles (measured), which made those stages issue-bound, so the boxes of a
    // stage are issued by different lanes of this warp in ONE warp instruction.
    const int nA = a_mn ? TC_BM / 32 : 1, nB = b_mn ? TC_BN / 32 : 1;
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % TC_STAGES;
      const uint32_t ph = (uint32_t)(kc / TC_STAGES) & 1u;
      if (lane == 0) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (kc < 16) TC_STAMP(2 + kc);
        mbar_expect_tx(full_bar(s), TC_A_BYTES + TC_B_BYTES);
      }
      __syncwarp();
      const uint32_t st = base + (uint32_t)s * TC_STAGE_BYTES;
      const uint32_t dA = st, dB = st + 2 * TC_A_BYTES;
      const int k0 = kc * TC_BK;
      if (lane < nA) {
        if (!a_mn) tma_load_2d(dA, &P->tmA, k0, m0, full_bar(s));
        else tma_load_2d(dA + lane * 4096, &P->tmA, m0 + 32 * lane, k0, full_bar(s));
      } else if (lane < nA + nB) {
        const int g = lane - nA;
        if (!b_mn) tma_load_2d(dB, &P->tmB, k0, n0, full_bar(s));
        else tma_load_2d(dB + g * 4096, &P->tmB, n0 + 32 * g, k0, full_bar(s));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = tc_instr_desc(a_mn, b_mn, TC_BN);
      const int nmain = tc_nmain(nk, TC_NMAIN);          // chunk kc accumulates into main[(kc * nmain) / nk]
      // (kept free of run-time divisions: this single thread's scalar work between two chunks is on the critical path --
      //  ~300 cycles per chunk with the divisions, measured with scripts/tc_timeline.py)
      int mi = 0, acc_pos = 0, acc_lim = nk, s = 0;
      uint32_t ph = 0;
      for (int kc = 0; kc < nk; ++kc) {
        bool new_main = kc == 0;
        if (acc_pos >= acc_lim) { ++mi; acc_lim += nk; new_main = true; }     // mi == (kc * nmain) / nk
        acc_pos += nmain;
        mbar_wait(ready_bar(s), ph);
        if (kc < 16) TC_STAMP(50 + 2 * kc);
        tc_fence_after();
        const uint32_t st = base + (uint32_t)s * TC_STAGE_BYTES;
        // One descriptor per operand per chunk; everything else is an add on the 14-bit address field
        // (units of 16 B): +A_BYTES -> lo copy, +2*A_BYTES -> B tile, k step = 32 B (K-major) / 1024 B (MN-major).
        const uint64_t dA0 = tc_smem_desc(st, a_mn ? 4096u : 16u, a_mn ? 512u : 1024u, a_mn ? 1u : 2u);
        const uint64_t dB0 = tc_smem_desc(st + 2 * TC_A_BYTES, b_mn ? 4096u : 16u, b_mn ? 512u : 1024u, b_mn ? 1u : 2u);
        const uint64_t stepA = a_mn ? (1024u >> 4) : (32u >> 4), stepB = b_mn ? (1024u >> 4) : (32u >> 4);
        if (Cfg::kPaired) {
          const uint32_t d_pair = tmem_base + (uint32_t)(2 * TC_BN * mi);       // [main_mi | cross_mi]
          const uint32_t idesc2 = tc_instr_desc(a_mn, b_mn, 2 * TC_BN);
#pragma unroll
          for (int ks = 0; ks < TC_BK / 8; ++ks) {
            const uint64_t dAh = dA0 + stepA * ks, dAl = dAh + (TC_A_BYTES >> 4);
            const uint64_t dBh = dB0 + stepB * ks;                              // N = 2*BN: hi rows, then the lo tile behind them
            tc_mma_tf32(d_pair, dAh, dBh, idesc2, (new_main && ks == 0) ? 0u : 1u);   // main += Ah*Bh ; cross += Ah*Bl
            tc_mma_tf32(d_pair + TC_BN, dAl, dBh, idesc, 1u);                          // cross += Al*Bh
          }
        } else {
          const uint32_t d_main = tmem_base + (uint32_t)(TC_BN * mi), d_cross = tmem_base + (uint32_t)(TC_BN * TC_NMAIN);
#pragma unroll
          for (int ks = 0; ks < TC_BK / 8; ++ks) {
            const uint64_t dAh = dA0 + stepA * ks, dAl = dAh + (TC_A_BYTES >> 4);
            const uint64_t dBh = dB0 + stepB * ks, dBl = dBh + (TC_B_BYTES >> 4);
            tc_mma_tf32(d_cross, dAh, dBl, idesc, (kc | ks) != 0 ? 1u : 0u);
            tc_mma_tf32(d_cross, dAl, dBh, idesc, 1u);
            tc_mma_tf32(d_main, dAh, dBh, idesc, (new_main && ks == 0) ? 0u : 1u);
          }
        }
        tc_commit(empty_bar(s));          // smem stage reusable once these MMAs have read it
        if (kc < 16) TC_STAMP(51 + 2 * kc);
        if (++s == TC_STAGES) { s = 0; ph ^= 1u; }
      }
      tc_commit(accum_bar);               // accumulator complete
    }
  } else {
    // =============================== splitter, then epilogue ===============================
    const int t = threadIdx.x - 64;       // 0..255
    const bool want_bsum = (P->mode == GEMM_WGRAD) && (P->C2 != nullptr) && (blockIdx.x == 0);
    float bs[4][4];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int q = 0; q < 4; ++q) bs[g][q] = 0.f;

    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % TC_STAGES;
      const uint32_t ph = (uint32_t)(kc / TC_STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      if (t == 0 && kc < 16) TC_STAMP(18 + 2 * kc);
      uint8_t* st = gbase + (size_t)s * TC_STAGE_BYTES;
      float4* aH = reinterpret_cast<float4*>(st);
      float4* aL = reinterpret_cast<float4*>(st + TC_A_BYTES);
      float4* bH = reinterpret_cast<float4*>(st + 2 * TC_A_BYTES);
      float4* bL = reinterpret_cast<float4*>(st + 2 * TC_A_BYTES + TC_B_BYTES);
      // tcgen05 kind::tf32 reads only the top 19 bits of each fp32 operand word (measured: using the raw
      // tile as "hi" is bit-identical to masking it first; a round-to-nearest model is off by 9e-4), so the
      // raw tile stays in place as the hi operand and only lo = x - trunc19(x) is written.
      auto split = [](float4 v, float4& lo) {
        lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
      };
      constexpr int NA = TC_A_BYTES / 16 / TC_SPLIT_THREADS, NB = TC_B_BYTES / 16 / TC_SPLIT_THREADS;   // 4; 2 | 4 | 5
      float4 va[NA], vb[NB];
#pragma unroll
      for (int i = 0; i < NA; ++i) va[i] = aH[t + i * TC_SPLIT_THREADS];   // all loads first
#pragma unroll
      for (int i = 0; i < NB; ++i) vb[i] = bH[t + i * TC_SPLIT_THREADS];
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        float4 lo;
        split(va[i], lo);
        aL[t + i * TC_SPLIT_THREADS] = lo;
        if (want_bsum) {   // MN-major A tile: [g = idx/256][k = (idx%256)/8][slot = idx%8], 32-B chunk (slot/2) ^= (k&3); idx = t + 256 i
          bs[i][0] += va[i].x; bs[i][1] += va[i].y; bs[i][2] += va[i].z; bs[i][3] += va[i].w;
        }
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        float4 lo;
        split(vb[i], lo);
        bL[t + i * TC_SPLIT_THREADS] = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
      mbar_arrive(ready_bar(s));
      if (t == 0 && kc < 16) TC_STAMP(19 + 2 * kc);
    }

    if (want_bsum) {
      // thread t saw, for every g, physical 16-B slot t%8 of row k = t/8; SW128_BASE32B XORs the 32-B chunk index (slot/2)
      // with k&3, so its logical 4-float m-vector is j below.  32 partials per column (one per k row of a chunk), folded in
      // two ordered passes into 16 slots, then summed in slot order: a fixed order, bit-reproducible.
      const int j = ((((t & 7) >> 1) ^ ((t >> 3) & 3)) << 1) | (t & 1), r = t >> 3;   // r = 0..31
      if (r < 16) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int q = 0; q < 4; ++q) bsum_g[r * TC_BM + g * 32 + j * 4 + q] = bs[g][q];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (r >= 16) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int q = 0; q < 4; ++q) bsum_g[(r - 16) * TC_BM + g * 32 + j * 4 + q] += bs[g][q];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int m = m0 + t;
      if (t < TC_BM && m < M) {
        float ssum = 0.f;
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) ssum += bsum_g[rr * TC_BM + t];
        P->C2[m] = ssum;
      }
    }

    // ---- epilogue: TMEM -> registers -> global ----
    const int q = warp & 3;                                   // TMEM lane quarter this warp may touch
    const int cg = (warp - 2) >> 2;                           // two warps per quarter: even / odd 32-column blocks
    const int mode = P->mode, relu = P->relu, ldc = P->ldc;
    const float* __restrict__ bias = P->bias;
    const float* __restrict__ mask = P->mask;
    float* __restrict__ Cout = P->C;
    const int ldmask = P->ldmask;
    // Output pass layout: one warp instruction covers 4 rows x 32 columns, lane = (row r0 + lane/8, column quad lane%8), so
    // global stores and mask loads are 128-bit and a 32x32 block takes 8 of them instead of 32 (measured: ~45 cycles per
    // store instruction in the 32-bit version = 1.5k cycles per block).
    const int rsub = lane >> 3, c4 = (lane & 7) << 2;
    const bool vec_c = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cout) & 15) == 0);
    const bool vec_m = mask && ((ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(mask) & 15) == 0);
    // Output path: the 32 x 32 block of a warp goes to shared memory in the thread = row layout tcgen05.ld delivers (8 swizzled
    // 128-bit stores per thread, conflict-free) and leaves as ONE cp.async.bulk.tensor store that clips ragged edges itself --
    // instead of a transposition through padded shared memory + 8 row stores per lane (1.5 k of the ~2 k cycles of a block,
    // measured).  Falls back to the row-store path when C cannot have a tensor map (ldc % 4 != 0) or the mask is unaligned.
    const bool use_tma = P->c_tma != 0 && (!(P->mode == GEMM_DGRAD && P->mask) || (((P->ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(P->mask) & 15) == 0)));
    constexpr int NBLK = (TC_BN / 32 + 1) / 2;                // 32-column blocks per warp (block c = 2 ci + cg)
    float4 bvs[NBLK];                                         // this lane's bias quad per block, requested while the MMAs still run
#pragma unroll
    for (int ci = 0; ci < NBLK; ++ci) {
      const int n = n0 + (2 * ci + cg) * 32 + c4;
      bvs[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mode == GEMM_FWD && bias) {
        if (n < N) bvs[ci].x = __ldg(bias + n);
        if (n + 1 < N) bvs[ci].y = __ldg(bias + n + 1);
        if (n + 2 < N) bvs[ci].z = __ldg(bias + n + 2);
        if (n + 3 < N) bvs[ci].w = __ldg(bias + n + 3);
      }
    }
    float bcol[NBLK];                                         // thread = row layout: lane j holds the bias of column j of the block
#pragma unroll
    for (int ci = 0; ci < NBLK; ++ci) {
      const int n = n0 + (2 * ci + cg) * 32 + lane;
      bcol[ci] = (use_tma && mode == GEMM_FWD && bias && n < N) ? __ldg(bias + n) : 0.f;
    }
    mbar_wait(accum_bar, 0);
    if (t == 0) TC_STAMP(82);
    tc_fence_after();
    const int nmain = tc_nmain(nk, TC_NMAIN);
    // All MMAs have completed (accum barrier), so the pipeline stages are free: each warp transposes its
    // 32x32 block through a padded smem scratch so that global stores / mask loads are row-contiguous.
    float* scratch = reinterpret_cast<float*>(gbase) + (warp - 2) * (32 * 33);
#pragma unroll
    for (int ci = 0; ci < NBLK; ++ci) {
      const int c = 2 * ci + cg;
      if (c >= TC_BN / 32) break;
      uint32_t v[32], w[32];
      float4 mkv[8];                                          // ReLU' mask quads of this lane's 8 output rows, in flight during the TMEM loads
      if (use_tma && mode == GEMM_DGRAD && mask) {            // thread = row: the 8 column quads of this thread's own row
        const int m_ = m0 + q * 32 + lane;
        const float* __restrict__ mp_ = mask + (long long)m_ * ldmask + n0 + c * 32;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int n_ = n0 + c * 32 + 4 * u;
          float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
          if (m_ < M) {
            if (n_ + 3 < N) {
              mk = __ldg(reinterpret_cast<const float4*>(mp_ + 4 * u));
            } else {
              if (n_ < N) mk.x = ldg_f32(mp_ + 4 * u);
              if (n_ + 1 < N) mk.y = ldg_f32(mp_ + 4 * u + 1);
              if (n_ + 2 < N) mk.z = ldg_f32(mp_ + 4 * u + 2);
              if (n_ + 3 < N) mk.w = ldg_f32(mp_ + 4 * u + 3);
            }
          }
          mkv[u] = mk;
        }
      } else if (mode == GEMM_DGRAD && mask) {
        const int n_ = n0 + c * 32 + c4;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int m_ = m0 + q * 32 + 4 * u + rsub;
          const float* __restrict__ mp_ = mask + (long long)m_ * ldmask + n_;
          float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
          if (m_ < M) {
            if (vec_m && n_ + 3 < N) {
              mk = __ldg(reinterpret_cast<const float4*>(mp_));
            } else {
              if (n_ < N) mk.x = ldg_f32(mp_);
              if (n_ + 1 < N) mk.y = ldg_f32(mp_ + 1);
              if (n_ + 2 < N) mk.z = ldg_f32(mp_ + 2);
              if (n_ + 3 < N) mk.w = ldg_f32(mp_ + 3);
            }
          }
          mkv[u] = mk;
        }
      }
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
      const uint32_t acc_stride = Cfg::kPaired ? 2u * TC_BN : (uint32_t)TC_BN;
      if (Cfg::kPaired) {                                     // [main_i | cross_i]: both halves of a pair per wait
        tc_ld32_nowait(lane_addr, v);
        tc_ld32_nowait(lane_addr + (uint32_t)TC_BN, w);
        tc_ld_wait();
        if (t == 0 && ci == 0) TC_STAMP(85);
        for (int mi = 1; mi < nmain; ++mi) {
          uint32_t v2[32], w2[32];
          tc_ld32_nowait(lane_addr + acc_stride * (uint32_t)mi, v2);
          tc_ld32_nowait(lane_addr + acc_stride * (uint32_t)mi + (uint32_t)TC_BN, w2);
          tc_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            v[jj] = __float_as_uint(__uint_as_float(v[jj]) + __uint_as_float(v2[jj]));
            w[jj] = __float_as_uint(__uint_as_float(w[jj]) + __uint_as_float(w2[jj]));     // cross terms summed apart (tiny values)
          }
        }
      } else {
        tc_ld32_nowait(lane_addr, v);                         // main[0]
        tc_ld32_nowait(lane_addr + (uint32_t)(TC_BN * TC_NMAIN), w);   // cross terms
        tc_ld_wait();
        if (t == 0 && ci == 0) TC_STAMP(85);
        for (int mi = 1; mi < nmain; ++mi) {
          uint32_t v2[32];
          tc_ld32(lane_addr + acc_stride * (uint32_t)mi, v2);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] = __float_as_uint(__uint_as_float(v[jj]) + __uint_as_float(v2[jj]));
        }
      }
      if (t == 0 && ci == 0) TC_STAMP(86);
      if (use_tma) {
        float x[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) x[jj] = __uint_as_float(v[jj]) + __uint_as_float(w[jj]);
        if (mode == GEMM_FWD) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            x[jj] += __shfl_sync(0xffffffffu, bcol[ci], jj);
            if (relu) x[jj] = fmaxf(x[jj], 0.f);
          }
        } else if (mode == GEMM_DGRAD && mask) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            x[4 * u] = mkv[u].x > 0.f ? x[4 * u] : 0.f; x[4 * u + 1] = mkv[u].y > 0.f ? x[4 * u + 1] : 0.f;
            x[4 * u + 2] = mkv[u].z > 0.f ? x[4 * u + 2] : 0.f; x[4 * u + 3] = mkv[u].w > 0.f ? x[4 * u + 3] : 0.f;
          }
        }
        // two 4-KB staging tiles per warp (the pipeline stages are free); a third block (160-wide tiles) waits for the first
        uint8_t* stg = gbase + (size_t)(warp - 2) * 8192 + (size_t)(ci & 1) * 4096;
        if (ci >= 2) { if (lane == 0) tma_store_wait_read(); __syncwarp(); }
#pragma unroll
        for (int u = 0; u < 8; ++u)                           // SWIZZLE_128B: 16-B chunk u of row r sits at chunk u ^ (r & 7)
          *reinterpret_cast<float4*>(stg + lane * 128 + ((u ^ (lane & 7)) << 4)) = make_float4(x[4 * u], x[4 * u + 1], x[4 * u + 2], x[4 * u + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (t == 0 && ci == 0) TC_STAMP(87);
        if (lane == 0 && m0 + q * 32 < M && n0 + c * 32 < N) {
          tma_store_2d(&P->tmC, smem_u32(stg), n0 + c * 32, m0 + q * 32);
          tma_store_commit();
        }
        if (t == 0 && ci == 0) TC_STAMP(88);
        continue;
      }
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) scratch[lane * 33 + jj] = __uint_as_float(v[jj]) + __uint_as_float(w[jj]);
      __syncwarp();
      if (t == 0 && ci == 0) TC_STAMP(87);
      const int n = n0 + c * 32 + c4;                         // first of this lane's four output columns
      const float4 bv = bvs[ci];
      const int mrow0 = m0 + q * 32;
      float4 xo[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {                           // all shared-memory reads of the block first ...
        const int row = 4 * u + rsub;
        xo[u].x = scratch[row * 33 + c4]; xo[u].y = scratch[row * 33 + c4 + 1];
        xo[u].z = scratch[row * 33 + c4 + 2]; xo[u].w = scratch[row * 33 + c4 + 3];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {                           // ... then the fused epilogue math ...
        float4 x = xo[u];
        if (mode == GEMM_DGRAD && mask) {
          x.x = mkv[u].x > 0.f ? x.x : 0.f; x.y = mkv[u].y > 0.f ? x.y : 0.f;
          x.z = mkv[u].z > 0.f ? x.z : 0.f; x.w = mkv[u].w > 0.f ? x.w : 0.f;
        } else if (mode == GEMM_FWD) {
          x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
          if (relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
        }
        xo[u] = x;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {                           // ... then eight back-to-back 128-bit row stores
        const int m = mrow0 + 4 * u + rsub;
        const float4 x = xo[u];
        if (m < M) {
          float* __restrict__ cp = Cout + (long long)m * ldc + n;
          if (vec_c && n + 3 < N) {
            asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
          } else {
            if (n < N) stg_f32(cp, x.x);
            if (n + 1 < N) stg_f32(cp + 1, x.y);
            if (n + 2 < N) stg_f32(cp + 2, x.z);
            if (n + 3 < N) stg_f32(cp + 3, x.w);
          }
        }
      }
      __syncwarp();
      if (t == 0 && ci == 0) TC_STAMP(88);
    }
    if (use_tma && lane == 0) tma_store_wait_read();          // the staging tiles must outlive the bulk reads
  }

  if (threadIdx.x == 64) TC_STAMP(83);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_STAMP(84);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

}  // namespace bsac
