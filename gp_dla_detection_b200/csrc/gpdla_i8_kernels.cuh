// INT8 tensor-core (tcgen05, TMEM) variant of the fused per-sample log-likelihood kernel: the Gram and the
// projection of log_mvnpdf_low_rank.m:22-28 leave the FP64 pipe.
//
// Exact-product scheme (prototype and error study: tools/ozaki_digits.py).  With a = absorption,
// d = a^2 omega^2 + v, w = a^2/d, u = a (y - a mu)/d  (process_qsos.m:192-198):
//   W'' = w cw,  cw_i = CAP (omega2_i + v_i)                      in [0, CAP]     (w is largest at a = 1)
//   U'' = u cu,  cu_i = CAP / (b_i (|y_i| + |mu_i|)),  b_i = max_{0<=a<=1} a / (a^2 omega2_i + v_i)   in [-CAP, CAP]
//   P''_ic = m_ip m_iq / cw_i 2^-e_c,   M''_ic = m_ic / cu_i 2^-e_c                  (column exponents, |.| <= CAP)
// Every factor is rounded to F = 8 L - 1 fractional bits and cut into L signed 8-bit digits; the slice pairs
// (i, j) with i + j >= L - 1 are contracted by tcgen05.mma kind::i8 into one s32 TMEM accumulator per
// diagonal i + j (exact: |sum| < 2^27), and the diagonals are recombined in FP64.  L = 6 reproduces the FP64
// Gram to 4e-15 of its scale (log-likelihoods to 8e-14), L = 5 to 9e-13 (1.4e-11).
//
// Schedule: a cluster of 4 CTAs owns 128 samples (MMA M) of one quasar.  Every CTA *produces* the W'' and U''
// digits of 32 samples per 32-pixel chunk (8 pairs of producer warps -- stage A: optical depth from the rest-frame
// table, exponential, instrument convolution; stage B: weights, digits -- the FP64 work that remains) and ships its
// 32-row block to the peers that contract it with one bulk shared-memory-to-shared-memory copy each (DSMEM,
// completion on the receiver's mbarrier).  CTAs 0..2
// contract W'' with a third of the Gram columns each (N = 80), CTA 3 contracts U'' with M'' (N = 32): TMEM
// holds L diagonals x N columns per CTA.  The P''/M'' digit chunk arrives by 1-D TMA.  One tcgen05.mma spans up
// to three consecutive digit planes / diagonals (issue_chunk_mmas_fixed: 9 instructions for the 21 slice pairs).  After
// the last chunk the accumulators are recombined, sent to the CTA that produced the sample (DSMEM stores) and
// factorised there (factor_staged, the same Cholesky as the FP64 kernels).
//
// Ranks whose pair columns exceed one cluster's TMEM (k = 40, Shape::EXT) also store their W'' digit tiles;
// gram_contract_i8_kernel contracts those with the remaining column blocks (no producers), and the accumulators of both
// kernels leave through global staging rows to cholesky_kernel (gpdla_kernels.cuh).
//
// dla_loglik_i8p_kernel keeps the clusters resident and overlaps the epilogue of a tile with the main loop of the
// next one (own warpgroup, setmaxnreg).  Measurements, failed variants and the interference probes: DESIGN.md 4.3;
// the one-tile-per-cluster development kernel of round 1 is kept as a record in profiles/experiments/.
#pragma once
#include "gpdla_kernels.cuh"

namespace gpdla {
namespace i8 {

constexpr int CLUSTER = 4;
constexpr int TS = 32;              // samples produced (and factorised) per CTA
constexpr int TM = CLUSTER * TS;    // samples per cluster = MMA M
constexpr int WCTAS = 3;            // CTAs contracting W''; the last CTA contracts U''
constexpr int NPROD = 8;            // producer warps
constexpr int SPB = TS / NPROD;     // samples per producer warp (interleaved in one instruction stream)
constexpr int NCTRL = 4;            // control / epilogue warps (MMA issue, P loader, row-block sender, spare)
constexpr int THREADS = 32 * (NCTRL + NPROD);
constexpr int STAGES = 3;           // A-operand stages
constexpr double CAP = 0.996;       // |x| <= CAP keeps the top signed digit within [-128, 127]
constexpr int TMEM_COLS = 512;
// Timing probes of the persistent kernel (compile-time, -DGPDLA_I8P_PROBE=bits; results are then wrong): 1 no
// tcgen05.mma issue (commits only), 2 no row-block copies, 4 producers skip digit stores and proxy fence, 8 epilogue
// skips the TMEM drain and the Cholesky, 16 MMAs with M = 64 instead of 128, 32 every MMA covers one digit plane only
// (same instruction count, N = 80 / 32), 64 producers skip the instrument convolution (no raw-row traffic through
// shared memory).  Measurements: DESIGN.md 4.3.
#ifndef GPDLA_I8P_PROBE
#define GPDLA_I8P_PROBE 0
#endif
constexpr int PROBE = GPDLA_I8P_PROBE;
// Barrier-wait accounting (-DGPDLA_I8P_PHASES=1 and GPDLA_I8_PHASES=1 in the environment): compiled out of the product
// build -- the null-pointer test and the clock reads around every wait cost the producers issue slots.
#ifndef GPDLA_I8P_SHFL_CONV
#define GPDLA_I8P_SHFL_CONV 0   // 1: instrument convolution through warp shuffles instead of shared-memory rows (measured 7 % SLOWER: 58.3 against 54.3 ms)
#endif
#ifndef GPDLA_I8P_PF
#define GPDLA_I8P_PF 0      // 1: stage A fetches the table cell of the next chunk a chunk ahead (measured 6 % SLOWER: 58.2 against 54.7 ms)
#endif
#ifndef GPDLA_I8P_NCA
#define GPDLA_I8P_NCA 4
#endif
#ifndef GPDLA_I8P_PHASES
#define GPDLA_I8P_PHASES 0
#endif

// Gram columns beyond what one cluster's TMEM holds (k = 40: 820 pair columns): the W'' columns are cut into WBLOCKS
// blocks of WCOLS columns (MMA N = NW = 80 each).  The producing kernel contracts blocks 0..2 (CTAs 0..2) and the U''
// columns (CTA 3) and, when there are more blocks (EXT), stores its W'' digit tiles; gram_contract_i8_kernel then
// contracts the stored tiles with blocks 3.. -- four per cluster pass, no FP64 work -- and the accumulators of all
// blocks leave through the global staging rows of the column-split FP64 path (LoglikArgs::gram) to cholesky_kernel.
template <int K, int L>
struct Shape {
  using G = GramShape<K>;
  static constexpr int WBLOCKS = (G::NPAIR <= WCTAS * 80) ? WCTAS : WCTAS + CLUSTER * ((G::NPAIR - WCTAS * 80 + CLUSTER * 80 - 1) / (CLUSTER * 80));
  static constexpr bool EXT = WBLOCKS > WCTAS;
  static constexpr int CPASSES = (WBLOCKS - WCTAS) / CLUSTER;     // cluster passes of the contract-only kernel
  static constexpr int NSLOT = WBLOCKS + 1;                       // column blocks: W'' blocks, then the U'' block
  static constexpr int USLOT = WBLOCKS;
  static constexpr int WCOLS = (G::NPAIR + WBLOCKS - 1) / WBLOCKS;   // useful Gram columns per W block (70; k = 40: 75)
  static constexpr int NW = (WCOLS + 15) / 16 * 16;               // MMA N of a W block (80)
  static constexpr int NU = (K + 15) / 16 * 16;                   // MMA N of the U block (32; k = 40: 48)
  static constexpr int NMAX = NW > NU ? NW : NU;
  static constexpr int F = 8 * L - 1;                             // fractional bits
  // A tile (128 samples x 32 pixels x L digits), MN-major (samples contiguous), no swizzle: a 16-byte granule holds one
  // pixel of 16 consecutive samples, 8 consecutive pixels make a 128-byte core matrix, pixel blocks of 8 are LBO_A apart,
  // the L digit planes of a 16-sample group lie side by side (PLANE_A apart), groups are SBO_A apart (canonical layout
  // ((1,n),(8,k)):((X,SBO),(1,LBO)) in 16-byte units, validated by bench_micro/tcgen05_i8.cu -DA_MN).  A producer lane
  // (= pixel) holds the digits of its warp's 4 consecutive samples: one 32-bit store per digit plane.
  static constexpr int LBO_A = 128;
  static constexpr int PLANE_A = (KC / 8) * LBO_A;
  static constexpr int SBO_A = PLANE_A * L;
  static constexpr int ROWBLOCK = (TS / 16) * SBO_A;              // one CTA's 32 rows, all digit planes: contiguous
  static constexpr int A_TILE = (TM / 16) * SBO_A;
  static constexpr int BW_PLANE = NW * KC, BU_PLANE = NU * KC;    // bytes of one digit plane of the B operand
  static constexpr int L_BW = L * BW_PLANE;                       // one W block's B operand of a chunk
  static constexpr int B_MAX = L * (BW_PLANE > BU_PLANE ? BW_PLANE : BU_PLANE);
  static constexpr int CHUNK_BYTES = L * (WBLOCKS * BW_PLANE + BU_PLANE);   // every column block's B operand, one chunk
  static constexpr int B_BUF = B_MAX;
  static constexpr int NCOLTAB = NSLOT * NMAX;                    // per-quasar column table [slot][NMAX]
  static constexpr int CSTR = TS + 4;
  static constexpr int NENT = (K + 1) * (K + 2) / 2;
  static_assert(L * NW <= TMEM_COLS && L * NU <= TMEM_COLS, "diagonal accumulators must fit in TMEM");
  static_assert(WBLOCKS * WCOLS >= G::NPAIR && NW <= 80, "column blocks must cover the pair columns");
  static_assert(L >= 2 && L <= 7, "digit count");
  __host__ __device__ static constexpr int b_offset(int slot) { return slot * L * BW_PLANE; }
  __host__ __device__ static constexpr int b_bytes(int slot) { return L * (slot < WBLOCKS ? BW_PLANE : BU_PLANE); }
  // column block of a CTA of the producing kernel
  __host__ __device__ static constexpr int slot_of_rank(int rank) { return rank < WCTAS ? rank : USLOT; }
  // Gram column (GramShape order: pair columns, then the projection at WT * 8) of accumulator column n of a block; -1 = padding
  __host__ __device__ static constexpr int gram_column(int slot, int n) {
    return slot < WBLOCKS ? ((n < WCOLS && slot * WCOLS + n < G::NPAIR) ? slot * WCOLS + n : -1) : (n < K ? G::WT * 8 + n : -1);
  }
  __host__ __device__ static constexpr uint64_t digit_bias() {
    uint64_t b = 0;
    for (int i = 0; i < L; ++i) b |= (uint64_t)0x80 << (8 * i);
    return b;
  }
  __host__ __device__ static constexpr double magic() {   // x + magic has ulp 2^-F:  1.5 * 2^(52 - F)
    double m = 1.5;
    for (int i = 0; i < 52 - F; ++i) m *= 2.0;
    return m;
  }
  static constexpr size_t OFF_A = 0;
  static constexpr size_t OFF_SX = OFF_A + (size_t)STAGES * A_TILE;
  static constexpr size_t OFF_B = OFF_SX + (size_t)STAGES * ROWBLOCK;
  static constexpr size_t OFF_CS = OFF_B + 2ull * B_BUF;
  static constexpr size_t OFF_RAW = OFF_CS + (EXT ? 0 : (size_t)NENT * CSTR * 8);   // EXT: no staging triangle in shared memory
  static constexpr size_t OFF_MISC = OFF_RAW + (size_t)TS * RAWS * 8;
};

// accumulator column (rank, n) -> augmented-triangle index of the staging area (-1: padding column)
__constant__ short c_i8_stage[CLUSTER * 128];

struct I8Args {
  double* pix2;            // [Q x NPIX x 2]  (cw, cu)
  double* pix8;            // [Q x 4 x NPIX] double2: the producers' per-pixel data as four planes (lambda, lh), (y, v),
                           // (mu, omega2), (cw, cu) -- a warp's 128-bit load of a plane is one contiguous 512-byte run
  uint8_t* bop;            // [Q x NPIX/KC x CHUNK_BYTES]  digit planes of P'' (ranks 0..2) and M'' (rank 3)
  double* colscale;        // [Q x 4 x NMAX]  2^(e_c - 2F + 8(L-1)): accumulator -> Gram entry
  double* colinv;          // [Q x 4 x NMAX]  2^-e_c
  int* status;             // != 0: a barrier wait timed out (kernel traps)
  int32_t* f64flag;        // [Q] set by i8_scales_kernel: 1 = a used pixel has zero noise variance, the fixed-point bound
                           // of U'' does not exist -> this quasar is left to the FP64 kernels (skipped here)
  int32_t* f64list;        // {count, q_0, q_1, ...}: the same quasars as a list (LoglikArgs::only_list of the fallback)
  uint8_t* adig;           // EXT ranks: [Q x tiles x NPIX/KC x A_TILE] W'' digit tiles written by the producing kernel
  unsigned long long* phase;   // nullable: [24] summed wait cycles per barrier (GPDLA_I8_PHASES diagnostics)
};

__device__ __forceinline__ unsigned long long* phase_ptr(const I8Args& xa) { return GPDLA_I8P_PHASES ? xa.phase : nullptr; }

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// wait with a deadline: a protocol error becomes a trapped kernel (CUDA error), not a hung GPU
__device__ __forceinline__ void mbar_wait_d(uint64_t* bar, uint32_t parity, int* status, int code,
                                            unsigned long long* phase = nullptr, unsigned backoff_ns = 0) {
  const uint32_t a = smem_u32(bar);
  const long long t0 = phase ? clock64() : 0;
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) break;
    if (backoff_ns) __nanosleep(backoff_ns);
    if (spins > 20000000u) {   // seconds: a protocol error, not a slow peer
      if (status) atomicExch(status, code + 100 * (int)(cluster_ctarank() + 1));
      __threadfence_system();
      asm volatile("trap;");
    }
  }
  if (phase) atomicAdd(&phase[8 + code], (unsigned long long)(clock64() - t0));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { mbar_expect_tx(bar, bytes); }
// shared::cta -> peer CTA's shared memory, completion (bytes) on the peer's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void st_cluster_f64(uint32_t addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// no-swizzle shared-memory matrix descriptor (K-major B operand and MN-major A operand validated by bench_micro/tcgen05_i8.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}
// kind::i8 instruction descriptor: D = s32, A = B = signed 8-bit, A MN-major (bit 15), B K-major, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// shared::cta -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_store_global(void* dst, uint32_t src_cta, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_cta), "r"(bytes) : "memory");
}
// global -> the same shared-memory offset of every CTA in `mask`, completion (bytes) on each destination's mbarrier
__device__ __forceinline__ void tma_load_1d_multicast(uint32_t dst_cta, const void* src, uint32_t bytes, uint32_t bar_cta, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst_cta),
               "l"(src), "r"(bytes), "r"(bar_cta), "h"(mask)
               : "memory");
}

// The L (L + 1) / 2 slice-pair products of one chunk.  Digit plane i of A pairs with the B planes j = L-1-i .. L-1,
// which are consecutive in shared memory, into the diagonals t = i + j - (L-1) = 0 .. i, which are consecutive in
// TMEM: up to 256 / n accumulator blocks are therefore covered by ONE tcgen05.mma whose N spans several planes.
// 6 digits: 9 instructions instead of 21 (5 digits: 7 instead of 15).  Every descriptor offset is an immediate (N is
// a template parameter): the issuing thread shares an SM sub-partition with two producer warps, and each instruction
// it spends on descriptor arithmetic is an issue slot taken from them (the cost of an MMA does not
// depend on its size).  da0 / db0 are the descriptors of digit plane 0 of the A stage / B buffer.
template <class Sh, int L, int N>
__device__ __forceinline__ void issue_chunk_mmas_fixed(uint32_t tmem_base, uint64_t da0, uint64_t db0, bool accumulate) {
  constexpr int group = 256 / N;
#pragma unroll
  for (int i = L - 1; i >= 0; --i) {
    const uint64_t da = da0 + (uint64_t)((i * Sh::PLANE_A) >> 4);
#pragma unroll
    for (int t0 = 0; t0 <= i; t0 += group) {
      const int nt = (PROBE & 32) ? 1 : ((group < i + 1 - t0) ? group : i + 1 - t0);
      const uint64_t db = db0 + (uint64_t)(((L - 1 - i + t0) * N * KC) >> 4);
      const uint32_t idesc = (PROBE & 16) ? (make_idesc(nt * N) & ~(0x1fu << 24)) | (4u << 24) : make_idesc(nt * N);
      mma_i8(tmem_base + (uint32_t)(t0 * N), da, db, idesc, (accumulate || i < L - 1) ? 1u : 0u);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K0c: per-pixel scales (cw, cu) and per-column exponents of the digit operands.  One CTA per quasar.
template <int K, int L>
__global__ void __launch_bounds__(NTHREADS) i8_scales_kernel(const QuasarMeta* __restrict__ meta, const double* __restrict__ pix,
                                                             const double* __restrict__ Mq, const double* __restrict__ lam_pad,
                                                             const double* __restrict__ lamh, I8Args xa, int NPIX) {
  using Sh = Shape<K, L>;
  using G = GramShape<K>;
  const int q = blockIdx.x, tid = threadIdx.x;
  const double* pq = pix + (int64_t)q * NPIX * 4;
  double* p2 = xa.pix2 + (int64_t)q * NPIX * 2;
  double2* p8 = reinterpret_cast<double2*>(xa.pix8 + (int64_t)q * NPIX * 8);
  const double* lq = lam_pad + (int64_t)q * (NPIX + 8) + 6;     // pixel i sits at position i + 6 of the padded grid
  const double* lhq = lamh + (int64_t)q * (NPIX + 8) + 6;
  bool unbounded = false;
  for (int i = tid; i < NPIX; i += NTHREADS) {
    const double y = pq[i * 4 + 0], v = pq[i * 4 + 1], mu = pq[i * 4 + 2], om2 = pq[i * 4 + 3];
    const double cw = CAP * (om2 + v);
    // b = max over 0 <= a <= 1 of a / (a^2 om2 + v): at a = sqrt(v / om2) if that is < 1, else at a = 1
    const double b = (v >= om2) ? 1.0 / (om2 + v) : 0.5 / sqrt(om2 * v);
    const double yy = fabs(y) + fabs(mu);
    // v = 0 (masked and padding pixels carry v = 1): u = a (y - a mu) / (a^2 om2) has no bound as a -> 0, and the
    // reference divides by zero once the profile saturates (process_qsos.m:194-198, log_mvnpdf_low_rank.m:13)
    unbounded |= !(v > 0.0) || !isfinite(b) || !isfinite(cw);
    const double cu = (yy > 0.0 && isfinite(b)) ? CAP / (b * yy) : 0.0;
    p2[i * 2 + 0] = cw; p2[i * 2 + 1] = cu;
    p8[i] = make_double2(lq[i], lhq[i]); p8[NPIX + i] = make_double2(y, v);
    p8[2 * NPIX + i] = make_double2(mu, om2); p8[3 * NPIX + i] = make_double2(cw, cu);
  }
  if (__syncthreads_or(unbounded) && tid == 0 && meta[q].nchunks > 0) {
    xa.f64flag[q] = 1;
    xa.f64list[1 + atomicAdd(&xa.f64list[0], 1)] = q;
  }
  // column maxima of |P''| and |M''| over the pixels, tiled through shared memory (M rows and 1/cw, 1/cu of 32 pixels)
  const int n_u = meta[q].n_u;
  const double* mq = Mq + (int64_t)q * NPIX * K;
  __shared__ double sM[KC][K + 1];
  __shared__ double sr[KC][2];
  constexpr int EPT = (Sh::NCOLTAB + NTHREADS - 1) / NTHREADS;   // column table entries per thread
  int cp[EPT], cq[EPT], crank[EPT];
  double mx2[EPT];
#pragma unroll
  for (int m = 0; m < EPT; ++m) {
    cp[m] = -1; cq[m] = 0; crank[m] = 0; mx2[m] = 0.0;
    const int idx = tid + m * NTHREADS;
    if (idx >= Sh::NCOLTAB) continue;
    const int rank = idx / Sh::NMAX, n = idx % Sh::NMAX;   // rank = column block: W'' blocks, then the U'' block
    crank[m] = rank;
    if (rank < Sh::WBLOCKS) {
      const int c = rank * Sh::WCOLS + n;
      if (n < Sh::WCOLS && c < G::NPAIR) {
        int p = 0;
        while (p + 1 < K && G::pair_index(p + 1, p + 1) <= c) ++p;
        cp[m] = p; cq[m] = p + (c - G::pair_index(p, p));
      }
    } else if (n < K) {
      cp[m] = n;
    }
  }
  for (int i0 = 0; i0 < n_u; i0 += KC) {
    for (int t = tid; t < KC * K; t += NTHREADS) sM[t / K][t % K] = mq[(int64_t)i0 * K + t];
    if (tid < KC * 2) {
      const double c = p2[(i0 + tid / 2) * 2 + (tid & 1)];
      sr[tid / 2][tid & 1] = c > 0.0 ? 1.0 / c : 0.0;       // the operand builder multiplies by the same reciprocals
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < EPT; ++m) {
      if (cp[m] < 0) continue;
      for (int r = 0; r < KC; ++r) {
        const double x = (crank[m] < Sh::WBLOCKS) ? __dmul_rn(__dmul_rn(sM[r][cp[m]], sM[r][cq[m]]), sr[r][0])
                                            : __dmul_rn(sM[r][cp[m]], sr[r][1]);
        mx2[m] = fmax(mx2[m], fabs(x));
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < EPT; ++m) {
    const int idx = tid + m * NTHREADS;
    if (idx >= Sh::NCOLTAB) continue;
    const double mx = mx2[m];
    int e = 0;
    if (mx > 0.0 && isfinite(mx)) {
      int ex;
      const double mant = frexp(mx, &ex);     // mx = mant 2^ex, mant in [0.5, 1)
      e = (mant <= CAP) ? ex : ex + 1;
    }
    xa.colscale[(int64_t)q * Sh::NCOLTAB + idx] = ldexp(1.0, e - 2 * Sh::F + 8 * (L - 1));
    xa.colinv[(int64_t)q * Sh::NCOLTAB + idx] = ldexp(1.0, -e);
  }
}

// K0d: digit planes of P'' and M'' in the tcgen05 K-major core-matrix layout, [Q][chunk][column block][digit][N x 32 B].
template <int K, int L>
__global__ void __launch_bounds__(NTHREADS) i8_build_operand_kernel(const QuasarMeta* __restrict__ meta,
                                                                    const double* __restrict__ Mq, I8Args xa, int NPIX) {
  using Sh = Shape<K, L>;
  using G = GramShape<K>;
  const int q = blockIdx.y, chunk = blockIdx.x;
  if (chunk >= meta[q].nchunks) return;
  __shared__ double sM[KC][K + 1];
  __shared__ double sc[KC][2];
  const double* src = Mq + ((int64_t)q * NPIX + (int64_t)chunk * KC) * K;
  for (int t = threadIdx.x; t < KC * K; t += NTHREADS) sM[t / K][t % K] = src[t];
  for (int t = threadIdx.x; t < KC * 2; t += NTHREADS) {   // 1/cw, 1/cu: the same reciprocals as in i8_scales_kernel
    const double c = xa.pix2[((int64_t)q * NPIX + (int64_t)chunk * KC) * 2 + t];
    sc[t / 2][t % 2] = c > 0.0 ? 1.0 / c : 0.0;
  }
  __syncthreads();
  uint8_t* dst = xa.bop + ((int64_t)q * (NPIX / KC) + chunk) * Sh::CHUNK_BYTES;
  const double* cinv = xa.colinv + (int64_t)q * Sh::NCOLTAB;
  constexpr int ROWS = Sh::WBLOCKS * Sh::NW + Sh::NU;
  for (int t = threadIdx.x; t < ROWS * KC; t += NTHREADS) {
    const int k = t % KC, row = t / KC;
    const int rank = row < Sh::WBLOCKS * Sh::NW ? row / Sh::NW : Sh::WBLOCKS;   // column block
    const int n = row - rank * Sh::NW;
    const int N = rank < Sh::WBLOCKS ? Sh::NW : Sh::NU;
    double x = 0.0;
    if (rank < Sh::WBLOCKS) {
      const int c = rank * Sh::WCOLS + n;
      if (n < Sh::WCOLS && c < G::NPAIR) {
        int p = 0;
        while (p + 1 < K && G::pair_index(p + 1, p + 1) <= c) ++p;
        const int qq = p + (c - G::pair_index(p, p));
        x = __dmul_rn(__dmul_rn(sM[k][p], sM[k][qq]), sc[k][0]);
      }
    } else if (n < K) {
      x = __dmul_rn(sM[k][n], sc[k][1]);
    }
    x = x * cinv[rank * Sh::NMAX + n];                       // exact: power of two
    const long long X = __double2ll_rn(ldexp(x, Sh::F));     // |X| <= CAP 2^F
    const uint64_t xb = ((uint64_t)X + Sh::digit_bias()) ^ Sh::digit_bias();
    uint8_t* d0 = dst + Sh::b_offset(rank) + (n / 8) * 256 + (k / 16) * 128 + (n % 8) * 16 + (k % 16);
#pragma unroll
    for (int j = 0; j < L; ++j) d0[(size_t)j * N * KC] = (uint8_t)(xb >> (8 * j));
  }
}

// ------------------------------------------------------------------------------------------
// Persistent variant (shipped).  The grid holds as many 4-CTA clusters as the GPU can keep resident; each cluster
// walks the (quasar, 128-sample tile) list with stride = number of clusters.  A CTA has 24 warps in six
// warpgroups with their own register budgets (setmaxnreg): control (MMA issuer, B loader, row-block sender; 40
// registers), two producer warpgroups of stage A (96: optical depth, exponential, convolution), two of stage B (80:
// weights, digit planes) and an EPILOGUE warpgroup (88) that recombines the TMEM accumulators and runs the Cholesky
// of tile t while the producers and the tensor pipe are already working on tile t + 1.
// Hand-overs: accumulators final (tcgen05.commit -> bar_acc), TMEM drained (bar_tfree), staging triangle delivered
// (remote arrivals on the owner's bar_csfull) and free again (remote arrivals on every writer's bar_csfree[owner]),
// per-sample scalars (bar_sq, double-buffered).  All main-loop barriers run on a chunk counter that spans tiles.
constexpr int P_THREADS = 32 * (NCTRL + 2 * NPROD + 4);
constexpr int REG_CTRL = 40, REG_A = 96, REG_B = 80, REG_EPI = 88;
constexpr int REG_BUDGET_TOTAL = REG_CTRL * 128 + (REG_A + REG_B) * 256 + REG_EPI * 128;
// 80 = the launch allocation ptxas gives a 768-thread CTA (65536 / 768 rounded down to a multiple of 8): setmaxnreg.inc can
// only draw on what the CTA owned at launch, not on the SM's spare registers (bench_micro/setmaxnreg_pool.cu; the host
// checks cudaFuncAttributes::numRegs against REG_BUDGET_TOTAL before the first launch)
static_assert(REG_BUDGET_TOTAL <= 80 * P_THREADS, "register budgets exceed the CTA's launch allocation");

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
// wait with cluster-scope acquire (the data was written by other CTAs of the cluster)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* status, int code,
                                                  unsigned long long* phase = nullptr) {
  const uint32_t a = smem_u32(bar);
  const long long t0 = phase ? clock64() : 0;
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) break;
    if (spins > 20000000u) {
      if (status) atomicExch(status, code + 100 * (int)(cluster_ctarank() + 1));
      __threadfence_system();
      asm volatile("trap;");
    }
  }
  if (phase) atomicAdd(&phase[8 + code], (unsigned long long)(clock64() - t0));
}

// The digits of a producer lane's 4 consecutive samples (x[ss] holds the L signed digits of sample ss in its low
// bytes) as one 32-bit word per digit plane -- byte b = sample b, the MN-major A layout -- by 4 x 4 byte transposes
// (8 PRMT for planes 0..3, 4 for planes 4..5).
template <int L, int PLANE>
__device__ __forceinline__ void store_digit_planes(uint8_t* dst, const uint64_t (&x)[SPB]) {
  static_assert(SPB == 4, "one 32-bit word = 4 samples");
  const uint32_t a0 = (uint32_t)x[0], a1 = (uint32_t)x[1], a2 = (uint32_t)x[2], a3 = (uint32_t)x[3];
  const uint32_t t0 = __byte_perm(a0, a1, 0x5140), t1 = __byte_perm(a2, a3, 0x5140);
  const uint32_t t2 = __byte_perm(a0, a1, 0x7362), t3 = __byte_perm(a2, a3, 0x7362);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  d[0] = __byte_perm(t0, t1, 0x5410);
  d[PLANE / 4] = __byte_perm(t0, t1, 0x7632);
  if (L > 2) d[2 * (PLANE / 4)] = __byte_perm(t2, t3, 0x5410);
  if (L > 3) d[3 * (PLANE / 4)] = __byte_perm(t2, t3, 0x7632);
  if (L > 4) {
    const uint32_t h0 = (uint32_t)(x[0] >> 32), h1 = (uint32_t)(x[1] >> 32), h2 = (uint32_t)(x[2] >> 32), h3 = (uint32_t)(x[3] >> 32);
    const uint32_t u0 = __byte_perm(h0, h1, 0x5140), u1 = __byte_perm(h2, h3, 0x5140);
    d[4 * (PLANE / 4)] = __byte_perm(u0, u1, 0x5410);
    if (L > 5) d[5 * (PLANE / 4)] = __byte_perm(u0, u1, 0x7632);
    if (L > 6) {
      const uint32_t u2 = __byte_perm(h0, h1, 0x7362), u3 = __byte_perm(h2, h3, 0x7362);
      d[6 * (PLANE / 4)] = __byte_perm(u2, u3, 0x5410);
    }
  }
}

// One warp's quarter of the TMEM accumulators (32 rows = lanes, the L diagonals of N columns each) recombined in FP64
// and written to the global staging rows: `grow` = this lane's row, columns in GramShape order.
template <class Sh, int L>
__device__ __forceinline__ void drain_to_gram(uint32_t taddr0, int N, int slot, const double* cs, double* grow, int c_begin = 0,
                                              int c_end = 1 << 30) {
  for (int c0 = c_begin; c0 < min(N, c_end); c0 += 8) {
    uint32_t v[L][8];
#pragma unroll
    for (int tt = 0; tt < L; ++tt) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[tt][0]), "=r"(v[tt][1]), "=r"(v[tt][2]), "=r"(v[tt][3]), "=r"(v[tt][4]), "=r"(v[tt][5]),
                     "=r"(v[tt][6]), "=r"(v[tt][7])
                   : "r"(taddr0 + (uint32_t)(tt * N + c0)));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int col = Sh::gram_column(slot, c0 + jj);
      if (col >= 0) {
        double acc = (double)(int32_t)v[L - 1][jj];
#pragma unroll
        for (int tt = L - 2; tt >= 0; --tt) acc = fma(acc, 256.0, (double)(int32_t)v[tt][jj]);
        grow[col] = acc * cs[c0 + jj];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Contract-only passes of the EXT ranks (k = 40).  The W'' digit tiles the producing kernel stored are contracted with
// the column blocks it did not cover: a cluster of 4 CTAs takes one (cluster pass, quasar, 128-sample tile) at a time, CTA r
// owning block WCTAS + 4 (pass) + r.  Per 32-pixel chunk every CTA fetches a quarter of the A tile and multicasts it to
// the four CTAs (one read of the tile from L2 / HBM per cluster) and fetches its own B block; no FP64 work, no producers:
// the tensor pipe is the only busy unit.  Warp 0: MMA issuer, warp 1: loader, warps 4..11: epilogue (two warps per TMEM
// lane quarter, half of the columns each: the tensor pipe idles while the single accumulator buffer is drained).
constexpr int C_THREADS = 384;
constexpr int C_EPI_WARPS = 8;
constexpr int C_STAGES = 4;
template <int K, int L>
struct CShape {
  using Sh = Shape<K, L>;
  static constexpr size_t OFF_A = 0;
  static constexpr size_t OFF_B = OFF_A + (size_t)C_STAGES * Sh::A_TILE;
  static constexpr size_t OFF_BAR = OFF_B + (size_t)C_STAGES * Sh::L_BW;
  static constexpr size_t SMEM = OFF_BAR + 32 * 8 + 16;
};

template <int K, int L>
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(C_THREADS, 1)
gram_contract_i8_kernel(LoglikArgs args, I8Args xa, int num_quasars, int tiles_per_quasar) {
  using Sh = Shape<K, L>;
  using CS = CShape<K, L>;
  static_assert(Sh::EXT, "ranks whose pair columns fit one cluster have no contract-only passes");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / CLUSTER, num_clusters = gridDim.x / CLUSTER;
  const int tiles_per_pass = num_quasars * tiles_per_quasar;
  const int num_tiles = tiles_per_pass * Sh::CPASSES;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint8_t* At = smem_raw + CS::OFF_A;
  uint8_t* Bt = smem_raw + CS::OFF_B;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + CS::OFF_BAR);   // [C_STAGES]
  uint64_t* bar_empty = bar_full + C_STAGES;                                   // [C_STAGES]
  uint64_t* bar_acc = bar_empty + C_STAGES;
  uint64_t* bar_tfree = bar_acc + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_full + 24);
  if (tid == 0) {
    for (int i = 0; i < C_STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], CLUSTER); }
    mbar_init(bar_acc, 1); mbar_init(bar_tfree, C_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = *s_tmem;
  cluster_sync_all();

  auto tile_live = [&](int q, const QuasarMeta& m) {
    return m.nchunks > 0 && !(args.active != nullptr && args.active[q] == 0) && xa.f64flag[q] == 0;
  };
  constexpr uint32_t b_bytes = (uint32_t)Sh::L_BW;
  if (warp == 0 && lane == 0) {
    // ---- MMA issuer
    const uint64_t da_stage0 = make_desc(smem_u32(At), Sh::LBO_A, Sh::SBO_A);
    const uint64_t db_buf0 = make_desc(smem_u32(Bt), 128, 256);
    int gc = 0, it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int q = (t % tiles_per_pass) / tiles_per_quasar;
      const QuasarMeta meta = args.meta[q];
      if (!tile_live(q, meta)) continue;
      if (it > 0) mbar_wait_d(bar_tfree, (it - 1) & 1, xa.status, 23, nullptr, 100);
      for (int c = 0; c < meta.nchunks; ++c, ++gc) {
        const int stage = gc % C_STAGES;
        mbar_wait_d(&bar_full[stage], (gc / C_STAGES) & 1, xa.status, 22, nullptr);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint64_t da0 = da_stage0 + (uint64_t)(uint32_t)(stage * (Sh::A_TILE >> 4));
        const uint64_t db0 = db_buf0 + (uint64_t)(uint32_t)(stage * (Sh::L_BW >> 4));
        issue_chunk_mmas_fixed<Sh, L, Sh::NW>(tmem_base, da0, db0, c > 0);
        mma_commit_multicast(&bar_empty[stage], (uint16_t)((1u << CLUSTER) - 1));
      }
      mma_commit(bar_acc);
      ++it;
    }
  } else if (warp == 1 && lane == 0) {
    // ---- loader: a quarter of the A tile to all four CTAs, this CTA's B block to itself
    int gc = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int pass = t / tiles_per_pass, qt = t % tiles_per_pass;
      const int q = qt / tiles_per_quasar;
      const QuasarMeta meta = args.meta[q];
      if (!tile_live(q, meta)) continue;
      const int slot = WCTAS + CLUSTER * pass + (int)rank;
      const uint8_t* asrc = xa.adig + (int64_t)qt * (args.NPIX / KC) * Sh::A_TILE + rank * Sh::ROWBLOCK;
      const uint8_t* bsrc = xa.bop + (int64_t)q * (args.NPIX / KC) * Sh::CHUNK_BYTES + Sh::b_offset(slot);
      for (int c = 0; c < meta.nchunks; ++c, ++gc) {
        const int stage = gc % C_STAGES;
        // the stage is free in ALL four CTAs (every CTA's MMAs of chunk gc - C_STAGES have committed to every peer)
        mbar_wait_d(&bar_empty[stage], ((gc / C_STAGES) & 1) ^ 1, xa.status, 21, nullptr, 50);
        mbar_expect_tx(&bar_full[stage], (uint32_t)Sh::A_TILE + b_bytes);
        tma_load_1d_multicast(smem_u32(At + stage * Sh::A_TILE) + rank * Sh::ROWBLOCK, asrc + (int64_t)c * Sh::A_TILE, Sh::ROWBLOCK,
                              smem_u32(&bar_full[stage]), (uint16_t)((1u << CLUSTER) - 1));
        tma_load_1d(Bt + stage * Sh::L_BW, bsrc + (int64_t)c * Sh::CHUNK_BYTES, b_bytes, &bar_full[stage]);
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: TMEM lane quarter e = rows 32 e .. 32 e + 31 of the tile (a warp may only read the quarter warp % 4)
    const int e = warp & 3, half = (warp - 4) >> 2;
    constexpr int CHALF = (Sh::NW / 2 + 7) / 8 * 8;
    const uint32_t taddr0 = tmem_base + ((uint32_t)(e * 32) << 16);
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int pass = t / tiles_per_pass, qt = t % tiles_per_pass;
      const int q = qt / tiles_per_quasar;
      const QuasarMeta meta = args.meta[q];
      if (!tile_live(q, meta)) continue;
      const int slot = WCTAS + CLUSTER * pass + (int)rank;
      if (lane == 0) mbar_wait_d(bar_acc, it & 1, xa.status, 24, nullptr, 200);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;");
      const int64_t row = (int64_t)(qt % tiles_per_quasar) * TM + e * 32 + lane;
      drain_to_gram<Sh, L>(taddr0, Sh::NW, slot, xa.colscale + (int64_t)q * Sh::NCOLTAB + slot * Sh::NMAX,
                           args.gram + ((int64_t)q * args.gram_rows + row) * Sh::G::NCOL, half * CHALF, (half + 1) * CHALF);
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tfree);
      ++it;
    }
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

template <int K, int L>
struct PShape {
  using Sh = Shape<K, L>;
  static constexpr size_t OFF_MISC = Sh::OFF_MISC;
  __host__ __device__ static constexpr size_t smem_bytes(int num_lines) {
    // per-sample arrays: nhi, q[2], ld[2], mult[num_lines + 1]; 128 barriers; partner indices; the stage A -> B hand-over
    // buffers (2 slots x 4 samples x 32 pixels per warp pair)
    return OFF_MISC + (size_t)TS * (num_lines + 6) * 8 + 128 * 8 + 4 * TS * 4 + 64 + (size_t)NPROD * 2 * SPB * KC * 8;
  }
};

template <int K, int L, int NL, int MODE>
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(P_THREADS, 1)
dla_loglik_i8p_kernel(LoglikArgs args, I8Args xa, int num_quasars, int tiles_per_quasar) {
  using Sh = Shape<K, L>;
  constexpr int CSTR = Sh::CSTR;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t S = args.S;
  const int cluster_id = blockIdx.x / CLUSTER, num_clusters = gridDim.x / CLUSTER;
  const int num_tiles = num_quasars * tiles_per_quasar;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint8_t* At = smem_raw + Sh::OFF_A;
  uint8_t* Sx = smem_raw + Sh::OFF_SX;
  uint8_t* Bt = smem_raw + Sh::OFF_B;
  double* Cs = reinterpret_cast<double*>(smem_raw + Sh::OFF_CS);
  double* rawbuf = reinterpret_cast<double*>(smem_raw + Sh::OFF_RAW);
  double* s_nhi = reinterpret_cast<double*>(smem_raw + Sh::OFF_MISC);      // [TS]
  double* s_q = s_nhi + TS;                                                // [2][TS]  sum r^2/d, by tile parity
  double* s_ld = s_q + 2 * TS;                                             // [2][TS]  sum log d
  double* s_mult = s_ld + 2 * TS;                                          // [num_lines][TS]
  const int num_lines = (NL > 0) ? NL : args.num_lines;
  double* s_K = s_mult + (size_t)TS * num_lines;                           // [TS]  rest-frame table offsets
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_K + TS);
  uint64_t* bar_full = bars;                    // [STAGES]
  uint64_t* bar_empty = bar_full + STAGES;      // [STAGES]
  uint64_t* bar_rows = bar_empty + STAGES;      // [STAGES]
  uint64_t* bar_pfull = bar_rows + STAGES;      // [2]
  uint64_t* bar_pempty = bar_pfull + 2;         // [2]
  uint64_t* bar_acc = bar_pempty + 2;           // accumulators of the tile final
  uint64_t* bar_tfree = bar_acc + 1;            // TMEM drained by the epilogue warps
  uint64_t* bar_sq = bar_tfree + 1;             // [2] per-sample scalars of the tile written
  uint64_t* bar_csfull = bar_sq + 2;            // this CTA's staging triangle is complete
  uint64_t* bar_csfree = bar_csfull + 1;        // [CLUSTER] CTA w has finished factorising: its triangle may be rewritten
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 32);
  uint64_t* bar_hfull = bars + 64;              // [NPROD][2] absorption of a chunk handed over by the pair's stage-A warp
  uint64_t* bar_hempty = bar_hfull + 2 * NPROD; // [NPROD][2] ... consumed by its stage-B warp
  int* s_part = reinterpret_cast<int*>(bars + 128);                        // [3][TS]
  int* s_so = s_part + 3 * TS;                                             // [TS]  sample index of every tile row
  double* hbuf = reinterpret_cast<double*>(s_so + TS + 16);                // [NPROD][2][SPB][KC]

  if (tid == 0) {
    // EXT: a stage is free once the four MMAs have read it AND the sender's store of its W'' block to global memory has
    for (int i = 0; i < STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], CLUSTER + (Sh::EXT ? 1 : 0)); mbar_init(&bar_rows[i], NPROD); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_pfull[i], 1); mbar_init(&bar_pempty[i], 1); mbar_init(&bar_sq[i], NPROD); }
    mbar_init(bar_acc, 1); mbar_init(bar_tfree, 4); mbar_init(bar_csfull, CLUSTER);
    for (int i = 0; i < 2 * NPROD; ++i) { mbar_init(&bar_hfull[i], 1); mbar_init(&bar_hempty[i], 1); }
    for (int i = 0; i < CLUSTER; ++i) mbar_init(&bar_csfree[i], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = *s_tmem;
  cluster_sync_all();

  const int N = rank < WCTAS ? Sh::NW : Sh::NU;
  // where the instrument convolution runs: in stage B for the single-DLA pass (stage A is the longer one), in stage A
  // when the convolved absorption is also cached for the later levels (MODE 1); MODE 2 has none
  // (NCA of a pair's 4 samples are convolved by stage A, the others by stage B: the split that balances the two stages)
  constexpr int NCA = (MODE == 0) ? GPDLA_I8P_NCA : SPB;
  constexpr bool CONV_B = NCA < SPB;
  constexpr bool SHFL_CONV = GPDLA_I8P_SHFL_CONV && !CONV_B && MODE != 2;
  const int slot = Sh::slot_of_rank((int)rank);
  const uint32_t b_bytes = (uint32_t)Sh::b_bytes(slot);
  // a tile is skipped by every role alike when its quasar has no usable pixel or is inactive
  auto tile_quasar = [&](int t) { return t / tiles_per_quasar; };
  auto tile_live = [&](int q, const QuasarMeta& m) {
    return m.nchunks > 0 && !(args.active != nullptr && args.active[q] == 0) && xa.f64flag[q] == 0;
  };

  if (warp >= NCTRL + 2 * NPROD) {
    // =========================================================================== EPILOGUE WARPGROUP
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_EPI));
    const int e = warp - NCTRL - 2 * NPROD;          // TMEM lane quarter = samples 32 e .. 32 e + 31 = CTA e's samples
    const uint32_t cs_remote = mapa(smem_u32(Cs), (uint32_t)e) + (uint32_t)lane * 8;
    const uint32_t csfull_remote = mapa(smem_u32(bar_csfull), (uint32_t)e);
    const uint32_t taddr0 = tmem_base + ((uint32_t)(e * 32) << 16);
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int q = tile_quasar(t);
      const QuasarMeta meta = args.meta[q];
      const int64_t s0 = (int64_t)(t % tiles_per_quasar) * TM + (int64_t)rank * TS;
      if (!tile_live(q, meta)) {
        if (lane < 8 && xa.f64flag[q] == 0) {   // dead quasar: NaN results; a flagged one belongs to the FP64 kernels
          const int64_t s = s0 + e * 8 + lane;
          if (s < S) args.sample_log_likelihoods[(int64_t)q * args.sll_stride + s] = NAN;
          else if (s == S && args.log_likelihoods_no_dla) args.log_likelihoods_no_dla[q] = NAN;
        }
        continue;
      }
      if (lane == 0) {
        mbar_wait_d(bar_acc, it & 1, xa.status, 6, phase_ptr(xa), 500);   // latency-insensitive: poll rarely
        if (!Sh::EXT && it > 0) mbar_wait_cluster(&bar_csfree[e], (it - 1) & 1, xa.status, 10, phase_ptr(xa));   // CTA e is done with its previous triangle
      }
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;");
      const double* cs = xa.colscale + (int64_t)q * Sh::NCOLTAB + slot * Sh::NMAX;
      if constexpr (Sh::EXT) {
        // accumulators of this CTA's column block, rows of CTA e's samples -> global staging rows (cholesky_kernel)
        drain_to_gram<Sh, L>(taddr0, N, slot, cs, args.gram + ((int64_t)q * args.gram_rows + (s0 - (int64_t)rank * TS) + e * 32 + lane) * Sh::G::NCOL);
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tfree);
        ++it;
        continue;
      }
      for (int c0 = 0; c0 < ((PROBE & 8) ? 0 : N); c0 += 8) {
        uint32_t v[L][8];
#pragma unroll
        for (int tt = 0; tt < L; ++tt) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(v[tt][0]), "=r"(v[tt][1]), "=r"(v[tt][2]), "=r"(v[tt][3]), "=r"(v[tt][4]), "=r"(v[tt][5]),
                         "=r"(v[tt][6]), "=r"(v[tt][7])
                       : "r"(taddr0 + (uint32_t)(tt * N + c0)));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int idx = c_i8_stage[rank * 128 + c0 + jj];
          if (idx >= 0) {
            double acc = (double)(int32_t)v[L - 1][jj];
#pragma unroll
            for (int tt = L - 2; tt >= 0; --tt) acc = fma(acc, 256.0, (double)(int32_t)v[tt][jj]);
            st_cluster_f64(cs_remote + (uint32_t)(idx * CSTR * 8), acc * cs[c0 + jj]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      asm volatile("fence.acq_rel.cluster;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_tfree);                      // the MMA issuer may overwrite the accumulators
        mbar_arrive_cluster(csfull_remote);          // this CTA's share of CTA e's triangle is in place
        mbar_wait_cluster(bar_csfull, it & 1, xa.status, 11, phase_ptr(xa));                 // all four shares of my triangle
        mbar_wait_d(&bar_sq[it & 1], (it >> 1) & 1, xa.status, 12, phase_ptr(xa), 200);            // my producers' scalars
      }
      __syncwarp();
      if constexpr (!Sh::EXT) {
        if (!(PROBE & 8)) factor_staged<K, CSTR>(Cs, s_q + (it & 1) * TS, s_ld + (it & 1) * TS, e * 8, lane, meta, args, q, s0);
      }
      __syncwarp();
      if (lane == 0) {
        const uint32_t freebar = smem_u32(&bar_csfree[rank]);
        for (uint32_t peer = 0; peer < (uint32_t)CLUSTER; ++peer) mbar_arrive_cluster(mapa(freebar, peer));
      }
      ++it;
    }
  } else if (warp >= NCTRL + NPROD) {
    // =========================================================================== PRODUCERS, STAGE B: weights and digits
    // The producers are a two-stage pipeline of warp pairs (A: absorption, B: weights and digits).  One warp doing both
    // (round 1 .. mid round 2, 8 warps at 168 registers) spent half of its time on dependency and load latency that two
    // warps per sub-partition could not cover; the split doubles the warps in flight at the same register total, with
    // the absorption of a warp's 4 samples x 32 pixels handed over through a 2-slot shared-memory buffer per pair.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_B));
    const int pr = warp - NCTRL - NPROD;
    const int row0 = pr * SPB;
    const uint32_t own_block = rank * Sh::ROWBLOCK;
    uint8_t* const wdst0 = (rank < WCTAS) ? At + own_block : Sx;
    uint8_t* const udst0 = (rank < WCTAS) ? Sx : At + own_block;
    const uint32_t wstride = (rank < WCTAS) ? Sh::A_TILE : Sh::ROWBLOCK;
    const uint32_t ustride = (rank < WCTAS) ? Sh::ROWBLOCK : Sh::A_TILE;
    const uint32_t rowoff = (row0 / 16) * Sh::SBO_A + (row0 % 16) + (lane / 8) * Sh::LBO_A + (lane % 8) * 16;
    constexpr uint64_t BIAS = Sh::digit_bias();
    const double MAGIC = Sh::magic();
    const uint64_t KADD = BIAS - (uint64_t)__double_as_longlong(MAGIC);
    const double* hb = hbuf + pr * (2 * SPB * KC) + lane;
    uint64_t* const hfull = bar_hfull + 2 * pr;
    uint64_t* const hempty = bar_hempty + 2 * pr;
    double* myraw = rawbuf + row0 * RAWS;             // CONV_B: this warp's raw-profile rows
    int it = 0;                                       // live-tile counter
    int stage = 0;                                    // = gc % STAGES (gc: chunk counter across tiles)
    uint32_t empty_parity = 1;                        // = ((gc / STAGES) & 1) ^ 1
    uint32_t gh = 0;                                  // hand-over counter across tiles
    long long w_empty = 0, w_hand = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int q = tile_quasar(t);
      const QuasarMeta meta = args.meta[q];
      if (!tile_live(q, meta)) continue;
      const int64_t s0 = (int64_t)(t % tiles_per_quasar) * TM + (int64_t)rank * TS;
      const int nchunks = meta.nchunks;
      const double2* prec = reinterpret_cast<const double2*>(xa.pix8 + (int64_t)q * args.NPIX * 8) + lane;
      const int64_t PL = args.NPIX;   // plane stride
      double qacc[SPB], ldm[SPB];
      int lde[SPB];
#pragma unroll
      for (int ss = 0; ss < SPB; ++ss) { qacc[ss] = 0.0; ldm[ss] = 1.0; lde[ss] = 0; }
      // pixel data of the next chunk is fetched one chunk ahead (global/L2 latency off the critical path)
      double2 p01n = prec[PL], p23n = prec[2 * PL], p45n = prec[3 * PL];
      if (CONV_B) {   // the six pad pixels in front of the window: a hand-over of their own (lanes 0..5)
        const uint32_t slot = gh & 1u;
        mbar_wait_d(&hfull[slot], (gh >> 1) & 1u, xa.status, 14);
        if (lane < 6) {
#pragma unroll
          for (int ss = NCA; ss < SPB; ++ss) myraw[ss * RAWS + lane] = hb[(slot * SPB + ss) * KC];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&hempty[slot]);
        ++gh;
      }
      for (int c = 0; c < nchunks; ++c) {
        const double y = p01n.x, v = p01n.y, mu = p23n.x, om2 = p23n.y, cw = p45n.x, cu = p45n.y;
        if (c + 1 < nchunks) {
          prec += KC;
          p01n = prec[PL]; p23n = prec[2 * PL]; p45n = prec[3 * PL];
        }
        // absorption of this chunk from the pair's stage-A warp
        const uint32_t slot = gh & 1u;
        long long tw0 = GPDLA_I8P_PHASES ? clock64() : 0;
        mbar_wait_d(&hfull[slot], (gh >> 1) & 1u, xa.status, 14);
        if (GPDLA_I8P_PHASES) w_hand += clock64() - tw0;
        double a[SPB];
#pragma unroll
        for (int ss = 0; ss < SPB; ++ss) a[ss] = hb[(slot * SPB + ss) * KC];
        __syncwarp();
        if (lane == 0) mbar_arrive(&hempty[slot]);
        ++gh;
        if (CONV_B) {
          // what arrived is the raw profile exp(-N tau) (-1 marks the rows of the null model): instrument convolution
          // over the warp's private rows (6 pixels carried over from the previous chunk)   voigt.c:297-299
#pragma unroll
          for (int ss = NCA; ss < SPB; ++ss) myraw[ss * RAWS + 6 + lane] = a[ss];
          __syncwarp();
          double carry[SPB];
#pragma unroll
          for (int ss = NCA; ss < SPB; ++ss) {
            const double* rb = myraw + ss * RAWS;
            double acc_a = 0.0;
#pragma unroll
            for (int tt = 0; tt < 6; ++tt) acc_a = fma(rb[lane + tt], c_lines.ip[tt], acc_a);
            acc_a = fma(a[ss], c_lines.ip[6], acc_a);   // the lane's own pixel is still in its register
            carry[ss] = rb[KC + (lane < 6 ? lane : 0)];
            a[ss] = (__double2hiint(a[ss]) < 0) ? 1.0 : acc_a;
          }
          __syncwarp();
          if (lane < 6) {
#pragma unroll
            for (int ss = NCA; ss < SPB; ++ss) myraw[ss * RAWS + lane] = carry[ss];
          }
        }
        uint64_t xw[SPB], xu[SPB];
#pragma unroll
        for (int ss = 0; ss < SPB; ++ss) {
          const double a2 = a[ss] * a[ss];
          const double d = fma(a2, om2, v);                // process_qsos.m:194,198
          const double rd = fast_rcp(d);
          const double r = fma(-a[ss], mu, y);
          const double t1 = r * rd;
          // W'' = w cw in [0, CAP], U'' = u cu in [-CAP, CAP], rounded once to F fractional bits by the fused add of
          // the magic constant; + KADD, ^ BIAS turn the two's-complement fixed-point value into L signed digits
          xw[ss] = ((uint64_t)__double_as_longlong(fma(a2 * rd, cw, MAGIC)) + KADD) ^ BIAS;
          xu[ss] = ((uint64_t)__double_as_longlong(fma(a[ss] * t1, cu, MAGIC)) + KADD) ^ BIAS;
          qacc[ss] = fma(r, t1, qacc[ss]);
          ldm[ss] *= d;
        }
        tw0 = GPDLA_I8P_PHASES ? clock64() : 0;
        mbar_wait_d(&bar_empty[stage], empty_parity, xa.status, 1);
        if (GPDLA_I8P_PHASES) w_empty += clock64() - tw0;
        uint8_t* dW = wdst0 + stage * wstride + rowoff;
        uint8_t* dU = udst0 + stage * ustride + rowoff;
        if (PROBE & 4) {   // keep the digits alive without storing them
          uint64_t sink = 0;
#pragma unroll
          for (int ss = 0; ss < SPB; ++ss) sink ^= xw[ss] ^ (xu[ss] << 1);
          if (sink == 0x123456789abcdefull) dW[0] = 1;
        } else {
          store_digit_planes<L, Sh::PLANE_A>(dW, xw);
          store_digit_planes<L, Sh::PLANE_A>(dU, xu);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_rows[stage]);
        if (++stage == STAGES) { stage = 0; empty_parity ^= 1u; }
        if ((c & 7) == 7) {
#pragma unroll
          for (int ss = 0; ss < SPB; ++ss) {
            const int hi = __double2hiint(ldm[ss]);
            const int e2 = ((hi >> 20) & 0x7ff) - 1023;
            lde[ss] += e2;
            ldm[ss] = __hiloint2double(hi - (e2 << 20), __double2loint(ldm[ss]));
          }
        }
      }
      // per-sample scalars into the buffer of this tile's parity (the epilogue of tile it - 2 has long finished:
      // the MMAs of tile it needed the TMEM drain of tile it - 1, which follows the factorisation of tile it - 2)
#pragma unroll
      for (int ss = 0; ss < SPB; ++ss) {
        const double qs = warp_sum(qacc[ss]);
        const double ld = warp_sum(log(ldm[ss]) + (double)lde[ss] * 0.693147180559945309417);
        if (lane == 0) {
          if (Sh::EXT) {   // straight to cholesky_kernel's scalars
            double* qd = args.qld + ((int64_t)q * args.gram_rows + s0 + row0 + ss) * 2;
            qd[0] = qs; qd[1] = ld;
          } else {
            s_q[(it & 1) * TS + row0 + ss] = qs; s_ld[(it & 1) * TS + row0 + ss] = ld;
          }
        }
      }
      __syncwarp();
      if (!Sh::EXT && lane == 0) mbar_arrive(&bar_sq[it & 1]);
      if (GPDLA_I8P_PHASES && xa.phase && lane == 0) {   // per-warp wait cycles of this tile
        atomicAdd(&xa.phase[8 + 1], (unsigned long long)w_empty); atomicAdd(&xa.phase[8 + 14], (unsigned long long)w_hand);
        w_empty = w_hand = 0;
        if (pr == 0) atomicAdd(&xa.phase[7], 1ull);
      }
      ++it;
    }
  } else if (warp >= NCTRL) {
    // =========================================================================== PRODUCERS, STAGE A: absorption
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_A));
    const int pr = warp - NCTRL;
    const int row0 = pr * SPB;
    double* myraw = rawbuf + row0 * RAWS;
    double* hb = hbuf + pr * (2 * SPB * KC) + lane;
    uint64_t* const hfull = bar_hfull + 2 * pr;
    uint64_t* const hempty = bar_hempty + 2 * pr;
    uint32_t gh = 0;                                  // hand-over counter across tiles
    long long w_hand = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int q = tile_quasar(t);
      const QuasarMeta meta = args.meta[q];
      if (!tile_live(q, meta)) continue;
      const int64_t s0 = (int64_t)(t % tiles_per_quasar) * TM + (int64_t)rank * TS;
      const int nchunks = meta.nchunks;
      // per-sample parameters of this warp's four samples (rows private to the warp)
      __syncwarp();
      if (lane < SPB) {
        const int i = row0 + lane;
        const int64_t s = s0 + i;
        const bool is_null = s >= S;
        const int64_t so = sample_at(args, is_null ? S - 1 : s);   // the null-model slot borrows a redshift (a == 1 anyway)
        const double z = __dadd_rn(meta.min_z_dla, __dmul_rn(meta.max_z_dla - meta.min_z_dla, args.offset_samples[so]));
        s_nhi[i] = is_null ? -1.0 : args.nhi_samples[so];
        s_K[i] = rest_table_offset(args.rt, z, meta.lam_ref);
        s_so[i] = (int)so;
        for (int j = 0; j < num_lines; ++j) s_mult[j * TS + i] = line_multiplier(j, z);
        if (MODE == 2) {
          for (int j = 0; j < args.num_partners; ++j)
            s_part[j * TS + i] = is_null ? 0 : args.partners[((int64_t)q * 3 + j) * S + so];
        }
      }
      __syncwarp();
      const double* lam = args.lam_pad + (int64_t)q * (args.NPIX + 8);
      const double2* pix8 = reinterpret_cast<const double2*>(xa.pix8 + (int64_t)q * args.NPIX * 8);
      double* const cache_q = (MODE != 0) ? args.acache + (int64_t)q * S * args.NPIX : nullptr;

      // tau / N: from the rest-frame table (one cell per lane serves the warp's four samples), directly where the
      // cell is near a line centre or the four samples are too far apart in redshift
      const double* lamh = args.lamh + (int64_t)q * (args.NPIX + 8);
      double K_mid;
      const int tab_mode = (MODE != 2) ? group_cell<SPB>(args.rt, s_K + row0, K_mid) : 0;
      auto eval_raw = [&](double lambda, double lh, double (&e)[SPB], const RestCell* cell = nullptr) {   // voigt.c:282-292, 4 samples at one wavelength
        double tau[SPB];
        tau_samples<NL, SPB>(args.rt, tab_mode, K_mid, lambda, lh, s_mult + row0, TS, s_K + row0, num_lines, tau, cell);
        raw_from_tau<SPB, false>(tau, s_nhi + row0, e);   // polynomial exponential: see exp_nonpos
      };
      double eprev[SPB];       // SHFL_CONV: raw profile of the previous chunk (this lane's pixel)
      if (MODE != 2) {   // leading pad pixels p = 0..5
        double e[SPB];
        eval_raw(lam[lane < 6 ? lane : 5], lamh[lane < 6 ? lane : 5], e);
        if (CONV_B) {
          const uint32_t slot = gh & 1u;
          mbar_wait_d(&hempty[slot], ((gh >> 1) & 1u) ^ 1u, xa.status, 15);
#pragma unroll
          for (int ss = NCA; ss < SPB; ++ss) hb[(slot * SPB + ss) * KC] = e[ss];
          __syncwarp();
          if (lane == 0) mbar_arrive(&hfull[slot]);
          ++gh;
        }
        if (SHFL_CONV) {   // the six pad pixels sit in front of pixel 0: lanes 26..31 of the "previous chunk"
#pragma unroll
          for (int ss = 0; ss < SPB; ++ss) eprev[ss] = __shfl_sync(0xffffffffu, e[ss], (lane + 6) & 31);
        } else if (lane < 6) {
#pragma unroll
          for (int ss = 0; ss < NCA; ++ss) myraw[ss * RAWS + lane] = e[ss];
        }
      }
      double rows_n[4][SPB];   // MODE 2: cached absorption rows (sample, partners) of the next chunk
      auto load_rows = [&](int ipix) {
#pragma unroll
        for (int ss = 0; ss < SPB; ++ss) rows_n[0][ss] = cache_q[(int64_t)s_so[row0 + ss] * args.NPIX + ipix];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (j < args.num_partners) {
#pragma unroll
            for (int ss = 0; ss < SPB; ++ss) rows_n[1 + j][ss] = cache_q[(int64_t)s_part[j * TS + row0 + ss] * args.NPIX + ipix];
          }
      };
      if (MODE == 2) load_rows(lane);
      // wavelength and grid position are fetched two chunks ahead, the table cell of the next chunk one chunk ahead: the
      // table (230 KB) lives in L2 -- shared memory leaves the L1 a few KB -- and stage A is the longer stage
      const double2* prec = pix8 + lane;
      double2 plln = prec[0];
      double2 plln2 = prec[nchunks > 1 ? KC : 0];
      RestCell cell_n;
      const bool use_cell = GPDLA_I8P_PF && MODE != 2 && tab_mode == 1;
      if (use_cell) rest_table_fetch(args.rt, plln.y, K_mid, cell_n);
      for (int c = 0; c < nchunks; ++c) {
        const int i = c * KC + lane;
        const double lambda = plln.x, lh = plln.y;
        const RestCell cell = cell_n;
        plln = plln2;
        if (c + 2 < nchunks) {
          prec += KC;
          plln2 = prec[KC];
        }
        if (use_cell && c + 1 < nchunks) rest_table_fetch(args.rt, plln.y, K_mid, cell_n);
        double a[SPB];
        if (MODE != 2) {
          double e[SPB];
          eval_raw(lambda, lh, e, use_cell ? &cell : nullptr);
          if (PROBE & 64) {   // no instrument convolution: no raw-row traffic through shared memory
#pragma unroll
            for (int ss = 0; ss < SPB; ++ss) a[ss] = (__double2hiint(s_nhi[row0 + ss]) < 0) ? 1.0 : e[ss];
          } else {
            // stage B convolves the samples ss >= NCA: their raw profile goes over, the rows of the null model (N marked
            // negative) as -1
#pragma unroll
            for (int ss = NCA; ss < SPB; ++ss) a[ss] = (__double2hiint(s_nhi[row0 + ss]) < 0) ? -1.0 : e[ss];
            if (SHFL_CONV) {
              // instrument convolution through warp shuffles (voigt.c:297-299, same summation order): pixel i needs the
              // raw profile of pixels i - 6 .. i; what lies before lane 0 is the previous chunk's lanes 26..31.  No
              // shared-memory round trip: the convolution's loads were 40 % of a chunk's shared-memory wavefronts
#pragma unroll
              for (int ss = 0; ss < SPB; ++ss) {
                double acc_a = 0.0;
#pragma unroll
                for (int tt = 0; tt < 6; ++tt) {
                  const int d = 6 - tt;
                  const double x = (lane >= 32 - d) ? eprev[ss] : e[ss];
                  acc_a = fma(__shfl_sync(0xffffffffu, x, (lane - d) & 31), c_lines.ip[tt], acc_a);
                }
                acc_a = fma(e[ss], c_lines.ip[6], acc_a);
                eprev[ss] = e[ss];
                a[ss] = (__double2hiint(s_nhi[row0 + ss]) < 0) ? 1.0 : acc_a;   // null model (N marked negative)
              }
            } else if (NCA > 0) {
#pragma unroll
              for (int ss = 0; ss < NCA; ++ss) myraw[ss * RAWS + 6 + lane] = e[ss];
              __syncwarp();
              double carry[SPB];
#pragma unroll
              for (int ss = 0; ss < NCA; ++ss) {
                const double* rb = myraw + ss * RAWS;
                double acc_a = 0.0;
#pragma unroll
                for (int tt = 0; tt < 6; ++tt) acc_a = fma(rb[lane + tt], c_lines.ip[tt], acc_a);   // voigt.c:297-299
                acc_a = fma(e[ss], c_lines.ip[6], acc_a);   // the lane's own pixel is still in its register
                carry[ss] = rb[KC + (lane < 6 ? lane : 0)];
                a[ss] = (__double2hiint(s_nhi[row0 + ss]) < 0) ? 1.0 : acc_a;   // null model (N marked negative)
              }
              __syncwarp();
              if (lane < 6) {
#pragma unroll
                for (int ss = 0; ss < NCA; ++ss) myraw[ss * RAWS + lane] = carry[ss];
              }
            }
          }
          if (MODE == 1) {
#pragma unroll
            for (int ss = 0; ss < SPB; ++ss)
              if (s0 + row0 + ss < S) cache_q[(int64_t)s_so[row0 + ss] * args.NPIX + i] = a[ss];
          }
        } else {
          // absorption = voigt(sample) .* voigt(partner 1) .* ...   (...meanflux.m:342-351): cached rows, fetched one
          // chunk ahead (HBM latency: the cache is 100 MB per quasar, and a producer has nothing else to overlap it with)
#pragma unroll
          for (int ss = 0; ss < SPB; ++ss) a[ss] = rows_n[0][ss];
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (j < args.num_partners) {
#pragma unroll
              for (int ss = 0; ss < SPB; ++ss) a[ss] = a[ss] * rows_n[1 + j][ss];
            }
          if (c + 1 < nchunks) load_rows(i + KC);
        }
        // hand the chunk's absorption to the pair's stage-B warp
        const uint32_t slot = gh & 1u;
        const long long tw0 = GPDLA_I8P_PHASES ? clock64() : 0;
        mbar_wait_d(&hempty[slot], ((gh >> 1) & 1u) ^ 1u, xa.status, 15);
        if (GPDLA_I8P_PHASES) w_hand += clock64() - tw0;
#pragma unroll
        for (int ss = 0; ss < SPB; ++ss) hb[(slot * SPB + ss) * KC] = a[ss];
        __syncwarp();
        if (lane == 0) mbar_arrive(&hfull[slot]);
        ++gh;
      }
      if (GPDLA_I8P_PHASES && xa.phase && lane == 0) { atomicAdd(&xa.phase[8 + 15], (unsigned long long)w_hand); w_hand = 0; }
    }
  } else {
    // =========================================================================== CONTROL WARPGROUP
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_CTRL));
    if (warp == 0 && lane == 0) {
      // ---- MMA issuer
      static_assert(Sh::A_TILE % 16 == 0 && Sh::B_BUF % 16 == 0 && Sh::PLANE_A % 16 == 0, "descriptor offsets are in 16-byte units");
      const uint64_t da_stage0 = make_desc(smem_u32(At), Sh::LBO_A, Sh::SBO_A);
      const uint64_t db_buf0 = make_desc(smem_u32(Bt), 128, 256);
      int gc = 0, it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int q = tile_quasar(t);
        const QuasarMeta meta = args.meta[q];
        if (!tile_live(q, meta)) continue;
        const long long tile_t0 = GPDLA_I8P_PHASES ? clock64() : 0;
        if (it > 0) mbar_wait_d(bar_tfree, (it - 1) & 1, xa.status, 13, phase_ptr(xa), 100);   // accumulators of the previous tile read out
        for (int c = 0; c < meta.nchunks; ++c, ++gc) {
          const int stage = gc % STAGES, buf = gc & 1;
          mbar_wait_d(&bar_full[stage], (gc / STAGES) & 1, xa.status, 2, phase_ptr(xa), 100);
          mbar_wait_d(&bar_pfull[buf], (gc >> 1) & 1, xa.status, 3, phase_ptr(xa));
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint64_t da0 = da_stage0 + (uint64_t)(uint32_t)(stage * (Sh::A_TILE >> 4));
          const uint64_t db0 = db_buf0 + (uint64_t)(uint32_t)(buf * (Sh::B_BUF >> 4));
          if (PROBE & 1) { (void)da0; (void)db0; }
          else if (rank < WCTAS) issue_chunk_mmas_fixed<Sh, L, Sh::NW>(tmem_base, da0, db0, c > 0);
          else issue_chunk_mmas_fixed<Sh, L, Sh::NU>(tmem_base, da0, db0, c > 0);
          mma_commit_multicast(&bar_empty[stage], (uint16_t)((1u << CLUSTER) - 1));
          mma_commit(&bar_pempty[buf]);
        }
        mma_commit(bar_acc);
        if (GPDLA_I8P_PHASES && xa.phase) atomicAdd(&xa.phase[6], (unsigned long long)(clock64() - tile_t0));
        ++it;
      }
    } else if (warp == 1 && lane == 0) {
      // ---- B-operand loader
      int gc = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int q = tile_quasar(t);
        const QuasarMeta meta = args.meta[q];
        if (!tile_live(q, meta)) continue;
        const uint8_t* bsrc = xa.bop + (int64_t)q * (args.NPIX / KC) * Sh::CHUNK_BYTES + Sh::b_offset(slot);
        for (int c = 0; c < meta.nchunks; ++c, ++gc) {
          const int buf = gc & 1;
          if (gc >= 2) mbar_wait_d(&bar_pempty[buf], ((gc >> 1) - 1) & 1, xa.status, 4, phase_ptr(xa), 100);
          mbar_expect_tx(&bar_pfull[buf], b_bytes);
          tma_load_1d(Bt + buf * Sh::B_BUF, bsrc + (int64_t)c * Sh::CHUNK_BYTES, b_bytes, &bar_pfull[buf]);
        }
      }
    } else if (warp == 2 && lane == 0) {
      // ---- row-block sender
      int gc = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int q = tile_quasar(t);
        const QuasarMeta meta = args.meta[q];
        if (!tile_live(q, meta)) continue;
        for (int c = 0; c < meta.nchunks; ++c, ++gc) {
          const int stage = gc % STAGES;
          mbar_wait_d(&bar_rows[stage], (gc / STAGES) & 1, xa.status, 5, phase_ptr(xa), 100);
          const uint32_t dst_off = smem_u32(At + stage * Sh::A_TILE) + rank * Sh::ROWBLOCK;
          const uint32_t other_rows = smem_u32(Sx + stage * Sh::ROWBLOCK);
          const uint32_t fullbar = smem_u32(&bar_full[stage]);
          if (PROBE & 2) { mbar_arrive(&bar_full[stage]); continue; }
          for (uint32_t peer = 0; peer < (uint32_t)CLUSTER; ++peer) {
            if (peer == rank) continue;
            const uint32_t src = (rank < WCTAS && peer < WCTAS) ? dst_off : other_rows;
            dsmem_bulk_copy(mapa(dst_off, peer), src, Sh::ROWBLOCK, mapa(fullbar, peer));
          }
          mbar_arrive_expect_tx(&bar_full[stage], (CLUSTER - 1) * Sh::ROWBLOCK);
          if constexpr (Sh::EXT) {
            // this CTA's W'' block of the chunk -> its place in the stored digit tile (the contract-only passes read whole
            // tiles); the stage is released only after the copy engine has read the block
            uint8_t* gdst = xa.adig + (((int64_t)q * tiles_per_quasar + (t % tiles_per_quasar)) * (args.NPIX / KC) + c) * Sh::A_TILE +
                            rank * Sh::ROWBLOCK;
            bulk_store_global(gdst, (rank < WCTAS) ? dst_off : other_rows, Sh::ROWBLOCK);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(&bar_empty[stage]);
          }
        }
      }
    }
    __syncwarp();
  }
  // nobody leaves while a peer may still address this CTA's shared memory or barriers
  asm volatile("tcgen05.fence::before_thread_sync;");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

}  // namespace i8
}  // namespace gpdla
