// Split-operator tensor-core GEMM on the 5th-generation (tcgen05) tensor cores, sm_100a.
//
//   acc[M x N] (fp32, TMEM) = A[M x K] * (B1 + B2)[N x K]^T          A, B1, B2 fp16, K-contiguous
//
// This is the low-precision half of the regulator-QP iteration ("FP32 with FP64 residual
// refinement"): the shared operator Top = (P + D)^-1 D is held as a two-term fp16 split
// B1 + B2 ~ s_T Top (22 significant bits), A holds the per-sample INCREMENT dw of the
// Douglas-Rachford operand (quantised to fp16 with a per-row power-of-two scale, the quantisation
// residual is fed forward exactly), and the epilogue keeps every piece of solver state in FP64:
//     x += acc / (s_T s_row);  d = x - clip(v);  v += alpha d;  dw+ = (2 clip(v) - v) - w_lp
// The increments shrink with the iteration, so the low-precision error is relative to a vanishing
// quantity; what error accumulates is removed by FP64 "anchor" GEMMs (x = Top w - c re-evaluated
// with the FP64 tensor-core kernel) and every returned point is verified with P in FP64.
// Replaces, like gemm_f64.cuh, the per-sample cvxopt solves of
// /root/reference/lib/linearMPC.py:503-504.
//
// Kernel structure (one CTA per SM, persistent over output tiles, warp specialised):
//   warp 0     TMA producer: cp.async.bulk.tensor 2D tiles (128B swizzle) of A, B1, B2 into a
//              STAGES-deep shared-memory ring, completion on mbarriers
//   warp 1     TMEM allocation + single-thread tcgen05.mma issue (M = 128, N = BN, K = 16 per
//              instruction; the B1 and B2 products accumulate into the same TMEM tile),
//              tcgen05.commit releases ring slots and publishes finished accumulators
//   warps 2-9  epilogue: tcgen05.ld (32 lanes x 16 columns per warp and step; two warps share a
//              lane quarter and split the columns) -> transpose through shared memory -> FP64
//              update on registers -> global; TMEM accumulators are double buffered so the epilogue
//              of tile i overlaps the MMAs of tile i+1.  The epilogue moves 42-46 B of state per
//              element through HBM and is what bounds the one-term pass (lp_iter.cuh); the two-term
//              pass is bound by board power (profiles/r02n_pass_energy_diagnosis.md).
// LpTile<BN, STAGES, MR, TERMS>: TERMS = 2 multiplies (B1 + B2) in every pass, TERMS = 1 only B1 (the
// second operator term is then delivered every m-th pass by a TERMS = 1 launch over the pending sums).
#pragma once
#include <cuda.h>   // CUtensorMap and its enums (types only: the encoder is fetched through cudart)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nnmpc {
namespace lp {

constexpr int BM = 128;          // rows (samples) per tile = TMEM lanes
constexpr int BK = 64;           // fp16 elements per k-block = one 128-byte swizzle span
constexpr int UMMA_K = 16;
#ifndef NNMPC_EPI_WARPS
#define NNMPC_EPI_WARPS 8
#endif
constexpr int EPI_WARPS = NNMPC_EPI_WARPS;   // EPI_WARPS / 4 warps per TMEM lane quarter, each owns CW of every 32 accumulator columns
constexpr int CW = 128 / EPI_WARPS;          // accumulator columns per warp and step (16 or 8)
static_assert(EPI_WARPS == 8 || EPI_WARPS == 16, "epilogue warps");
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int ACC_STAGES = 2;

// Per-warp shared memory of the epilogue: the 32 x 16 accumulator block of one step, transposed through shared
// memory so that the FP64 state is read and written in 64-byte runs per row (4 lanes x 16 B), and the per-row
// scalars of the warp's 32 rows.
struct EpiRowInfo {
  int row;          // sample row (-1: not taking part)
  int pw;           // operand row written for it
  double inv_in, s_out, inv_out;
  float g;          // deferred second term: scale of the pending-sum operand relative to the operand written now
  int boff;         // row * nu: element offset of the row's bounds (row 0 for a row that does not take part)
  long long xoff;   // row * n: element offset of the row's state (row 0 for a row that does not take part)
  long long doff;   // pw * ldd: element offset of the operand row written for it
};
// leading dimension of the staging block: bank-conflict-free both for the writes (one column, 32 consecutive rows)
// and for the epilogue's reads (lane = (row group, column pair), see EpiDelta)
constexpr int STG_LD = CW == 16 ? 34 : 36;
struct EpiWarpSmem {
  float stg[CW * STG_LD];
  EpiRowInfo info[32];
};
constexpr int EPI_SMEM_BYTES = EPI_WARPS * (int)sizeof(EpiWarpSmem);

// L2 eviction-priority policies for TMA loads (createpolicy encodings)
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;

// MR_ = 128-row blocks per CTA tile (1 or 2).  With MR = 2 a CTA owns 256 x BN outputs: one staged operator tile
// (B1, B2) feeds the MMAs of both row blocks, so the operator bytes pulled from L2 per flop halve - the pass is
// bound by L2->SM throughput (ncu: 7.5 kB/clk chip-wide at 48 % tensor-pipe activity with MR = 1).
// TERMS_ = operator terms per pass: 2 = A x (B1 + B2) (both products share the staged A tile and the accumulator);
// 1 = A x B1 only (a stage holds A and B1: deeper ring for the same shared memory) - the form of the regulator pass when
// the second operator term is delivered every m-th pass by its own GEMM (lp_iter.cuh, "deferred second term").
template <int BN_, int STAGES_, int MR_ = 1, int TERMS_ = 2>
struct LpTile {
  static constexpr int BN = BN_, STAGES = STAGES_, MR = MR_, TERMS = TERMS_;
  static constexpr int TILE_M = MR * BM;
  static constexpr int A_BYTES = TILE_M * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + TERMS * B_BYTES;
  static_assert(TERMS == 1 || TERMS == 2, "operator terms");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*1 KB alignment slack*/ + 256 /*barriers*/ + EPI_SMEM_BYTES;
  static constexpr int TMEM_COLS = ACC_STAGES * MR * BN;
  static_assert(MR == 1 || MR == 2, "row blocks per tile");
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns: power of two");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// 2D tile load global -> shared, completion (bytes) on an mbarrier; c0 = innermost (k) coordinate
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// same with an L2 eviction-priority hint (the shared operator is re-read by every row tile: keep it resident
// while the per-sample state streams through)
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1,
                                                 uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 TMEM lanes (this warp's quarter) x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_cw(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x16(taddr, r); }
__device__ __forceinline__ void tmem_ld_cw(uint32_t taddr, uint32_t (&r)[8]) { tmem_ld_32x8(taddr, r); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand tile written by TMA with 128-byte swizzle:
// rows of 128 bytes (64 fp16), 8-row groups 1024 bytes apart (SBO), tile base 1024-byte aligned.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address        bits [ 0,14)
  d |= (uint64_t)1 << 16;                     // leading byte offset  bits [16,30): unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset   bits [32,46)
  d |= (uint64_t)1 << 46;                     // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                     // layout type SWIZZLE_128B
  return d;
}
// Instruction descriptor: kind::f16, A = B = fp16 (0), D = fp32 (1), both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------ the kernel
struct LpShape {
  int M;              // rows of A / of the output (samples); upper bound when m_dev is given
  int N;              // rows of B1/B2 = output columns
  int K;              // contraction length (TMA zero-fills beyond it)
  const int* m_dev;   // optional: the row count lives in device memory (0 = nothing to do)
  int group_cols;     // column tiles per L2 group (see tile_coords); <= 0: all
  // Optional: per 128-row tile, does the tile need the second operator term B2?  (null: every tile does.)  Tiles
  // whose rows are all in the late phase of their QP - increments so small that the 2^-11 relative error of a
  // one-term product is far below the solver tolerance - run A x B1 only: half the MMA work, a third less operand
  // traffic.  tile_stat (nullable): [0] += tiles run with one term, [1] += tiles run with both (accounting).
  const unsigned char* need2;
  unsigned long long* tile_stat;
  // Three-product form for a two-term A operand (structured-network layers, mlp_tc.cuh): A = [A_hi | A_lo] stacked
  // along K.  k-blocks >= b_wrap read the operator again from its first k-block (0: no wrap), and only the first
  // kb2 k-blocks multiply the second operator term (<= 0: all), so that
  //     acc = A_hi B1' + A_hi B2' + A_lo B1'      (the 2^-22 term A_lo B2' is never formed).
  int kb2;
  int b_wrap;
  // Probe only (tools/probes/lp_pass_split.py): 1 = no TMA loads and no MMAs, the accumulators are handed to the
  // epilogue as they are - the time of the epilogue alone.
  int skip_mma;
};

// Tile order of the persistent CTAs.  The operator (2 x n x n fp16, 80 MB at n = 4480) does not fit the part of
// the L2 one die can keep while the per-sample state streams through, so the column tiles are walked in groups
// whose operator slice (group_cols x BN rows of T1 and T2) does: within a group the order is row tile major,
// column tile minor, so the ~148 tiles in flight share the same few MB of operand rows and the group's operator
// slice is read from HBM once per pass instead of once per wave of CTAs.
__device__ __forceinline__ void tile_coords(int t, int ntm, int ntn, int group_cols, int& bm, int& bn) {
  if (group_cols <= 0 || group_cols >= ntn) {
    bm = t / ntn;
    bn = t - bm * ntn;
    return;
  }
  const int per_group = ntm * group_cols;
  const int gi = t / per_group;
  const int rem = t - gi * per_group;
  const int c0 = gi * group_cols;
  const int gc = (ntn - c0 < group_cols) ? ntn - c0 : group_cols;   // only the last group can be narrower
  bm = rem / gc;
  bn = c0 + (rem - bm * gc);
}

// Epilogue concept:
//   struct Epi { struct Params {...};
//     __device__ Epi(const Params&);
//     // warp-collective: the warp owns 32 consecutive output rows (lane l holds the accumulators of row pos0 + l)
//     __device__ Epi(const Params&, EpiWarpSmem* warp_smem, int lane);
//     __device__ void begin_tile(int pos0, int M);
//     __device__ void chunk(int col0, const uint32_t (&acc)[CW] /*fp32 bit patterns*/, int N);   // CW columns
//     __device__ void end_tile();
//   };
// optional epilogue hook: prefetch(col0, N) - the state the epilogue will read for the chunk at col0 (after begin_tile)
template <class E>
__device__ __forceinline__ auto epi_prefetch(E& e, int col0, int N, int) -> decltype(e.prefetch(col0, N), void()) {
  e.prefetch(col0, N);
}
template <class E>
__device__ __forceinline__ void epi_prefetch(E&, int, int, long) {}

// optional epilogue forms: prime(col0, N) after begin_tile (start the loads of the tile's first chunk), and
// chunk(col0, acc, N, next_col0) (next_col0 < 0: last chunk of the row block) for epilogues that pipeline their
// loads across chunks
template <class E>
__device__ __forceinline__ auto epi_prime(E& e, int col0, int N, int) -> decltype(e.prime(col0, N), void()) {
  e.prime(col0, N);
}
template <class E>
__device__ __forceinline__ void epi_prime(E&, int, int, long) {}
template <class E>
__device__ __forceinline__ auto epi_chunk(E& e, int col0, const uint32_t (&acc)[CW], int N, int next_col0, int)
    -> decltype(e.chunk(col0, acc, N, next_col0), void()) {
  e.chunk(col0, acc, N, next_col0);
}
template <class E>
__device__ __forceinline__ void epi_chunk(E& e, int col0, const uint32_t (&acc)[CW], int N, int, long) {
  e.chunk(col0, acc, N);
}

template <class T, class Epi>
__global__ void __launch_bounds__(THREADS, 1)
lp_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmB2, LpShape g, typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned operand ring (128B-swizzle atoms span 8 rows x 128 B)
  // (pointer arithmetic on smem_raw, not an integer round trip: the compiler must keep seeing SHARED addresses, or
  //  every access of the epilogue's staging block becomes a generic LD/ST that may alias its global stores)
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + T::STAGES * T::STAGE_BYTES);
  uint64_t* full = bars;                               // [STAGES]      TMA -> MMA
  uint64_t* empty = bars + T::STAGES;                  // [STAGES]      MMA -> TMA
  uint64_t* acc_full = bars + 2 * T::STAGES;           // [ACC_STAGES]  MMA -> epilogue
  uint64_t* acc_empty = acc_full + ACC_STAGES;         // [ACC_STAGES]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
  EpiWarpSmem* epi_smem = reinterpret_cast<EpiWarpSmem*>(ring + T::STAGES * T::STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = g.m_dev ? min(*g.m_dev, g.M) : g.M;
  const int ntn = (g.N + T::BN - 1) / T::BN;
  const int ntm = (M + T::TILE_M - 1) / T::TILE_M;
  const int tiles = ntn * ntm;
  const int KB = (g.K + BK - 1) / BK;
  // does row tile bm need the second operator term?  (flags are per 128 rows)
  auto tile_two = [&](int bm) -> bool {
    if (T::TERMS == 1) return false;
    if (!g.need2) return true;
    bool two = false;
#pragma unroll
    for (int r = 0; r < T::MR; ++r) two = two || g.need2[bm * T::MR + r] != 0;
    return two;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmB2);
    for (int s = 0; s < T::STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, EPI_WARPS);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, T::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (one elected lane) =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < tiles && !g.skip_mma; t += gridDim.x) {
        int bm, bn;
        tile_coords(t, ntm, ntn, g.group_cols, bm, bn);
        const bool two = tile_two(bm);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(empty + s, ph ^ 1);
          uint8_t* st = ring + s * T::STAGE_BYTES;
          const bool two_kb = two && (g.kb2 <= 0 || kb < g.kb2);
          const int kbb = (g.b_wrap > 0 && kb >= g.b_wrap) ? kb - g.b_wrap : kb;
          mbar_expect_tx(full + s, two_kb ? T::STAGE_BYTES : T::A_BYTES + T::B_BYTES);
          tma_load_2d(st, &tmA, full + s, kb * BK, bm * T::TILE_M);      // the A map's box is TILE_M rows
          tma_load_2d_hint(st + T::A_BYTES, &tmB1, full + s, kbb * BK, bn * T::BN, L2_EVICT_LAST);
          if (T::TERMS == 2 && two_kb)
            tma_load_2d_hint(st + T::A_BYTES + T::B_BYTES, &tmB2, full + s, kbb * BK, bn * T::BN, L2_EVICT_LAST);
          if (++s == T::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, T::BN);
      int s = 0;
      uint32_t ph = 0;
      int i = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
        int bm, bn;
        tile_coords(t, ntm, ntn, g.group_cols, bm, bn);
        const bool two = tile_two(bm);
        if (g.tile_stat) atomicAdd(g.tile_stat + (two ? 1 : 0), (unsigned long long)T::MR);   // in 128 x BN units
        const int as = i & 1;
        const uint32_t aph = (uint32_t)(i >> 1) & 1u;
        mbar_wait(acc_empty + as, aph ^ 1);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(as * T::MR * T::BN);
        for (int kb = 0; kb < (g.skip_mma ? 0 : KB); ++kb) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + s * T::STAGE_BYTES);
          const uint64_t db1 = make_sw128_kmajor_desc(sa + T::A_BYTES);
          const uint64_t db2 = make_sw128_kmajor_desc(sa + T::A_BYTES + T::B_BYTES);
          const bool two_kb = two && (g.kb2 <= 0 || kb < g.kb2);
#pragma unroll
          for (int r = 0; r < T::MR; ++r) {       // row block r: 128 rows of A (16 KB apart in the box), own accumulator
            const uint64_t da = make_sw128_kmajor_desc(sa + r * (BM * BK * 2));
            const uint32_t tr = tacc + (uint32_t)(r * T::BN);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)   // +32 bytes (2 x 16 B units) per K = 16 step inside the swizzle span
              umma_f16(tr, da + 2 * k, db1 + 2 * k, idesc, (kb | k) ? 1u : 0u);
            if (T::TERMS == 2 && two_kb) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) umma_f16(tr, da + 2 * k, db2 + 2 * k, idesc, 1u);
            }
          }
          umma_commit(empty + s);                 // ring slot free once these MMAs have read it
          if (++s == T::STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(acc_full + as);               // accumulator complete
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    Epi epi(ep, epi_smem + (warp - 2), lane);
    int i = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      int bm, bn;
      tile_coords(t, ntm, ntn, g.group_cols, bm, bn);
      const int as = i & 1;
      const uint32_t aph = (uint32_t)(i >> 1) & 1u;
#pragma unroll 1
      for (int r = 0; r < T::MR; ++r) {
        epi.begin_tile(bm * T::TILE_M + r * BM + q * 32, M);
        epi_prefetch(epi, bn * T::BN + hsel * CW, g.N, 0);
        epi_prime(epi, bn * T::BN + hsel * CW, g.N, 0);
        if (r == 0) {
          mbar_wait(acc_full + as, aph);
          tc_fence_after();
        }
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * T::MR + r) * T::BN);
#pragma unroll 1
        for (int cc = 0; cc < T::BN / 32; ++cc) {
          uint32_t acc[CW];
          tmem_ld_cw(tacc + (uint32_t)(cc * 32 + hsel * CW), acc);
          if (cc + 1 < T::BN / 32) epi_prefetch(epi, bn * T::BN + (cc + 1) * 32 + hsel * CW, g.N, 0);
          tmem_ld_wait();
          epi_chunk(epi, bn * T::BN + cc * 32 + hsel * CW, acc, g.N,
                    cc + 1 < T::BN / 32 ? bn * T::BN + (cc + 1) * 32 + hsel * CW : -1, 0);
        }
        epi.end_tile();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + as);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, T::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ CTA-pair kernel
// Same contraction with cta_group::2: a cluster of two CTAs (one TPC) owns a 256 x 256 output tile.  Each CTA
// loads its own 128 rows of A and HALF of the B1/B2 tiles (128 of the 256 operator rows); one thread of the
// leader CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads both CTAs' shared memory and writes each
// CTA's 128 accumulator rows into its own TMEM.  Per output element this halves the operand bytes pulled
// from L2, which is what bounds the single-CTA kernel (6.3 kB/clk chip-wide TMA throughput).
constexpr int BN2 = 256;          // output columns per pair tile (128 operator rows staged per CTA)
constexpr int STAGES2 = EPI_WARPS == 16 ? 3 : 4;
constexpr int STAGE2_BYTES = BM * BK * 2 + 2 * (BN2 / 2) * BK * 2;   // A + B1 half + B2 half = 48 KB
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256 + EPI_SMEM_BYTES;
static_assert(SMEM2_BYTES <= 227 * 1024, "shared memory");
constexpr int TMEM2_COLS = ACC_STAGES * BN2;                          // 512: all of TMEM

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(rank)
      : "memory");
}
// TMA load issued by either CTA of a pair; the bytes are credited to the LEADER CTA's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1,
                                                      uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once the pair MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
lp_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmB2, LpShape g, typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + STAGES2 * STAGE2_BYTES);
  uint64_t* full = bars;                          // [STAGES2]  leader's copy is used: both CTAs' TMA bytes + 2 arrivals
  uint64_t* empty = bars + STAGES2;               // [STAGES2]  per CTA, released by the leader's multicast commit
  uint64_t* acc_full = bars + 2 * STAGES2;        // [ACC_STAGES] per CTA, multicast commit
  uint64_t* acc_empty = acc_full + ACC_STAGES;    // [ACC_STAGES] leader's copy is used: both CTAs' epilogue warps arrive
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
  EpiWarpSmem* epi_smem = reinterpret_cast<EpiWarpSmem*>(ring + STAGES2 * STAGE2_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int M = g.m_dev ? min(*g.m_dev, g.M) : g.M;
  const int ntn = (g.N + BN2 - 1) / BN2;
  const int ntm = (M + 2 * BM - 1) / (2 * BM);
  const int tiles = ntn * ntm;
  const int KB = (g.K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmB2);
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(full + s, 2);
      mbar_init(empty + s, 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, TMEM2_COLS);
  tc_fence_before();
  cluster_sync_all();          // both CTAs' barriers and TMEM allocations exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (one lane in EACH CTA: own A rows, own half of the operator tile) =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = cluster_id; t < tiles; t += nclusters) {
        int bm, bn;
      tile_coords(t, ntm, ntn, g.group_cols, bm, bn);
        const int arow = bm * 2 * BM + (int)rank * BM;
        const int brow = bn * BN2 + (int)rank * (BN2 / 2);
        const bool two = !g.need2 || (g.need2[2 * bm] | g.need2[2 * bm + 1]) != 0;     // a pair tile spans two 128-row tiles
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(empty + s, ph ^ 1);
          uint8_t* st = ring + s * STAGE2_BYTES;
          tma_load_2d_pair(st, &tmA, full + s, kb * BK, arow);
          tma_load_2d_pair_hint(st + BM * BK * 2, &tmB1, full + s, kb * BK, brow, L2_EVICT_LAST);
          if (two) tma_load_2d_pair_hint(st + BM * BK * 2 + (BN2 / 2) * BK * 2, &tmB2, full + s, kb * BK, brow, L2_EVICT_LAST);
          if (leader) mbar_expect_tx(full + s, two ? 2 * STAGE2_BYTES : 2 * (BM * BK * 2 + (BN2 / 2) * BK * 2));
          else mbar_arrive_remote(full + s, 0);
          if (++s == STAGES2) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one lane of the leader CTA =====
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(2 * BM, BN2);
      int s = 0;
      uint32_t ph = 0;
      int i = 0;
      for (int t = cluster_id; t < tiles; t += nclusters, ++i) {
        int bm, bn;
        tile_coords(t, ntm, ntn, g.group_cols, bm, bn);
        const bool two = !g.need2 || (g.need2[2 * bm] | g.need2[2 * bm + 1]) != 0;
        if (g.tile_stat) atomicAdd(g.tile_stat + (two ? 1 : 0), 4ull);      // counted in 128 x 128 tiles
        const int as = i & 1;
        const uint32_t aph = (uint32_t)(i >> 1) & 1u;
        mbar_wait(acc_empty + as, aph ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(as * BN2);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + s * STAGE2_BYTES);
          const uint64_t da = make_sw128_kmajor_desc(sa);
          const uint64_t db1 = make_sw128_kmajor_desc(sa + BM * BK * 2);
          const uint64_t db2 = make_sw128_kmajor_desc(sa + BM * BK * 2 + (BN2 / 2) * BK * 2);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_f16_pair(tacc, da + 2 * k, db1 + 2 * k, idesc, (kb | k) ? 1u : 0u);
          if (two) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma_f16_pair(tacc, da + 2 * k, db2 + 2 * k, idesc, 1u);
          }
          umma_commit_pair(empty + s);
          if (++s == STAGES2) { s = 0; ph ^= 1; }
        }
        umma_commit_pair(acc_full + as);
      }
    }
  } else {
    // ===== epilogue warps of both CTAs: own 128 accumulator rows =====
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    Epi epi(ep, epi_smem + (warp - 2), lane);
    int i = 0;
    for (int t = cluster_id; t < tiles; t += nclusters, ++i) {
      int bm, bn;
      tile_coords(t, ntm, ntn, g.group_cols, bm, bn);
      const int as = i & 1;
      const uint32_t aph = (uint32_t)(i >> 1) & 1u;
      epi.begin_tile(bm * 2 * BM + (int)rank * BM + q * 32, M);
      epi_prime(epi, bn * BN2 + hsel * CW, g.N, 0);
      mbar_wait(acc_full + as, aph);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN2);
#pragma unroll 1
      for (int cc = 0; cc < BN2 / 32; ++cc) {
        uint32_t acc[CW];
        tmem_ld_cw(tacc + (uint32_t)(cc * 32 + hsel * CW), acc);
        tmem_ld_wait();
        epi_chunk(epi, bn * BN2 + cc * 32 + hsel * CW, acc, g.N, cc + 1 < BN2 / 32 ? bn * BN2 + (cc + 1) * 32 + hsel * CW : -1, 0);
      }
      epi.end_tile();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(acc_empty + as);
        else mbar_arrive_remote(acc_empty + as, 0);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();          // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM2_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side
// cuTensorMapEncodeTiled fetched through the runtime (libnnmpc links cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 row-major matrix [rows][cols] with leading dimension ld (elements, multiple of 8) -> tensor map with
// boxes of box_rows x 64 elements, 128-byte swizzle, zero fill outside [rows] x [cols]
inline bool make_tmap_f16(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <class T, class Epi>
inline cudaError_t launch_lp_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB1, const CUtensorMap& tmB2,
                                  const LpShape& g, const typename Epi::Params& ep, int num_sms, cudaStream_t st) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(lp_gemm_kernel<T, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = true;
  }
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return cudaSuccess;
  const long long tiles = (long long)((g.N + T::BN - 1) / T::BN) * ((g.M + T::TILE_M - 1) / T::TILE_M);
  const unsigned grid = (unsigned)(tiles < num_sms ? tiles : num_sms);
  lp_gemm_kernel<T, Epi><<<grid, THREADS, T::SMEM_BYTES, st>>>(tmA, tmB1, tmB2, g, ep);
  return cudaGetLastError();
}

template <class Epi>
inline cudaError_t launch_lp_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB1, const CUtensorMap& tmB2,
                                       const LpShape& g, const typename Epi::Params& ep, int num_sms, cudaStream_t st) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(lp_gemm_pair_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = true;
  }
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return cudaSuccess;
  const long long tiles = (long long)((g.N + BN2 - 1) / BN2) * ((g.M + 2 * BM - 1) / (2 * BM));
  long long clusters = num_sms / 2;
  if (tiles < clusters) clusters = tiles;
  lp_gemm_pair_kernel<Epi><<<(unsigned)(2 * clusters), THREADS, SMEM2_BYTES, st>>>(tmA, tmB1, tmB2, g, ep);
  return cudaGetLastError();
}

}  // namespace lp
}  // namespace nnmpc
