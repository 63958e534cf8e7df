// FP64 tensor-core GEMM with a pluggable fused epilogue, sm_100a.
//
//   C[M x N] = A[M x K] * Bt[N x K]^T            (both operands K-contiguous, "TN")
//
// This is the one dense contraction every hot-path operator of the linear-MPC path reduces to:
// the batched operator apply of the regulator QP iteration (A = per-sample vectors, Bt = the
// shared condensed operator), the q-build, the plant step and the structured-MLP layers.  The
// reference does these one sample at a time with NumPy matvecs and cvxopt
// (/root/reference/lib/linearMPC.py:503-504, :860).
//
// Design (B200): FP64 has no tcgen05 kind, so the tensor path is the legacy-encoded
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), measured 37.1 TFLOP/s raw on B200 vs 35.5 cuBLAS
// DGEMM.  CTA tile BM x BN x 16, operands staged global->shared with 16-byte cp.async in a
// STAGES-deep ring (rows padded to 20 doubles = conflict-free 64-bit fragment loads), accumulators
// in registers, epilogue functor applied on the accumulator fragments (no C round trip).
// A rows can be gathered through an index list whose length lives in device memory, so the
// solver can drop converged samples without a host round trip or a physical compaction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nnmpc {

struct GemmOperands {
  const double* A;      // M x K, row stride lda (even, 16-byte aligned rows)
  long long lda;
  const double* Bt;     // N x K, row stride ldb
  long long ldb;
  int M, N, K;
  const int* rows;      // optional: logical row i of A/C -> physical row rows[i]
  const int* m_count;   // optional: device-side logical row count (<= M)
  // device-side tile-shape selection: the launch does nothing unless m_lo < row count <= m_hi, so
  // several tile shapes can be enqueued for a row list whose length only the device knows
  int m_lo = 0, m_hi = 0x7fffffff;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int BK_ = 16>
struct GemmTile {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_;
  static constexpr int BK = BK_, LDS = BK + 4;
  static constexpr int KC = BK / 2;                          // 16-byte chunks per tile row
  static constexpr int THREADS = 32 * WM * WN;
  static constexpr int MT = BM / WM / 8, NT = BN / WN / 8;
  static constexpr int A_ELEMS = BM * LDS, B_ELEMS = BN * LDS;
  static constexpr int SMEM_BYTES = STAGES * (A_ELEMS + B_ELEMS) * 8;
  static constexpr int A_CHUNKS = BM * KC / THREADS;         // 16-byte chunks per thread
  static constexpr int B_CHUNKS = BN * KC / THREADS;
  static_assert(BM % (WM * 8) == 0 && BN % (WN * 8) == 0, "warp tiling");
  static_assert((BM * KC) % THREADS == 0 && (BN * KC) % THREADS == 0, "loader tiling");
  static_assert(THREADS % KC == 0, "every chunk of a thread must share its k offset");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

// Epilogue concept:
//   struct Epi { struct Params {...};
//     __device__ Epi(const Params&, int tile_m0, int tile_n0);
//     __device__ void apply(int prow /*physical row*/, int lrow /*logical row*/, int col, double v0, double v1,
//                           bool c0_ok, bool c1_ok);      // col, col+1; called by ALL lanes (ok flags predicate)
//     __device__ void begin_row();
//     __device__ void finish_row(int prow, int lrow, int slot, bool row_ok);  // once per (thread,row); all lanes call it
//   };
// `slot` identifies the (column tile, warp_n) pair for deterministic per-row partial reductions.
template <class T, class Epi>
__global__ void __launch_bounds__(T::THREADS, 1)
gemm_f64_kernel(GemmOperands g, typename Epi::Params ep) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / T::WN, wn = warp % T::WN;
  const int gq = lane >> 2, tq = lane & 3;     // fragment coordinates
  const int Mraw = g.m_count ? *g.m_count : g.M;   // the tile-shape window tests the true row count
  const int Mact = min(Mraw, g.M);
  // linearised grid, column tiles fastest: CTAs resident together share A row panels and sweep Bt
  const int ntn = (g.N + T::BN - 1) / T::BN;
  const int bn = blockIdx.x % ntn;
  const int n0 = bn * T::BN;
  if (Mraw <= g.m_lo || Mraw > g.m_hi) return;
  // Row tiles: one per CTA when the row count is known at launch; for a device-side count the grid is capped
  // (see launch_gemm) and each CTA strides over the row tiles, so a short list costs few empty CTAs.
  const int bm_stride = gridDim.x / ntn;
  for (int bm = blockIdx.x / ntn; bm * T::BM < Mact; bm += bm_stride) {
  const int m0 = bm * T::BM;
  __syncthreads();   // the previous row tile's last fragments have been read before the ring is refilled

  double* As = smem;
  double* Bs = smem + T::STAGES * T::A_ELEMS;

  // per-thread loader coordinates
  const double* a_src[T::A_CHUNKS];
  int a_dst[T::A_CHUNKS];
  bool a_ok[T::A_CHUNKS];
#pragma unroll
  for (int i = 0; i < T::A_CHUNKS; ++i) {
    int c = tid + i * T::THREADS;
    int r = c / T::KC, kc = c % T::KC;
    int lr = m0 + r;
    a_ok[i] = lr < Mact;
    long long pr = a_ok[i] ? (g.rows ? (long long)g.rows[lr] : (long long)lr) : 0;
    a_src[i] = g.A + pr * g.lda + kc * 2;
    a_dst[i] = r * T::LDS + kc * 2;
  }
  const double* b_src[T::B_CHUNKS];
  int b_dst[T::B_CHUNKS];
  bool b_ok[T::B_CHUNKS];
#pragma unroll
  for (int i = 0; i < T::B_CHUNKS; ++i) {
    int c = tid + i * T::THREADS;
    int r = c / T::KC, kc = c % T::KC;
    int col = n0 + r;
    b_ok[i] = col < g.N;
    b_src[i] = g.Bt + (long long)(b_ok[i] ? col : 0) * g.ldb + kc * 2;
    b_dst[i] = r * T::LDS + kc * 2;
  }
  const int kcoff = (tid % T::KC) * 2;   // every chunk of this thread has the same kc (THREADS % KC == 0)

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * T::BK;
    int rem = (g.K - (k0 + kcoff)) * 8;
    rem = rem < 0 ? 0 : (rem > 16 ? 16 : rem);
    double* as = As + stage * T::A_ELEMS;
    double* bs = Bs + stage * T::B_ELEMS;
#pragma unroll
    for (int i = 0; i < T::A_CHUNKS; ++i) cp_async16(as + a_dst[i], a_src[i] + k0, a_ok[i] ? rem : 0);
#pragma unroll
    for (int i = 0; i < T::B_CHUNKS; ++i) cp_async16(bs + b_dst[i], b_src[i] + k0, b_ok[i] ? rem : 0);
  };

  double acc[T::MT][T::NT][2];
#pragma unroll
  for (int i = 0; i < T::MT; ++i)
#pragma unroll
    for (int j = 0; j < T::NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = (g.K + T::BK - 1) / T::BK;
#pragma unroll
  for (int s = 0; s < T::STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  const int a_frag = (wm * (T::BM / T::WM) + gq) * T::LDS + tq;
  const int b_frag = (wn * (T::BN / T::WN) + gq) * T::LDS + tq;

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<T::STAGES - 2>();
    __syncthreads();
    {  // prefetch tile kt+STAGES-1 into the slot consumed at iteration kt-1
      int nk = kt + T::STAGES - 1;
      if (nk < KT) load_stage(nk % T::STAGES, nk);
      cp_async_commit();
    }
    const double* as = As + (kt % T::STAGES) * T::A_ELEMS + a_frag;
    const double* bs = Bs + (kt % T::STAGES) * T::B_ELEMS + b_frag;
#pragma unroll
    for (int kk = 0; kk < T::BK; kk += 4) {
      double af[T::MT], bf[T::NT];
#pragma unroll
      for (int i = 0; i < T::MT; ++i) af[i] = as[i * 8 * T::LDS + kk];
#pragma unroll
      for (int j = 0; j < T::NT; ++j) bf[j] = bs[j * 8 * T::LDS + kk];
#pragma unroll
      for (int i = 0; i < T::MT; ++i)
#pragma unroll
        for (int j = 0; j < T::NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // fused epilogue on the accumulator fragments
  Epi epi(ep, m0, n0);
  const int slot = bn * T::WN + wn;
#pragma unroll
  for (int i = 0; i < T::MT; ++i) {
    int lr = m0 + wm * (T::BM / T::WM) + i * 8 + gq;
    bool rok = lr < Mact;
    int pr = rok ? (g.rows ? g.rows[lr] : lr) : 0;
    epi.begin_row();
#pragma unroll
    for (int j = 0; j < T::NT; ++j) {
      int col = n0 + wn * (T::BN / T::WN) + j * 8 + tq * 2;
      epi.apply(pr, lr, col, acc[i][j][0], acc[i][j][1], rok && col < g.N, rok && col + 1 < g.N);
    }
    epi.finish_row(pr, lr, slot, rok);   // all 32 lanes participate (warp shuffles inside)
  }
  }  // row tiles of this CTA
}

template <class T, class Epi>
inline cudaError_t launch_gemm(const GemmOperands& g, const typename Epi::Params& ep, cudaStream_t st) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_f64_kernel<T, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         T::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = true;
  }
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  const long long ntn = (g.N + T::BN - 1) / T::BN;
  long long ntm = (g.M + T::BM - 1) / T::BM;
  if (g.m_count) {   // device-side row count: about eight waves of CTAs at most, the CTAs stride over the row tiles
    const long long cap = (8 * 148 + ntn - 1) / ntn;
    if (ntm > cap) ntm = cap;
  }
  gemm_f64_kernel<T, Epi><<<(unsigned)(ntn * ntm), T::THREADS, T::SMEM_BYTES, st>>>(g, ep);
  return cudaGetLastError();
}

// number of per-row partial slots an epilogue with row reductions needs for a given N
template <class T>
inline int gemm_row_slots(int N) { return ((N + T::BN - 1) / T::BN) * T::WN; }

}  // namespace nnmpc
