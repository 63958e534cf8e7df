// Internal interface of the batched regulator-QP solver (shared by qp.cu and sim.cu).
#pragma once
#include "nnmpc_common.cuh"
#include "gemm_f64.cuh"

namespace nnmpc {

// CTA tiles of the FP64 tensor-core GEMM: big for full batches, small for few samples.
using TileBig = GemmTile<128, 128, 2, 4, 4>;
using TileSmall = GemmTile<64, 64, 2, 2, 4>;
inline bool use_big_tile(long long M, int N) { return M > 64 && N > 64; }

// plain store epilogue: C = scale*acc (+bias) (ReLU)
struct EpiStore {
  struct Params {
    double* C;
    long long ldc;
    const double* bias;  // nullable, per column
    int relu;
  };
  Params p;
  __device__ EpiStore(const Params& p_, int, int) : p(p_) {}
  __device__ void begin_row() {}
  __device__ void apply(int pr, int, int col, double v0, double v1, bool ok0, bool ok1) {
    if (p.bias) {
      if (ok0) v0 += p.bias[col];
      if (ok1) v1 += p.bias[col + 1];
    }
    if (p.relu) {
      v0 = v0 > 0.0 ? v0 : 0.0;
      v1 = v1 > 0.0 ? v1 : 0.0;
    }
    double* c = p.C + (long long)pr * p.ldc + col;
    if (ok1 && ((reinterpret_cast<uintptr_t>(c) & 15) == 0)) {
      *reinterpret_cast<double2*>(c) = make_double2(v0, v1);
    } else {
      if (ok0) c[0] = v0;
      if (ok1) c[1] = v1;
    }
  }
  __device__ void finish_row(int, int, int, bool) {}
};

template <class Epi>
inline int gemm_auto(const GemmOperands& g, const typename Epi::Params& ep, cudaStream_t st) {
  cudaError_t e = use_big_tile(g.M, g.N) ? launch_gemm<TileBig, Epi>(g, ep, st)
                                         : launch_gemm<TileSmall, Epi>(g, ep, st);
  count_launch();
  if (e != cudaSuccess) return set_error(NNMPC_ERR_CUDA, "gemm launch failed: %s", cudaGetErrorString(e));
  return 0;
}
inline int row_slots_auto(long long M, int N) {
  return use_big_tile(M, N) ? gemm_row_slots<TileBig>(N) : gemm_row_slots<TileSmall>(N);
}

struct QpOutputs {
  double* cost;      // nullable
  double* kkt;       // nullable
  int* iters;        // nullable
  long long stride;  // element stride between samples in the three arrays
};

int qp_solve_device(nnmpc_qp* h, int B, const double* x0, const double* lb, const double* ub, double* u,
                    double* v_state, int warm, QpOutputs out, double tol, int max_iter, cudaStream_t st,
                    long long* iter_sum_out);

}  // namespace nnmpc

struct nnmpc_qp {
  int n, nxa, nu, N, device;
  double alpha;
  double *P, *Top, *tq, *Mtq, *Kunc;  // device operators
  // scratch, sized for `cap` samples
  long long cap;
  int nslots_cap;
  nnmpc::DevBuf<double> V, W0, W1, C, Ql, part_max, part_sum;
  nnmpc::DevBuf<int> rows0, rows1;
  int* counts;                    // device: [0]=count A, [1]=count B, [2]=hit-maxiter flag
  unsigned long long* iter_sum;   // device accumulator of per-sample iterations
  int* h_pinned;                  // pinned host mirror (4 ints + 1 u64)
  // host-buffer staging for *_host entry points
  nnmpc::DevBuf<double> hx0, hlb, hub, hu, hcost, hkkt;
  nnmpc::DevBuf<int> hiters;
};
