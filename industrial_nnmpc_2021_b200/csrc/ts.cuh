// Internal interface of the batched target selector (shared by ts.cu and sim.cu).
#pragma once
#include "nnmpc_common.cuh"

struct nnmpc_ts {
  int nx, nu, ny, nd, device;
  int* fail = nullptr;                               // device flag: some solve of the last call missed its optimum
  double *Ht, *Fy, *Fd, *f0, *Gx, *Gd, *ulb, *uub;  // device operators
  // output-constrained targets (nnmpc_ts_set_output_bounds): constraint rows Abar = [C Gx; I] (mb x nu), Hinv = Ht^-1,
  // AH = Abar Hinv, Mbar = Abar Hinv Abar', Ryd = C Gd + Cd (ny x nd), output bounds ylb, yub (ny)
  bool general = false;
  int mb = 0;
  double *Hinv = nullptr, *Abar = nullptr, *AH = nullptr, *Mbar = nullptr, *Ryd = nullptr, *ylb = nullptr, *yub = nullptr;
  nnmpc::DevBuf<double> hysp, hd, hxs, hus;          // staging for the host entry point
  nnmpc::DevBuf<int> hiters;
};

namespace nnmpc {

// Optional fused outputs for the closed-loop generator: when `x` is non-null the kernel also
// forms the regulator inputs x0 = [x-xs; uprev-us], lb = ulb-us, ub = uub-us
// (LinearMPCController.get_control_sequence, linearMPC.py:685-688), the warm-start shift
// dus = us_prev - us, and copies (x, uprev) into the dataset row.
struct TsFused {
  const double* x;        // B x nx   current state
  const double* uprev;    // B x nu
  double* x0;             // B x nxa_ld
  int nxa_ld;
  double* lb;             // B x nu
  double* ub;             // B x nu
  double* us_prev;        // B x nu  (in: previous target, out: this target)
  double* dus;            // B x nu
  double* row_x;          // dataset rows, stride row_stride_x / row_stride_u doubles between samples
  double* row_uprev;
  long long row_stride_x, row_stride_u;
};

// Optional indirection for the closed-loop engine: logical sample b is trajectory slot rows[b]
// (b < *count), at its own time index tcur[slot]; the strided arrays (ysp, d, xs, us, iters and the
// dataset rows of TsFused) are then indexed by chunk[slot]*T + tcur[slot], the per-slot buffers of TsFused
// (x, uprev, x0, lb, ub, us_prev, dus) by slot.
struct TsIndex {
  const int* rows;
  const int* count;
  const int* tcur;
  int T;
  const int* chunk = nullptr;   // optional: slot -> trajectory chunk it currently works on (default: the slot itself)
};

int ts_solve_device(nnmpc_ts* h, int B, const double* ysp, long long ysp_stride, const double* d,
                    long long d_stride, double* xs, long long xs_stride, double* us, long long us_stride,
                    int* iters, long long iters_stride, const TsFused* fused, const TsIndex* index,
                    int* fail_flag, cudaStream_t st);

}  // namespace nnmpc
