// Closed-loop offline data generation for B trajectories at once.
//
// Replaces simulate_offline (/root/reference/lib/linearMPC.py:827-880) run in one OS process per
// trajectory chunk (:803-825).  Per time step, for all trajectories together:
//   1. k_target_selector (fused): (ysp_t, d_t) -> (xs, us); dataset rows x, uprev, xs, us;
//      regulator inputs x0 = [x-xs; uprev-us], lb = ulb-us, ub = uub-us       (:851-855, :685-688)
//   2. k_warm_shift: previous solver state shifted one stage and re-centred on the new target
//      (the reference cold-starts cvxopt every step; consecutive QPs are near-identical)
//   3. qp_solve_device: batched regulator QP                                   (:853, :495-512)
//   4. k_advance: u = useq[0:nu] + us -> dataset row u; plant input [x | u | d]  (:856, :689)
//   5. plant step x+ = [x|u|d] [A|B|Bd]' through the FP64 tensor-core GEMM     (:860-861)
#include "qp.cuh"
#include "ts.cuh"

struct nnmpc_sim {
  nnmpc_qp* qp;
  nnmpc_ts* ts;
  int nx, nu, nd, ny, device;
  int kin_ld;    // nx+nu+nd rounded up to even
  int nxa_ld;    // = qp->nxa
  double* ABd;   // device nx x kin_ld
  long long cap;
  long long warm_B;   // batch size whose solver state (Va/Vb, us_prev) is valid for `resume`; 0 = none
  double* warm_V;     // which of Va/Vb holds it
  nnmpc::DevBuf<double> x0, lb, ub, us_prev, dus, Va, Vb, U, xin, xcur, upcur;
  // staging for the host entry point
  nnmpc::DevBuf<double> h_sp, h_dist, h_x, h_uprev, h_xs, h_us, h_u, h_kkt, h_xio, h_upio;
  nnmpc::DevBuf<int> h_iters;
};

namespace nnmpc {

// v_new[j] = v_old[j+nu] + dus[j % nu]  (last stage repeated)
__global__ void k_warm_shift(const double* __restrict__ Vo, double* __restrict__ Vn, const double* __restrict__ dus,
                             long long total, int n, int nu) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    long long r = i / n;
    int j = (int)(i - r * n);
    int js = j + nu < n ? j + nu : j;
    Vn[i] = Vo[r * n + js] + dus[r * nu + (j % nu)];
  }
}

// first move + dataset row u + plant-step input [x | u | d | 0-pad]
__global__ void k_advance(const double* __restrict__ U, const double* __restrict__ us, long long us_stride,
                          const double* __restrict__ xcur, const double* __restrict__ d, long long d_stride,
                          double* __restrict__ row_u, long long row_stride_u, double* __restrict__ upcur,
                          double* __restrict__ xin, int B, int n, int nx, int nu, int nd, int kin_ld) {
  long long total = (long long)B * kin_ld;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    long long b = i / kin_ld;
    int c = (int)(i - b * kin_ld);
    double v;
    if (c < nx) {
      v = xcur[b * nx + c];
    } else if (c < nx + nu) {
      int k = c - nx;
      v = U[b * n + k] + us[b * us_stride + k];
      row_u[b * row_stride_u + k] = v;
      upcur[b * nu + k] = v;
    } else if (c < nx + nu + nd) {
      v = d[b * d_stride + (c - nx - nu)];
    } else {
      v = 0.0;
    }
    xin[i] = v;
  }
}

static int sim_ensure(nnmpc_sim* h, long long B) {
  if (B <= h->cap) return 0;
  const long long n = h->qp->n;
  NNMPC_TRY(h->x0.ensure(B * h->nxa_ld));
  NNMPC_TRY(h->lb.ensure(B * h->nu));
  NNMPC_TRY(h->ub.ensure(B * h->nu));
  NNMPC_TRY(h->us_prev.ensure(B * h->nu));
  NNMPC_TRY(h->dus.ensure(B * h->nu));
  NNMPC_TRY(h->Va.ensure(B * n));
  NNMPC_TRY(h->Vb.ensure(B * n));
  NNMPC_TRY(h->U.ensure(B * n));
  NNMPC_TRY(h->xin.ensure(B * h->kin_ld));
  NNMPC_TRY(h->xcur.ensure(B * h->nx));
  NNMPC_TRY(h->upcur.ensure(B * h->nu));
  h->cap = B;
  h->warm_B = 0;
  return 0;
}

static int sim_run_device(nnmpc_sim* h, int B, int T, double* x_io, double* uprev_io, const double* sp,
                          const double* dist, double* ox, double* ouprev, double* oxs, double* ous, double* ou,
                          int* oiters, double* okkt, double tol, int max_iter, int resume, cudaStream_t st) {
  if (B <= 0 || T <= 0) return 0;
  NNMPC_TRY(sim_ensure(h, B));
  const bool cont = resume && h->warm_B == B;
  h->warm_B = 0;
  const int nx = h->nx, nu = h->nu, nd = h->nd, ny = h->ny, n = h->qp->n;
  const long long sx = (long long)T * nx, su = (long long)T * nu;
  int rc_warn = 0;
  NNMPC_CUDA(cudaMemcpyAsync(h->xcur.p, x_io, (size_t)B * nx * 8, cudaMemcpyDeviceToDevice, st));
  NNMPC_CUDA(cudaMemcpyAsync(h->upcur.p, uprev_io, (size_t)B * nu * 8, cudaMemcpyDeviceToDevice, st));
  if (!cont) NNMPC_CUDA(cudaMemsetAsync(h->us_prev.p, 0, (size_t)B * nu * 8, st));
  double* Vold = cont ? h->warm_V : h->Va.p;
  double* Vnew = Vold == h->Va.p ? h->Vb.p : h->Va.p;
  const int ew_blocks = 148 * 16;
  for (int t = 0; t < T; ++t) {
    TsFused F{};
    F.x = h->xcur.p; F.uprev = h->upcur.p; F.x0 = h->x0.p; F.nxa_ld = h->nxa_ld; F.lb = h->lb.p; F.ub = h->ub.p;
    F.us_prev = h->us_prev.p; F.dus = h->dus.p;
    F.row_x = ox + (long long)t * nx; F.row_uprev = ouprev + (long long)t * nu;
    F.row_stride_x = sx; F.row_stride_u = su;
    NNMPC_TRY(ts_solve_device(h->ts, B, sp + (long long)t * ny, (long long)T * ny, dist + (long long)t * nd,
                              (long long)T * nd, oxs + (long long)t * nx, sx, ous + (long long)t * nu, su, nullptr, 0,
                              &F, st));
    int warm = 0;
    if (t > 0 || cont) {
      k_warm_shift<<<ew_blocks, 256, 0, st>>>(Vold, Vnew, h->dus.p, (long long)B * n, n, nu);
      count_launch();
      warm = 1;
    }
    QpOutputs out{nullptr, okkt ? okkt + t : nullptr, oiters ? oiters + t : nullptr, T};
    int rc = qp_solve_device(h->qp, B, h->x0.p, h->lb.p, h->ub.p, h->U.p, Vnew, warm, out, tol, max_iter, st, nullptr);
    if (rc < 0) return rc;
    rc_warn |= rc;
    k_advance<<<ew_blocks, 256, 0, st>>>(h->U.p, ous + (long long)t * nu, su, h->xcur.p, dist + (long long)t * nd,
                                         (long long)T * nd, ou + (long long)t * nu, su, h->upcur.p, h->xin.p, B, n, nx,
                                         nu, nd, h->kin_ld);
    count_launch();
    GemmOperands g{h->xin.p, h->kin_ld, h->ABd, h->kin_ld, B, nx, h->kin_ld, nullptr, nullptr};
    NNMPC_TRY(gemm_auto<EpiStore>(g, EpiStore::Params{h->xcur.p, nx, nullptr, 0}, st));
    double* tmp = Vold; Vold = Vnew; Vnew = tmp;
  }
  NNMPC_CUDA(cudaMemcpyAsync(x_io, h->xcur.p, (size_t)B * nx * 8, cudaMemcpyDeviceToDevice, st));
  NNMPC_CUDA(cudaMemcpyAsync(uprev_io, h->upcur.p, (size_t)B * nu * 8, cudaMemcpyDeviceToDevice, st));
  NNMPC_CUDA(cudaGetLastError());
  h->warm_B = B;
  h->warm_V = Vold;   // after the final swap: the state the last solve left behind
  return rc_warn;
}

}  // namespace nnmpc

using namespace nnmpc;

extern "C" {

int nnmpc_sim_create(nnmpc_sim_t** out, nnmpc_qp_t* qp, nnmpc_ts_t* ts, int nx, int nu, int nd, int ny,
                     const double* ABd_host, int device) {
  if (!out || !qp || !ts || !ABd_host) return set_error(NNMPC_ERR_BADARG, "nnmpc_sim_create: null argument");
  if (qp->device != device || ts->device != device)
    return set_error(NNMPC_ERR_BADARG, "nnmpc_sim_create: handles live on different devices");
  if (qp->nu != nu || ts->nu != nu || ts->nx != nx || ts->nd != nd || ts->ny != ny || qp->nxa < nx + nu)
    return set_error(NNMPC_ERR_BADARG, "nnmpc_sim_create: inconsistent sizes");
  DeviceGuard dg(device);
  nnmpc_sim* h = new (std::nothrow) nnmpc_sim();
  if (!h) return set_error(NNMPC_ERR_NOMEM, "out of host memory");
  h->qp = qp; h->ts = ts; h->nx = nx; h->nu = nu; h->nd = nd; h->ny = ny; h->device = device;
  h->nxa_ld = qp->nxa;
  const int kin = nx + nu + nd;
  h->kin_ld = (kin + 1) & ~1;
  h->cap = 0;
  h->warm_B = 0;
  h->warm_V = nullptr;
  // pad [A|B|Bd] rows to an even leading dimension for the 16-byte operand loader
  double* tmp = new (std::nothrow) double[(size_t)nx * h->kin_ld];
  if (!tmp) return set_error(NNMPC_ERR_NOMEM, "out of host memory");
  for (int r = 0; r < nx; ++r) {
    for (int c = 0; c < kin; ++c) tmp[(size_t)r * h->kin_ld + c] = ABd_host[(size_t)r * kin + c];
    for (int c = kin; c < h->kin_ld; ++c) tmp[(size_t)r * h->kin_ld + c] = 0.0;
  }
  int rc = upload(&h->ABd, tmp, (size_t)nx * h->kin_ld);
  delete[] tmp;
  if (rc < 0) return rc;
  *out = h;
  return 0;
}

int nnmpc_sim_destroy(nnmpc_sim_t* h) {
  if (!h) return 0;
  DeviceGuard dg(h->device);
  cudaFree(h->ABd);
  h->x0.release(); h->lb.release(); h->ub.release(); h->us_prev.release(); h->dus.release(); h->Va.release();
  h->Vb.release(); h->U.release(); h->xin.release(); h->xcur.release(); h->upcur.release();
  h->h_sp.release(); h->h_dist.release(); h->h_x.release(); h->h_uprev.release(); h->h_xs.release();
  h->h_us.release(); h->h_u.release(); h->h_kkt.release(); h->h_xio.release(); h->h_upio.release();
  h->h_iters.release();
  delete h;
  return 0;
}

int nnmpc_sim_run(nnmpc_sim_t* h, int B, int T, double* x_io, double* uprev_io, const double* setpoints,
                  const double* disturbances, double* x, double* uprev, double* xs, double* us, double* u, int* iters,
                  double* kkt, double tol, int max_iter, int resume, void* stream) {
  if ((B == 0 || T == 0) && h) return 0;
  if (!h || !x_io || !uprev_io || !setpoints || !disturbances || !x || !uprev || !xs || !us || !u)
    return set_error(NNMPC_ERR_BADARG, "nnmpc_sim_run: null argument");
  if (B < 0 || T < 0) return set_error(NNMPC_ERR_BADARG, "nnmpc_sim_run: negative size");
  DeviceGuard dg(h->device);
  return sim_run_device(h, B, T, x_io, uprev_io, setpoints, disturbances, x, uprev, xs, us, u, iters, kkt, tol,
                        max_iter, resume, (cudaStream_t)stream);
}

int nnmpc_sim_run_host(nnmpc_sim_t* h, int B, int T, double* x_io, double* uprev_io, const double* setpoints,
                       const double* disturbances, double* x, double* uprev, double* xs, double* us, double* u,
                       int* iters, double* kkt, double tol, int max_iter, int resume) {
  if ((B == 0 || T == 0) && h) return 0;
  if (!h || !x_io || !uprev_io || !setpoints || !disturbances || !x || !uprev || !xs || !us || !u)
    return set_error(NNMPC_ERR_BADARG, "nnmpc_sim_run_host: null argument");
  if (B <= 0 || T <= 0) return (B == 0 || T == 0) ? 0 : set_error(NNMPC_ERR_BADARG, "nnmpc_sim_run_host: negative size");
  DeviceGuard dg(h->device);
  const size_t bt = (size_t)B * T;
  const int nx = h->nx, nu = h->nu, nd = h->nd, ny = h->ny;
  NNMPC_TRY(h->h_sp.ensure(bt * ny));
  NNMPC_TRY(h->h_dist.ensure(bt * (nd > 0 ? nd : 1)));
  NNMPC_TRY(h->h_x.ensure(bt * nx));
  NNMPC_TRY(h->h_xs.ensure(bt * nx));
  NNMPC_TRY(h->h_uprev.ensure(bt * nu));
  NNMPC_TRY(h->h_us.ensure(bt * nu));
  NNMPC_TRY(h->h_u.ensure(bt * nu));
  NNMPC_TRY(h->h_kkt.ensure(bt));
  NNMPC_TRY(h->h_iters.ensure(bt));
  NNMPC_TRY(h->h_xio.ensure((size_t)B * nx));
  NNMPC_TRY(h->h_upio.ensure((size_t)B * nu));
  cudaStream_t st = 0;
  NNMPC_CUDA(cudaMemcpyAsync(h->h_sp.p, setpoints, bt * ny * 8, cudaMemcpyHostToDevice, st));
  if (nd > 0) NNMPC_CUDA(cudaMemcpyAsync(h->h_dist.p, disturbances, bt * nd * 8, cudaMemcpyHostToDevice, st));
  NNMPC_CUDA(cudaMemcpyAsync(h->h_xio.p, x_io, (size_t)B * nx * 8, cudaMemcpyHostToDevice, st));
  NNMPC_CUDA(cudaMemcpyAsync(h->h_upio.p, uprev_io, (size_t)B * nu * 8, cudaMemcpyHostToDevice, st));
  int rc = sim_run_device(h, B, T, h->h_xio.p, h->h_upio.p, h->h_sp.p, h->h_dist.p, h->h_x.p, h->h_uprev.p, h->h_xs.p,
                          h->h_us.p, h->h_u.p, h->h_iters.p, h->h_kkt.p, tol, max_iter, resume, st);
  if (rc < 0) return rc;
  NNMPC_CUDA(cudaMemcpyAsync(x, h->h_x.p, bt * nx * 8, cudaMemcpyDeviceToHost, st));
  NNMPC_CUDA(cudaMemcpyAsync(xs, h->h_xs.p, bt * nx * 8, cudaMemcpyDeviceToHost, st));
  NNMPC_CUDA(cudaMemcpyAsync(uprev, h->h_uprev.p, bt * nu * 8, cudaMemcpyDeviceToHost, st));
  NNMPC_CUDA(cudaMemcpyAsync(us, h->h_us.p, bt * nu * 8, cudaMemcpyDeviceToHost, st));
  NNMPC_CUDA(cudaMemcpyAsync(u, h->h_u.p, bt * nu * 8, cudaMemcpyDeviceToHost, st));
  if (iters) NNMPC_CUDA(cudaMemcpyAsync(iters, h->h_iters.p, bt * 4, cudaMemcpyDeviceToHost, st));
  if (kkt) NNMPC_CUDA(cudaMemcpyAsync(kkt, h->h_kkt.p, bt * 8, cudaMemcpyDeviceToHost, st));
  NNMPC_CUDA(cudaMemcpyAsync(x_io, h->h_xio.p, (size_t)B * nx * 8, cudaMemcpyDeviceToHost, st));
  NNMPC_CUDA(cudaMemcpyAsync(uprev_io, h->h_upio.p, (size_t)B * nu * 8, cudaMemcpyDeviceToHost, st));
  NNMPC_CUDA(cudaStreamSynchronize(st));
  return rc;
}

}  // extern "C"
