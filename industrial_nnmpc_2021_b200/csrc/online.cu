// Batched ONLINE closed loop: S independent scenarios advanced in lock step, every step
//     Kalman filter -> target selector -> controller -> average stage cost -> plant step + measurement
// exactly as LinearMPCController.control_law / online_simulation do for one scenario at a time
// (/root/reference/lib/linearMPC.py:646-669, :703-718, filter :133-176, plant :87-131), for the three controller
// kinds of the reference's validation study (lib/controller_evaluation.py:322-523): the linear MPC itself, the
// structured neural network (:841-892) and the saturated LQR (:895-1006).
//
// Everything a step needs is one of the batched kernels of this library:
//   filter       xhat+ = (I - L C)(A xhat + B uprev) + L y  =  Fkf [xhat; uprev; y]        one FP64 GEMM
//   targets      nnmpc_ts kernel, fused with x0 = [xhat - xs; uprev - us], lb/ub = bounds - us (:685-688)
//   controller   batched regulator QP (warm started from the previous step) | structured network | clip(K x0 + us)
//   plant        [x+; y+ - v] = Fpl [x; u; p],  Fpl = [A B Bp; CA CB CBp]                    one FP64 GEMM
// The scenario loop stays on the host (T steps), the scenario dimension is the GEMM M dimension.
#include "qp.cuh"
#include "ts.cuh"

namespace nnmpc {
int mlp_forward_dispatch(nnmpc_mlp* h, long long B, const double* x, const double* uprev, const double* xs,
                         const double* us, const double* xscale, const double* ulb, const double* uub, double* out,
                         cudaStream_t st);
int mlp_dims(const nnmpc_mlp* h, int* nx, int* nu, int* with_uprev);
}

struct nnmpc_online {
  nnmpc_qp* qp;
  nnmpc_ts* ts;
  nnmpc_mlp* mlp;
  int nx, nu, ny, nd, np, device;
  int nxa, nxa_ld, kf_ld, pl_ld;
  double *Fkf, *Fpl, *Kaug, *Qaug, *Raug, *Maug, *ulb, *uub, *xscale;
  nnmpc::DevBuf<double> Zkf, xh, dh, x0, lb, ub, xs, us, U, V, uc, Zpl, xy, usp, dus, avg;
  int* fail;
};

namespace nnmpc {

// Zkf row = [xhat | dhat | uprev | y | 0-pad]
__global__ void k_online_pack_kf(double* __restrict__ Z, int ld, const double* __restrict__ xh, const double* __restrict__ dh,
                                 const double* __restrict__ up, const double* __restrict__ y, long long y_stride, int S,
                                 int nx, int nd, int nu, int ny) {
  const long long total = (long long)S * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / ld;
    const int c = (int)(i - s * ld);
    double v = 0.0;
    if (c < nx) v = xh[s * nx + c];
    else if (c < nx + nd) v = dh[s * nd + (c - nx)];
    else if (c < nx + nd + nu) v = up[s * nu + (c - nx - nd)];
    else if (c < nx + nd + nu + ny) v = y[s * y_stride + (c - nx - nd - nu)];
    Z[i] = v;
  }
}

// one CTA per scenario: control input of this step, running average stage cost (linearMPC.py:691-701), records,
// next uprev and the plant-step operand [x | u | p | 0-pad]
__global__ void __launch_bounds__(128)
k_online_finish(int kind, int S, int T, int t, int n, const double* __restrict__ U, const double* __restrict__ uc,
                const double* __restrict__ us, const double* __restrict__ x0, int nxa, int nxa_ld,
                const double* __restrict__ Qaug, const double* __restrict__ Raug, const double* __restrict__ Maug,
                const double* __restrict__ ulb, const double* __restrict__ uub, double* __restrict__ avg,
                double* __restrict__ out_u, double* __restrict__ out_ell, double* __restrict__ uprev,
                const double* __restrict__ xpl, const double* __restrict__ p, double* __restrict__ Zpl, int pl_ld, int nx,
                int nu, int np) {
  extern __shared__ double sh[];       // xa (nxa) | du (nu)
  double* xa = sh;
  double* du = sh + nxa;
  __shared__ double red[4];
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    for (int j = threadIdx.x; j < nxa; j += blockDim.x) xa[j] = x0[(long long)s * nxa_ld + j];
    for (int j = threadIdx.x; j < nu; j += blockDim.x) {
      const double usj = us[(long long)s * nu + j];
      double u;
      if (kind == 0) u = U[(long long)s * n + j] + usj;                          // first move + target (:689, :661)
      else if (kind == 1) u = uc[(long long)s * nu + j];                         // network output, already us + ... clipped
      else u = fmin(fmax(uc[(long long)s * nu + j] + usj, ulb[j]), uub[j]);      // saturated LQR (controller_evaluation.py:985-987)
      du[j] = u - usj;
      out_u[((long long)s * T + t) * nu + j] = u;
      uprev[(long long)s * nu + j] = u;
      Zpl[(long long)s * pl_ld + nx + j] = u;
    }
    for (int j = threadIdx.x; j < nx; j += blockDim.x) Zpl[(long long)s * pl_ld + j] = xpl[(long long)s * nx + j];
    for (int j = threadIdx.x; j < pl_ld - nx - nu; j += blockDim.x)
      Zpl[(long long)s * pl_ld + nx + nu + j] = j < np ? p[((long long)s * T + t) * np + j] : 0.0;
    __syncthreads();
    // ell = xa'Q xa + du'R du + xa'M du + du'M'xa
    double acc = 0.0;
    for (int i = threadIdx.x; i < nxa; i += blockDim.x) {
      double q = 0.0, m = 0.0;
      for (int j = 0; j < nxa; ++j) q += Qaug[(long long)i * nxa + j] * xa[j];
      for (int j = 0; j < nu; ++j) m += Maug[(long long)i * nu + j] * du[j];
      acc += xa[i] * (q + 2.0 * m);
    }
    for (int i = threadIdx.x; i < nu; i += blockDim.x) {
      double r = 0.0;
      for (int j = 0; j < nu; ++j) r += Raug[(long long)i * nu + j] * du[j];
      acc += du[i] * r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      const double ell = red[0] + red[1] + red[2] + red[3];
      const double a = (avg[s] * (double)t + ell) / (double)(t + 1);
      avg[s] = a;
      if (out_ell) out_ell[(long long)s * T + t] = a;
    }
    __syncthreads();
  }
}

// after the plant GEMM: x+ and the next measurement y+ = C x+ + v  (linearMPC.py:113-121)
__global__ void k_online_next(const double* __restrict__ xy, int S, int T, int t, int nx, int ny, double* __restrict__ xpl,
                              const double* __restrict__ noise, double* __restrict__ out_x, double* __restrict__ out_y) {
  const int w = nx + ny;
  const long long total = (long long)S * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / w;
    const int c = (int)(i - s * w);
    const double v = xy[i];
    if (c < nx) {
      xpl[s * nx + c] = v;
      if (out_x) out_x[(s * (T + 1) + t + 1) * nx + c] = v;
    } else {
      const int k = c - nx;
      out_y[(s * (T + 1) + t + 1) * ny + k] = v + (noise ? noise[(s * (T + 1) + t + 1) * ny + k] : 0.0);
    }
  }
}

}  // namespace nnmpc

using namespace nnmpc;

static int up_pad(double** dst, const double* host, int rows, int cols, int ld) {
  double* tmp = new (std::nothrow) double[(size_t)rows * ld];
  if (!tmp) return set_error(NNMPC_ERR_NOMEM, "out of host memory");
  for (int r = 0; r < rows; ++r) {
    for (int c = 0; c < cols; ++c) tmp[(size_t)r * ld + c] = host[(size_t)r * cols + c];
    for (int c = cols; c < ld; ++c) tmp[(size_t)r * ld + c] = 0.0;
  }
  int rc = upload(dst, tmp, (size_t)rows * ld);
  delete[] tmp;
  return rc;
}

extern "C" {

int nnmpc_online_create(nnmpc_online_t** out, nnmpc_qp_t* qp, nnmpc_ts_t* ts, nnmpc_mlp_t* mlp, int nx, int nu, int ny,
                        int nd, int np, const double* Fkf, const double* Fpl, const double* Kaug, const double* Qaug,
                        const double* Raug, const double* Maug, const double* ulb, const double* uub,
                        const double* xscale, int device) {
  if (!out || !ts || !Fkf || !Fpl || !Qaug || !Raug || !Maug || !ulb || !uub)
    return set_error(NNMPC_ERR_BADARG, "nnmpc_online_create: null argument");
  if (ts->nx != nx || ts->nu != nu || ts->ny != ny || ts->nd != nd || ts->device != device)
    return set_error(NNMPC_ERR_BADARG, "nnmpc_online_create: target selector does not match the sizes / device");
  if (qp && (qp->nu != nu || qp->nxa < nx + nu || qp->device != device))
    return set_error(NNMPC_ERR_BADARG, "nnmpc_online_create: regulator does not match the sizes / device");
  if (mlp) {
    int mx, mu, wu;
    const int mdev = mlp_dims(mlp, &mx, &mu, &wu);
    if (mx != nx || mu != nu || mdev != device)
      return set_error(NNMPC_ERR_BADARG, "nnmpc_online_create: network does not match the sizes / device");
  }
  DeviceGuard dg(device);
  nnmpc_online* h = new (std::nothrow) nnmpc_online();
  if (!h) return set_error(NNMPC_ERR_NOMEM, "out of host memory");
  h->qp = qp; h->ts = ts; h->mlp = mlp;
  h->nx = nx; h->nu = nu; h->ny = ny; h->nd = nd; h->np = np; h->device = device;
  h->nxa = nx + nu;
  h->nxa_ld = qp ? qp->nxa : ((nx + nu + 1) & ~1);
  h->kf_ld = (nx + nd + nu + ny + 1) & ~1;
  h->pl_ld = (nx + nu + np + 1) & ~1;
  h->Fkf = h->Fpl = h->Kaug = h->Qaug = h->Raug = h->Maug = h->ulb = h->uub = h->xscale = nullptr;
  h->fail = nullptr;
  int rc = up_pad(&h->Fkf, Fkf, nx + nd, nx + nd + nu + ny, h->kf_ld);
  if (rc == 0) rc = up_pad(&h->Fpl, Fpl, nx + ny, nx + nu + np, h->pl_ld);
  if (rc == 0 && Kaug) rc = up_pad(&h->Kaug, Kaug, nu, nx + nu, h->nxa_ld);
  if (rc == 0) rc = upload(&h->Qaug, Qaug, (size_t)h->nxa * h->nxa);
  if (rc == 0) rc = upload(&h->Raug, Raug, (size_t)nu * nu);
  if (rc == 0) rc = upload(&h->Maug, Maug, (size_t)h->nxa * nu);
  if (rc == 0) rc = upload(&h->ulb, ulb, (size_t)nu);
  if (rc == 0) rc = upload(&h->uub, uub, (size_t)nu);
  if (rc == 0 && xscale) rc = upload(&h->xscale, xscale, (size_t)nx);
  if (rc == 0 && cudaMalloc((void**)&h->fail, sizeof(int)) != cudaSuccess)
    rc = set_error(NNMPC_ERR_NOMEM, "nnmpc_online_create: cudaMalloc failed");
  if (rc < 0) {
    nnmpc_online_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

int nnmpc_online_destroy(nnmpc_online_t* h) {
  if (!h) return 0;
  DeviceGuard dg(h->device);
  for (double* p : {h->Fkf, h->Fpl, h->Kaug, h->Qaug, h->Raug, h->Maug, h->ulb, h->uub, h->xscale})
    if (p) cudaFree(p);
  if (h->fail) cudaFree(h->fail);
  for (DevBuf<double>* b : {&h->Zkf, &h->xh, &h->dh, &h->x0, &h->lb, &h->ub, &h->xs, &h->us, &h->U, &h->V, &h->uc, &h->Zpl,
                            &h->xy, &h->usp, &h->dus, &h->avg})
    b->release();
  delete h;
  return 0;
}

int nnmpc_online_run(nnmpc_online_t* h, int kind, int S, int T, double* x_io, double* xhat_io, double* uprev_io,
                     const double* setpoints, const double* disturbances, const double* noise, double* y, double* u,
                     double* x, double* xhat, double* xs_out, double* us_out, double* ell_avg, int* iters, double* kkt,
                     double tol, int max_iter, void* stream) {
  if ((S == 0 || T == 0) && h) return 0;
  if (!h || !x_io || !xhat_io || !uprev_io || !setpoints || !disturbances || !y || !u)
    return set_error(NNMPC_ERR_BADARG, "nnmpc_online_run: null argument");
  if (S < 0 || T < 0) return set_error(NNMPC_ERR_BADARG, "nnmpc_online_run: negative size");
  if (kind == NNMPC_ONLINE_MPC && !h->qp) return set_error(NNMPC_ERR_BADARG, "nnmpc_online_run: no regulator in this handle");
  if (kind == NNMPC_ONLINE_NN && !h->mlp) return set_error(NNMPC_ERR_BADARG, "nnmpc_online_run: no network in this handle");
  if (kind == NNMPC_ONLINE_SATDLQR && !h->Kaug) return set_error(NNMPC_ERR_BADARG, "nnmpc_online_run: no LQR gain in this handle");
  if (kind < 0 || kind > 2) return set_error(NNMPC_ERR_BADARG, "nnmpc_online_run: unknown controller kind %d", kind);
  DeviceGuard dg(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int nx = h->nx, nu = h->nu, ny = h->ny, nd = h->nd, np = h->np, nxa_ld = h->nxa_ld;
  const int n = h->qp ? h->qp->n : 0;
  const size_t s = (size_t)S;
  NNMPC_TRY(h->Zkf.ensure(s * h->kf_ld));
  NNMPC_TRY(h->xh.ensure(s * nx));
  NNMPC_TRY(h->dh.ensure(s * (nd > 0 ? nd : 1)));
  NNMPC_TRY(h->x0.ensure(s * nxa_ld));
  NNMPC_TRY(h->lb.ensure(s * nu));
  NNMPC_TRY(h->ub.ensure(s * nu));
  NNMPC_TRY(h->xs.ensure(s * nx));
  NNMPC_TRY(h->us.ensure(s * nu));
  NNMPC_TRY(h->uc.ensure(s * nu));
  NNMPC_TRY(h->Zpl.ensure(s * h->pl_ld));
  NNMPC_TRY(h->xy.ensure(s * (nx + ny)));
  NNMPC_TRY(h->usp.ensure(s * nu));
  NNMPC_TRY(h->dus.ensure(s * nu));
  NNMPC_TRY(h->avg.ensure(s));
  if (kind == NNMPC_ONLINE_MPC) {
    NNMPC_TRY(h->U.ensure(s * n));
    NNMPC_TRY(h->V.ensure(s * n));
  }
  NNMPC_CUDA(cudaMemsetAsync(h->avg.p, 0, s * sizeof(double), st));
  NNMPC_CUDA(cudaMemsetAsync(h->usp.p, 0, s * nu * sizeof(double), st));
  NNMPC_CUDA(cudaMemsetAsync(h->fail, 0, sizeof(int), st));
  // xhat_io = [xhat | dhat] per scenario -> split buffers
  NNMPC_CUDA(cudaMemcpy2DAsync(h->xh.p, (size_t)nx * 8, xhat_io, (size_t)(nx + nd) * 8, (size_t)nx * 8, s, cudaMemcpyDeviceToDevice, st));
  if (nd > 0)
    NNMPC_CUDA(cudaMemcpy2DAsync(h->dh.p, (size_t)nd * 8, xhat_io + nx, (size_t)(nx + nd) * 8, (size_t)nd * 8, s, cudaMemcpyDeviceToDevice, st));
  if (x) NNMPC_CUDA(cudaMemcpy2DAsync(x, (size_t)(T + 1) * nx * 8, x_io, (size_t)nx * 8, (size_t)nx * 8, s, cudaMemcpyDeviceToDevice, st));
  const unsigned eg = (unsigned)((s * h->kf_ld + 255) / 256 < 148 * 8 ? (s * h->kf_ld + 255) / 256 : 148 * 8);
  int warn = 0;
  for (int t = 0; t < T; ++t) {
    // filter (linearMPC.py:159-168) on the measurement y[:, t]
    k_online_pack_kf<<<eg, 256, 0, st>>>(h->Zkf.p, h->kf_ld, h->xh.p, h->dh.p, uprev_io, y + (size_t)t * ny, (long long)(T + 1) * ny,
                                         S, nx, nd, nu, ny);
    count_launch();
    GemmOperands g{};
    g.A = h->Zkf.p; g.lda = h->kf_ld; g.Bt = h->Fkf; g.ldb = h->kf_ld; g.M = S; g.N = nx; g.K = h->kf_ld;
    NNMPC_TRY(gemm_auto<EpiStore>(g, EpiStore::Params{h->xh.p, nx, nullptr, 0}, st));
    if (nd > 0) {
      g.Bt = h->Fkf + (size_t)nx * h->kf_ld; g.N = nd;
      NNMPC_TRY(gemm_auto<EpiStore>(g, EpiStore::Params{h->dh.p, nd, nullptr, 0}, st));
    }
    // targets + regulator inputs (:655-657, :682-688); records xhat and the uprev this step starts from
    TsFused F{};
    F.x = h->xh.p; F.uprev = uprev_io; F.x0 = h->x0.p; F.nxa_ld = nxa_ld; F.lb = h->lb.p; F.ub = h->ub.p;
    F.us_prev = h->usp.p; F.dus = h->dus.p;
    F.row_x = xhat ? xhat + (size_t)t * nx : h->xy.p;            // scratch sink when the record is not wanted
    F.row_stride_x = xhat ? (long long)T * nx : 0;
    F.row_uprev = h->uc.p; F.row_stride_u = 0;
    NNMPC_TRY(ts_solve_device(h->ts, S, setpoints + (size_t)t * ny, (long long)T * ny, h->dh.p, nd, h->xs.p, nx, h->us.p, nu,
                              nullptr, 0, &F, nullptr, h->fail, st));
    if (xs_out) NNMPC_CUDA(cudaMemcpy2DAsync(xs_out + (size_t)t * nx, (size_t)T * nx * 8, h->xs.p, (size_t)nx * 8, (size_t)nx * 8, s, cudaMemcpyDeviceToDevice, st));
    if (us_out) NNMPC_CUDA(cudaMemcpy2DAsync(us_out + (size_t)t * nu, (size_t)T * nu * 8, h->us.p, (size_t)nu * 8, (size_t)nu * 8, s, cudaMemcpyDeviceToDevice, st));
    // controller
    if (kind == NNMPC_ONLINE_MPC) {
      QpOutputs qo{nullptr, kkt ? kkt + t : nullptr, iters ? iters + t : nullptr, T};
      int rc = qp_solve_device(h->qp, S, h->x0.p, h->lb.p, h->ub.p, h->U.p, h->V.p, t > 0 ? 1 : 0, qo, tol, max_iter, st, nullptr);
      if (rc < 0) return rc;
      warn |= rc;
    } else if (kind == NNMPC_ONLINE_NN) {
      NNMPC_TRY(mlp_forward_dispatch(h->mlp, S, h->xh.p, uprev_io, h->xs.p, h->us.p, h->xscale, h->ulb, h->uub, h->uc.p, st));
    } else {
      GemmOperands gk{};
      gk.A = h->x0.p; gk.lda = nxa_ld; gk.Bt = h->Kaug; gk.ldb = nxa_ld; gk.M = S; gk.N = nu; gk.K = nxa_ld;
      NNMPC_TRY(gemm_auto<EpiStore>(gk, EpiStore::Params{h->uc.p, nu, nullptr, 0}, st));
    }
    k_online_finish<<<row_grid(S), 128, (size_t)(h->nxa + nu) * sizeof(double), st>>>(
        kind, S, T, t, n, h->U.p, h->uc.p, h->us.p, h->x0.p, h->nxa, nxa_ld, h->Qaug, h->Raug, h->Maug, h->ulb, h->uub,
        h->avg.p, u, ell_avg, uprev_io, x_io, disturbances, h->Zpl.p, h->pl_ld, nx, nu, np);
    count_launch();
    // plant step and next measurement (:110-121)
    GemmOperands gp{};
    gp.A = h->Zpl.p; gp.lda = h->pl_ld; gp.Bt = h->Fpl; gp.ldb = h->pl_ld; gp.M = S; gp.N = nx + ny; gp.K = h->pl_ld;
    NNMPC_TRY(gemm_auto<EpiStore>(gp, EpiStore::Params{h->xy.p, nx + ny, nullptr, 0}, st));
    k_online_next<<<eg, 256, 0, st>>>(h->xy.p, S, T, t, nx, ny, x_io, noise, x, y);
    count_launch();
  }
  // hand the estimator state back: [xhat | dhat]
  NNMPC_CUDA(cudaMemcpy2DAsync(xhat_io, (size_t)(nx + nd) * 8, h->xh.p, (size_t)nx * 8, (size_t)nx * 8, s, cudaMemcpyDeviceToDevice, st));
  if (nd > 0)
    NNMPC_CUDA(cudaMemcpy2DAsync(xhat_io + nx, (size_t)(nx + nd) * 8, h->dh.p, (size_t)nd * 8, (size_t)nd * 8, s, cudaMemcpyDeviceToDevice, st));
  int failed = 0;
  NNMPC_CUDA(cudaMemcpyAsync(&failed, h->fail, sizeof(int), cudaMemcpyDeviceToHost, st));
  NNMPC_CUDA(cudaStreamSynchronize(st));
  NNMPC_CUDA(cudaGetLastError());
  if (failed) warn |= NNMPC_WARN_TARGET;
  return warn;
}

}  // extern "C"
