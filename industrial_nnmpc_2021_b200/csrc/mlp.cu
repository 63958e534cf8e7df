// Batched forward of the structured regulator network.
//
// Replaces RegulatorLayerWithUprev.call / RegulatorLayerWithoutUprev.call
// (/root/reference/lib/LinearMPCLayers.py:40-61, :91-112; Keras, float64) and the NumPy
// deployment form NeuralNetworkController._get_control_input
// (/root/reference/lib/controller_evaluation.py:863-892):
//     u = us + f(x, [uprev], xs, us) - f(xs, [us], xs, us),   f = (Dense+ReLU)^(L-1), Dense(no bias)
// Both passes of a sample are interleaved as rows 2b / 2b+1 of one activation matrix, so every
// layer is a single FP64 tensor-core GEMM with bias+ReLU fused into the epilogue, and the last
// layer's epilogue forms us + out(2b) - out(2b+1) (+ clip) with one warp shuffle - the difference
// never goes through memory.
#include "qp.cuh"

struct nnmpc_mlp {
  int nx, nu, with_uprev, L, device;
  int dims[17];
  int ld[17];          // even leading dimensions of the activations (ld[i] >= dims[i])
  double* Wt[16];      // device, transposed weights: dims[i+1] x ld[i]
  double* bias[16];    // device, dims[i+1] (null for the last layer)
  int maxw;            // widest hidden activation (ld)
  nnmpc::DevBuf<double> act0, act1;
  nnmpc::DevBuf<double> hx, hup, hxs, hus, hout, hscale, hlb, hub;
};

namespace nnmpc {

// rows 2b and 2b+1 of the first-layer input: [x/s, (uprev), xs/s, us] and [xs/s, (us), xs/s, us]
__global__ void k_pack_inputs(const double* __restrict__ x, const double* __restrict__ uprev,
                              const double* __restrict__ xs, const double* __restrict__ us,
                              const double* __restrict__ xscale, double* __restrict__ in, long long B, int nx,
                              int nu, int with_uprev, int ld) {
  const long long total = B * ld;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int o1 = nx, o2 = nx + (with_uprev ? nu : 0), o3 = o2 + nx, o4 = o3 + nu;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long b = i / ld;
    const int c = (int)(i - b * ld);
    double v1, v2;
    if (c < o1) {
      v1 = x[b * nx + c];
      v2 = xs[b * nx + c];
      if (xscale) { v1 /= xscale[c]; v2 /= xscale[c]; }
    } else if (c < o2) {
      v1 = uprev[b * nu + (c - o1)];
      v2 = us[b * nu + (c - o1)];
    } else if (c < o3) {
      v1 = xs[b * nx + (c - o2)];
      if (xscale) v1 /= xscale[c - o2];
      v2 = v1;
    } else if (c < o4) {
      v1 = us[b * nu + (c - o3)];
      v2 = v1;
    } else {
      v1 = v2 = 0.0;
    }
    in[(2 * b) * ld + c] = v1;
    in[(2 * b + 1) * ld + c] = v2;
  }
}

// last layer: out[b] = us[b] + acc(row 2b) - acc(row 2b+1), optional clip to [ulb, uub].
// Fragment rows g and g^1 sit in lanes 4 apart, tiles start at multiples of 8 rows, so a pair
// never straddles a fragment; every lane calls apply (ok flags predicate the store).
struct EpiStructOut {
  struct Params {
    double* out;        // B x nu
    const double* us;   // B x nu
    const double* ulb;  // nu or null
    const double* uub;
    int nu;
  };
  Params p;
  __device__ EpiStructOut(const Params& p_, int, int) : p(p_) {}
  __device__ void begin_row() {}
  __device__ void apply(int pr, int, int col, double a0, double a1, bool ok0, bool ok1) {
    const double b0 = __shfl_xor_sync(0xffffffffu, a0, 4);
    const double b1 = __shfl_xor_sync(0xffffffffu, a1, 4);
    if ((pr & 1) || !ok0) return;
    const long long b = pr >> 1;
    double v0 = p.us[b * p.nu + col] + (a0 - b0);   // us + (out1 - out2), the reference order (:58-60)
    if (p.ulb) v0 = fmin(fmax(v0, p.ulb[col]), p.uub[col]);
    p.out[b * p.nu + col] = v0;
    if (ok1) {
      double v1 = p.us[b * p.nu + col + 1] + (a1 - b1);
      if (p.ulb) v1 = fmin(fmax(v1, p.ulb[col + 1]), p.uub[col + 1]);
      p.out[b * p.nu + col + 1] = v1;
    }
  }
  __device__ void finish_row(int, int, int, bool) {}
};

static int mlp_forward_device(nnmpc_mlp* h, long long B, const double* x, const double* uprev, const double* xs,
                              const double* us, const double* xscale, const double* ulb, const double* uub,
                              double* out, cudaStream_t st) {
  if (B <= 0) return 0;
  // chunk the batch so the two activation buffers stay below ~2 GiB each
  long long chunk = (long long)(1ll << 28) / (2ll * h->maxw);
  chunk = chunk < 1024 ? 1024 : (chunk > (1 << 20) ? (1 << 20) : chunk);
  if (chunk > B) chunk = B;
  NNMPC_TRY(h->act0.ensure((size_t)2 * chunk * h->maxw));
  NNMPC_TRY(h->act1.ensure((size_t)2 * chunk * h->maxw));
  const int nx = h->nx, nu = h->nu, L = h->L;
  for (long long b0 = 0; b0 < B; b0 += chunk) {
    const long long nb = B - b0 < chunk ? B - b0 : chunk;
    k_pack_inputs<<<148 * 16, 256, 0, st>>>(x + b0 * nx, uprev ? uprev + b0 * nu : nullptr, xs + b0 * nx,
                                            us + b0 * nu, xscale, h->act0.p, nb, nx, nu, h->with_uprev, h->ld[0]);
    count_launch();
    double* cur = h->act0.p;
    double* nxt = h->act1.p;
    for (int l = 0; l < L; ++l) {
      GemmOperands g{cur, h->ld[l], h->Wt[l], h->ld[l], (int)(2 * nb), h->dims[l + 1], h->ld[l], nullptr, nullptr};
      if (l < L - 1) {
        // zero the padding column of an odd-width hidden layer once per chunk (it is read as K)
        if (h->ld[l + 1] != h->dims[l + 1])
          NNMPC_CUDA(cudaMemsetAsync(nxt, 0, (size_t)2 * nb * h->ld[l + 1] * 8, st));
        NNMPC_TRY(gemm_auto<EpiStore>(g, EpiStore::Params{nxt, h->ld[l + 1], h->bias[l], 1}, st));
        double* t = cur; cur = nxt; nxt = t;
      } else {
        EpiStructOut::Params ep{out + b0 * nu, us + b0 * nu, ulb, uub, nu};
        NNMPC_TRY(gemm_auto<EpiStructOut>(g, ep, st));
      }
    }
  }
  NNMPC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace nnmpc

using namespace nnmpc;

extern "C" {

int nnmpc_mlp_create(nnmpc_mlp_t** out, int nx, int nu, int with_uprev, int num_layers, const int* dims,
                     const double* const* weights_host, const double* const* biases_host, int device) {
  if (!out || !dims || !weights_host) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_create: null argument");
  if (num_layers < 1 || num_layers > 16) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_create: 1..16 layers");
  const int in_w = 2 * nx + (with_uprev ? 2 : 1) * nu;
  if (dims[0] != in_w) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_create: dims[0]=%d, expected %d", dims[0], in_w);
  if (dims[num_layers] != nu) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_create: last width must equal nu");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_error(NNMPC_ERR_CUDA, "nnmpc_mlp_create: no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_create: bad device %d", device);
  DeviceGuard dg(device);
  nnmpc_mlp* h = new (std::nothrow) nnmpc_mlp();
  if (!h) return set_error(NNMPC_ERR_NOMEM, "out of host memory");
  h->nx = nx; h->nu = nu; h->with_uprev = with_uprev; h->L = num_layers; h->device = device;
  h->maxw = 0;
  for (int i = 0; i <= num_layers; ++i) {
    if (dims[i] < 1) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_create: bad width");
    h->dims[i] = dims[i];
    h->ld[i] = (dims[i] + 1) & ~1;
    if (i < num_layers && h->ld[i] > h->maxw) h->maxw = h->ld[i];
  }
  for (int l = 0; l < num_layers; ++l) {
    const int in = dims[l], outw = dims[l + 1], ld = h->ld[l];
    double* tmp = new (std::nothrow) double[(size_t)outw * ld];
    if (!tmp) return set_error(NNMPC_ERR_NOMEM, "out of host memory");
    const double* W = weights_host[l];  // in x out, row-major (Keras kernel layout)
    for (int o = 0; o < outw; ++o) {
      for (int i = 0; i < in; ++i) tmp[(size_t)o * ld + i] = W[(size_t)i * outw + o];
      for (int i = in; i < ld; ++i) tmp[(size_t)o * ld + i] = 0.0;
    }
    int rc = upload(&h->Wt[l], tmp, (size_t)outw * ld);
    delete[] tmp;
    if (rc < 0) return rc;
    h->bias[l] = nullptr;
    if (l < num_layers - 1) {
      if (!biases_host || !biases_host[l]) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_create: missing bias %d", l);
      NNMPC_TRY(upload(&h->bias[l], biases_host[l], (size_t)outw));
    }
  }
  *out = h;
  return 0;
}

int nnmpc_mlp_destroy(nnmpc_mlp_t* h) {
  if (!h) return 0;
  DeviceGuard dg(h->device);
  for (int l = 0; l < h->L; ++l) {
    cudaFree(h->Wt[l]);
    if (h->bias[l]) cudaFree(h->bias[l]);
  }
  h->act0.release(); h->act1.release();
  h->hx.release(); h->hup.release(); h->hxs.release(); h->hus.release(); h->hout.release();
  h->hscale.release(); h->hlb.release(); h->hub.release();
  delete h;
  return 0;
}

int nnmpc_mlp_forward(nnmpc_mlp_t* h, long long B, const double* x, const double* uprev, const double* xs,
                      const double* us, const double* xscale, const double* ulb, const double* uub, double* out,
                      void* stream) {
  if (B == 0 && h) return 0;
  if (!h || !x || !xs || !us || !out) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_forward: null argument");
  if (h->with_uprev && !uprev) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_forward: uprev required");
  if ((ulb == nullptr) != (uub == nullptr)) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_forward: ulb/uub both or none");
  if (B < 0) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_forward: negative batch");
  DeviceGuard dg(h->device);
  return mlp_forward_device(h, B, x, uprev, xs, us, xscale, ulb, uub, out, (cudaStream_t)stream);
}

int nnmpc_mlp_forward_host(nnmpc_mlp_t* h, long long B, const double* x, const double* uprev, const double* xs,
                           const double* us, const double* xscale, const double* ulb, const double* uub, double* out) {
  if (B == 0 && h) return 0;
  if (!h || !x || !xs || !us || !out) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_forward_host: null argument");
  if (h->with_uprev && !uprev) return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_forward_host: uprev required");
  if ((ulb == nullptr) != (uub == nullptr))
    return set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_forward_host: ulb/uub both or none");
  if (B <= 0) return B == 0 ? 0 : set_error(NNMPC_ERR_BADARG, "nnmpc_mlp_forward_host: negative batch");
  DeviceGuard dg(h->device);
  const size_t b = (size_t)B;
  const int nx = h->nx, nu = h->nu;
  NNMPC_TRY(h->hx.ensure(b * nx));
  NNMPC_TRY(h->hxs.ensure(b * nx));
  NNMPC_TRY(h->hus.ensure(b * nu));
  NNMPC_TRY(h->hup.ensure(b * nu));
  NNMPC_TRY(h->hout.ensure(b * nu));
  NNMPC_TRY(h->hscale.ensure(nx));
  NNMPC_TRY(h->hlb.ensure(nu));
  NNMPC_TRY(h->hub.ensure(nu));
  cudaStream_t st = 0;
  NNMPC_CUDA(cudaMemcpyAsync(h->hx.p, x, b * nx * 8, cudaMemcpyHostToDevice, st));
  NNMPC_CUDA(cudaMemcpyAsync(h->hxs.p, xs, b * nx * 8, cudaMemcpyHostToDevice, st));
  NNMPC_CUDA(cudaMemcpyAsync(h->hus.p, us, b * nu * 8, cudaMemcpyHostToDevice, st));
  if (h->with_uprev) NNMPC_CUDA(cudaMemcpyAsync(h->hup.p, uprev, b * nu * 8, cudaMemcpyHostToDevice, st));
  if (xscale) NNMPC_CUDA(cudaMemcpyAsync(h->hscale.p, xscale, (size_t)nx * 8, cudaMemcpyHostToDevice, st));
  if (ulb) {
    NNMPC_CUDA(cudaMemcpyAsync(h->hlb.p, ulb, (size_t)nu * 8, cudaMemcpyHostToDevice, st));
    NNMPC_CUDA(cudaMemcpyAsync(h->hub.p, uub, (size_t)nu * 8, cudaMemcpyHostToDevice, st));
  }
  NNMPC_TRY(mlp_forward_device(h, B, h->hx.p, h->with_uprev ? h->hup.p : nullptr, h->hxs.p, h->hus.p,
                               xscale ? h->hscale.p : nullptr, ulb ? h->hlb.p : nullptr, ulb ? h->hub.p : nullptr,
                               h->hout.p, st));
  NNMPC_CUDA(cudaMemcpyAsync(out, h->hout.p, b * nu * 8, cudaMemcpyDeviceToHost, st));
  NNMPC_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
