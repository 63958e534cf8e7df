// Mixed-precision Douglas-Rachford iteration of the regulator QP: epilogues for the tcgen05
// split-operator GEMM (lp_gemm.cuh) and the small FP64 kernels around the FP64 anchors.
//
// State per sample (row), all FP64 unless noted:
//   V     Douglas-Rachford state v                      X     x = Top w_lp - c  (tracked incrementally)
//   E     fp32: w - w_lp, the part of the operand the fp16 increments have not delivered yet
//   D     fp16: scaled increment s_row (w_lp+ - w_lp) = the A operand of the next tensor-core pass
// One pass:  x += (T1 + T2) dq / (s_T s_row);  d = x - clip(v);  v += alpha d;
//            dw = (2 clip(v) - v) - w_lp;  dq = fp16(s_row' dw);  w_lp += dq / s_row'
// Fixed points are exactly those of the FP64 iteration in qp.cu (E -> 0, dq -> 0).  What the fp16
// operator split and the fp32 accumulation lose is proportional to |w_lp - w_anchor|; an FP64
// anchor GEMM (x = Top w - c with w_lp := w, E := 0) resets it, and every returned point is
// checked with P in FP64, so the reference tolerances (/root/repo/BASELINE.json north_star:
// KKT <= 1e-8) are certified by FP64 arithmetic, not by the fp16 path.
#pragma once
#include "lp.cuh"
#include "lp_gemm.cuh"

namespace nnmpc {

using LpTileN128 = lp::LpTile<128, 4>;

// one element of the Douglas-Rachford delta update (shared by the tensor-core epilogue and k_dr_first)
__device__ __forceinline__ void dr_delta_one(double& x, double& v, double wl, double l, double u, double alpha,
                                             double& dw, double& dabs) {
  const double z0 = clipd(v, l, u);
  const double d = x - z0;
  v += alpha * d;
  const double z1 = clipd(v, l, u);
  dw = (2.0 * z1 - v) - wl;
  dabs = fabs(d);
}
__device__ __forceinline__ __half quantise_dw(double dw, double s, double inv_s, float& e) {
  double t = dw * s;
  t = fmin(fmax(t, -60000.0), 60000.0);          // saturate: the residual e carries what did not fit
  const __half q = __double2half(t);
  e = (float)(dw - (double)__half2float(q) * inv_s);
  return q;
}

// ---- epilogue of the tensor-core pass: the whole iteration on the accumulator registers ----------
struct EpiDelta {
  struct Params {
    double* X;
    double* V;
    float* E;
    __half* Dn;          // next operand (the TMA reads the current one)
    long long ldd;       // row stride of Dn (elements)
    const double* lb;
    const double* ub;    // B x nu
    const int* state;    // row takes part iff state[row] == iter_state
    int iter_state;
    const int* list_r;   // position (row of the operand the TMA reads) -> sample row; null = identity
    const int* pos_w;    // sample row -> position in the operand written now; null = identity
    const double* sc_in;   // scale the current operand row was quantised with
    const double* sc_out;  // scale for the operand written now
    unsigned long long* dres;  // per row: max |d| (bit pattern of a non-negative double)
    int n, nu;
    double alpha;
    double inv_sT;       // 1 / operator scale
  };
  Params p;
  bool ok;
  int row;
  long long pw;
  double inv_in, s_out, inv_out, dmax;
  const double* lbr;
  const double* ubr;
  __device__ explicit EpiDelta(const Params& p_) : p(p_), ok(false), row(0), pw(0), inv_in(0), s_out(0), inv_out(0), dmax(0), lbr(nullptr), ubr(nullptr) {}
  __device__ void begin_row(int pos, bool in_range) {
    row = in_range ? (p.list_r ? p.list_r[pos] : pos) : 0;
    ok = in_range && p.state[row] == p.iter_state;
    dmax = 0.0;
    if (ok) {
      pw = p.pos_w ? p.pos_w[row] : row;
      inv_in = p.inv_sT / p.sc_in[row];
      s_out = p.sc_out[row];
      inv_out = 1.0 / s_out;
      lbr = p.lb + (long long)row * p.nu;
      ubr = p.ub + (long long)row * p.nu;
    }
  }
  // 16 consecutive columns of this thread's row.  Fast path (stage width a multiple of 16, full chunk): every
  // state load of the chunk is issued before the first use and results are stored at the end, so each
  // epilogue thread keeps 24 x 16 B of HBM traffic in flight.
  __device__ void chunk(int col0, const uint32_t (&acc)[16], int N) {
    if (!ok || col0 >= N) return;
    const long long base = (long long)row * p.n + col0;
    __half* dn = p.Dn + pw * p.ldd + col0;
    const int k0 = col0 % p.nu;
    if ((p.nu & 15) == 0 && col0 + 16 <= N) {
      double2 x[8], v[8];
      float2 e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[i] = *reinterpret_cast<const double2*>(p.X + base + 2 * i);
        v[i] = *reinterpret_cast<const double2*>(p.V + base + 2 * i);
        e[i] = *reinterpret_cast<const float2*>(p.E + base + 2 * i);
      }
      __half2 q[8];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        // the stage bounds of these 8 columns: L1/L2 hits (every column tile re-reads the same 2 x 256 B per row),
        // loaded half a chunk at a time to stay inside the 168 registers ten warps leave per thread
        double2 lbc[4], ubc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          lbc[i] = *reinterpret_cast<const double2*>(lbr + k0 + 8 * hh + 2 * i);
          ubc[i] = *reinterpret_cast<const double2*>(ubr + k0 + 8 * hh + 2 * i);
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int i = 4 * hh + ii;
          x[i].x += (double)__uint_as_float(acc[2 * i]) * inv_in;
          x[i].y += (double)__uint_as_float(acc[2 * i + 1]) * inv_in;
          const double wl0 = (2.0 * clipd(v[i].x, lbc[ii].x, ubc[ii].x) - v[i].x) - (double)e[i].x;
          const double wl1 = (2.0 * clipd(v[i].y, lbc[ii].y, ubc[ii].y) - v[i].y) - (double)e[i].y;
          double dw0, dw1, a0, a1;
          dr_delta_one(x[i].x, v[i].x, wl0, lbc[ii].x, ubc[ii].x, p.alpha, dw0, a0);
          dr_delta_one(x[i].y, v[i].y, wl1, lbc[ii].y, ubc[ii].y, p.alpha, dw1, a1);
          const __half q0 = quantise_dw(dw0, s_out, inv_out, e[i].x);
          const __half q1 = quantise_dw(dw1, s_out, inv_out, e[i].y);
          q[i] = __halves2half2(q0, q1);
          const double a = (a0 <= a1) ? a1 : a0;       // NaN propagates
          dmax = (a <= dmax) ? dmax : a;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        *reinterpret_cast<double2*>(p.X + base + 2 * i) = x[i];
        *reinterpret_cast<double2*>(p.V + base + 2 * i) = v[i];
        *reinterpret_cast<float2*>(p.E + base + 2 * i) = e[i];
      }
      uint4* d4 = reinterpret_cast<uint4*>(dn);
      d4[0] = make_uint4(*reinterpret_cast<uint32_t*>(&q[0]), *reinterpret_cast<uint32_t*>(&q[1]),
                         *reinterpret_cast<uint32_t*>(&q[2]), *reinterpret_cast<uint32_t*>(&q[3]));
      d4[1] = make_uint4(*reinterpret_cast<uint32_t*>(&q[4]), *reinterpret_cast<uint32_t*>(&q[5]),
                         *reinterpret_cast<uint32_t*>(&q[6]), *reinterpret_cast<uint32_t*>(&q[7]));
      return;
    }
    // general path (any stage width, ragged last chunk): element pairs, n is even
    int k = k0;
#pragma unroll 1
    for (int j = 0; j < 16; j += 2) {
      const int k1 = (k + 1 == p.nu) ? 0 : k + 1;
      if (col0 + j < N) {
        double2 x = *reinterpret_cast<const double2*>(p.X + base + j);
        double2 v = *reinterpret_cast<const double2*>(p.V + base + j);
        float2 e = *reinterpret_cast<const float2*>(p.E + base + j);
        const double l0 = lbr[k], u0 = ubr[k], l1 = lbr[k1], u1 = ubr[k1];
        x.x += (double)__uint_as_float(acc[j]) * inv_in;
        x.y += (double)__uint_as_float(acc[j + 1]) * inv_in;
        const double wl0 = (2.0 * clipd(v.x, l0, u0) - v.x) - (double)e.x;
        const double wl1 = (2.0 * clipd(v.y, l1, u1) - v.y) - (double)e.y;
        double dw0, dw1, a0, a1;
        dr_delta_one(x.x, v.x, wl0, l0, u0, p.alpha, dw0, a0);
        dr_delta_one(x.y, v.y, wl1, l1, u1, p.alpha, dw1, a1);
        const __half q0 = quantise_dw(dw0, s_out, inv_out, e.x);
        const __half q1 = quantise_dw(dw1, s_out, inv_out, e.y);
        *reinterpret_cast<double2*>(p.X + base + j) = x;
        *reinterpret_cast<double2*>(p.V + base + j) = v;
        *reinterpret_cast<float2*>(p.E + base + j) = e;
        *reinterpret_cast<__half2*>(dn + j) = __halves2half2(q0, q1);
        const double a = (a0 <= a1) ? a1 : a0;
        dmax = (a <= dmax) ? dmax : a;
      }
      k = (k1 + 1 == p.nu) ? 0 : k1 + 1;
    }
  }
  __device__ void end_row() {
    if (!ok) return;
    double m = dmax;
    if (!(m <= 1.7e308)) m = __longlong_as_double(0x7ff0000000000000ll);   // NaN/Inf must not look converged
    atomicMax(p.dres + row, (unsigned long long)__double_as_longlong(m));
  }
};

// plain store (self test): C = scale * acc
struct EpiLpStore {
  struct Params {
    double* C;
    long long ldc;
    double scale;
  };
  Params p;
  bool ok;
  int row;
  __device__ explicit EpiLpStore(const Params& p_) : p(p_), ok(false), row(0) {}
  __device__ void begin_row(int pos, bool in_range) { ok = in_range; row = pos; }
  __device__ void chunk(int col0, const uint32_t (&acc)[16], int N) {
    if (!ok) return;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col0 + j < N) p.C[(long long)row * p.ldc + col0 + j] = (double)__uint_as_float(acc[j]) * p.scale;
  }
  __device__ void end_row() {}
};

// ---- FP64 anchor: x = Top w - c on the accumulators of the DMMA GEMM ---------------------------
struct EpiAnchor {
  struct Params {
    double* X;
    const double* C;
    int n;
  };
  Params p;
  __device__ EpiAnchor(const Params& p_, int, int) : p(p_) {}
  __device__ void begin_row() {}
  __device__ void apply(int pr, int, int col, double a0, double a1, bool ok0, bool ok1) {
    if (!ok0) return;
    const long long off = (long long)pr * p.n + col;
    if (ok1) {
      const double2 c = *reinterpret_cast<const double2*>(p.C + off);
      *reinterpret_cast<double2*>(p.X + off) = make_double2(a0 - c.x, a1 - c.y);
    } else {
      p.X[off] = a0 - p.C[off];
    }
  }
  __device__ void finish_row(int, int, int, bool) {}
};

// operator split: T1 = fp16(s Top), T2 = fp16(s Top - T1)  (same scale: both products share one accumulator)
__global__ void k_split_f16(const double* __restrict__ T, int n, long long ldh, double s, __half* __restrict__ T1,
                            __half* __restrict__ T2) {
  const long long total = (long long)n * ldh;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / ldh;
    const int c = (int)(i - r * ldh);
    double t = c < n ? T[r * n + c] * s : 0.0;
    const __half h1 = __double2half(t);
    const __half h2 = __double2half(t - (double)__half2float(h1));
    T1[i] = h1;
    T2[i] = h2;
  }
}

// anchor rows: the FP64 operand w = 2 clip(v) - v, nothing outstanding (E = 0)
__global__ void k_anchor_prep(const int* __restrict__ rows, const int* __restrict__ count, const double* __restrict__ V,
                              double* __restrict__ W, float* __restrict__ E, const double* __restrict__ lb,
                              const double* __restrict__ ub, int n, int nu) {
  if ((int)blockIdx.x >= *count) return;
  const long long s = rows[blockIdx.x];
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int k = j % nu;
    const double v = V[s * n + j];
    W[s * n + j] = 2.0 * clipd(v, lb[s * nu + k], ub[s * nu + k]) - v;
    E[s * n + j] = 0.f;
  }
}

// anchor rows after the FP64 GEMM: one full-precision Douglas-Rachford step from the exact x, then the
// first fp16 increment with a scale taken from the row's own max |dw| (one CTA per row)
__global__ void __launch_bounds__(256)
k_dr_first(const int* __restrict__ rows, const int* __restrict__ count, double* __restrict__ X, double* __restrict__ V,
           double* __restrict__ W, float* __restrict__ E, __half* __restrict__ D, long long ldd,
           const double* __restrict__ lb, const double* __restrict__ ub, double* __restrict__ sc_in,
           double* __restrict__ sc_out, int* __restrict__ state, int* __restrict__ it, int iter_state, int n, int nu,
           double alpha, const int* __restrict__ pos_r) {
  if ((int)blockIdx.x >= *count) return;
  const long long s = rows[blockIdx.x];
  __shared__ double red[2][8];
  double dmax = 0.0, wmax = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int k = j % nu;
    const double l = lb[s * nu + k], u = ub[s * nu + k];
    double x = X[s * n + j], v = V[s * n + j], dw, a;
    dr_delta_one(x, v, W[s * n + j], l, u, alpha, dw, a);
    V[s * n + j] = v;
    W[s * n + j] = dw;                    // staged for the second pass
    dmax = (a <= dmax) ? dmax : a;
    const double b = fabs(dw);
    wmax = (b <= wmax) ? wmax : b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double a = __shfl_xor_sync(0xffffffffu, dmax, o), b = __shfl_xor_sync(0xffffffffu, wmax, o);
    dmax = (a <= dmax) ? dmax : a;
    wmax = (b <= wmax) ? wmax : b;
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = dmax;
    red[1][threadIdx.x >> 5] = wmax;
  }
  __syncthreads();
  dmax = red[0][0];
  wmax = red[1][0];
  for (int w = 1; w < 8; ++w) {
    dmax = (red[0][w] <= dmax) ? dmax : red[0][w];
    wmax = (red[1][w] <= wmax) ? wmax : red[1][w];
  }
  const double sq = pow2_scale(wmax), inv = 1.0 / sq;
  const long long dpos = pos_r ? pos_r[s] : s;      // row of the operand buffer the next pass reads for this sample
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float e;
    D[dpos * ldd + j] = quantise_dw(W[s * n + j], sq, inv, e);
    E[s * n + j] = e;
  }
  if (threadIdx.x == 0) {
    sc_in[s] = sq;
    sc_out[s] = pow2_scale(3.0 * alpha * dmax);
    state[s] = iter_state;
    it[s] += 1;
  }
}

}  // namespace nnmpc
