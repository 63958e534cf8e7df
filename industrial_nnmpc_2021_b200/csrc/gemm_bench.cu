// Tile-shape tuning harness for the FP64 tensor-core GEMM (not on the product path): times the
// same C = A Bt' launch under several CTA tile configurations so one GPU call ranks them.
#include "qp.cuh"

namespace nnmpc {
template <class T>
static int bench_one(const GemmOperands& g, double* C, int reps, float* ms, cudaStream_t st) {
  cudaEvent_t a, b;
  NNMPC_CUDA(cudaEventCreate(&a));
  NNMPC_CUDA(cudaEventCreate(&b));
  EpiStore::Params ep{C, g.N, nullptr, 0};
  cudaError_t ce = cudaSuccess;
  for (int i = 0; i < 2 && ce == cudaSuccess; ++i) ce = launch_gemm<T, EpiStore>(g, ep, st);
  NNMPC_CUDA(ce);
  NNMPC_CUDA(cudaEventRecord(a, st));
  for (int i = 0; i < reps && ce == cudaSuccess; ++i) ce = launch_gemm<T, EpiStore>(g, ep, st);
  NNMPC_CUDA(ce);
  NNMPC_CUDA(cudaEventRecord(b, st));
  NNMPC_CUDA(cudaEventSynchronize(b));
  NNMPC_CUDA(cudaEventElapsedTime(ms, a, b));
  *ms /= reps;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return 0;
}
}  // namespace nnmpc
using namespace nnmpc;

extern "C" int nnmpc_gemm_bench(int cfg, int M, int N, int K, const double* A, const double* Bt, double* C, int reps,
                                float* ms) {
  if (!A || !Bt || !C || !ms || (K & 1)) return set_error(NNMPC_ERR_BADARG, "nnmpc_gemm_bench: bad argument");
  GemmOperands g{A, K, Bt, K, M, N, K, nullptr, nullptr};
  cudaStream_t st = 0;
  switch (cfg) {
    case 0: return bench_one<GemmTile<128, 128, 2, 4, 4, 16>>(g, C, reps, ms, st);   // current
    case 1: return bench_one<GemmTile<128, 128, 4, 4, 4, 16>>(g, C, reps, ms, st);   // 16 warps
    case 2: return bench_one<GemmTile<128, 128, 2, 4, 3, 32>>(g, C, reps, ms, st);   // deeper k tile
    case 3: return bench_one<GemmTile<128, 256, 2, 4, 3, 16>>(g, C, reps, ms, st);   // wide tile
    case 4: return bench_one<GemmTile<256, 128, 4, 2, 3, 16>>(g, C, reps, ms, st);   // tall tile
    case 5: return bench_one<GemmTile<128, 128, 4, 4, 3, 32>>(g, C, reps, ms, st);   // 16 warps, deep k
    case 6: return bench_one<GemmTile<64, 64, 2, 2, 4, 16>>(g, C, reps, ms, st);
    case 7: return bench_one<GemmTile<32, 32, 2, 2, 6, 16>>(g, C, reps, ms, st);
    case 8: return bench_one<GemmTile<64, 128, 2, 4, 4, 16>>(g, C, reps, ms, st);
    case 9: return bench_one<GemmTile<32, 64, 2, 2, 6, 16>>(g, C, reps, ms, st);
    case 10: return bench_one<GemmTile<16, 32, 1, 4, 8, 16>>(g, C, reps, ms, st);
    case 11: return bench_one<GemmTile<128, 128, 2, 4, 2, 32>>(g, C, reps, ms, st);
    default: return set_error(NNMPC_ERR_BADARG, "nnmpc_gemm_bench: unknown config %d", cfg);
  }
}
