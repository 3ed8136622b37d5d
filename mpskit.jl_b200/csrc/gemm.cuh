// Grouped FP64 / ComplexF64 "NT" GEMM engine for sm_100a.
//   C[M x N] (+)= A[M x K] * B[N x K]^T     all row-major, A and B K-contiguous
// This is the one tensor-core kernel behind every contraction on the hot path (SURVEY.md 8a rows
// a2-a5, a10-a11: derivatives.jl:77-197, transfer.jl:105-110,166-259).  Blackwell has no FP64
// tcgen05 kind; FP64 tensor math is DMMA.8x8x4 (mma.sync.m8n8k4.f64), measured 37.0 TFLOP/s on
// B200 (profiles/r01_peak_fp64_microbench.json).  Operand tiles are staged by TMA
// (cp.async.bulk.tensor.2d) into a multi-stage shared-memory ring guarded by mbarriers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpsk {

struct alignas(128) GemmProblemDev {
  CUtensorMap tmA;      // rows = M, inner = K (doubles: K or 2K)
  CUtensorMap tmB;      // rows = N
  void* C;
  const void* A;        // raw pointers (used by the debug kernel)
  const void* B;
  int64_t lda, ldb, ldc;  // in elements
  int32_t M, N, K;
  int32_t accumulate;   // 0: C = AB^T, 1: C += AB^T
};

struct GemmTile { int32_t prob, m0, n0, pad; };

struct GemmProblemHost {
  const void* A; const void* B; void* C;
  int64_t lda, ldb, ldc; int32_t M, N, K; int32_t accumulate;
};

// Tile configurations.  Warp tile: 32 x WN (complex) / 64 x WN (real); CTA tile = BM x (WN * WARPS_N).
//   wide     : 8 warps, 128 x 64 complex / 128 x 128 real, 1 CTA per SM, 4 / 3 pipeline stages
//   narrow   : 4 warps, 128 x 32 complex, 2 CTAs per SM, 2 stages -- for plans dominated by narrow-N problems (the row
//              panels of a sharded matvec), where 64-wide tiles would waste up to 40 % of the DMMA work
//   narrow16 : 4 warps, 128 x 16 complex (warp tile 32 x 16), 3 CTAs per SM -- row panels whose width is far from a
//              multiple of 32 (136 rows per sector at 8 GPUs: 144 instead of 160 padded columns)
//   3M (M3_) : the same three complex tile shapes with the 3-multiplication complex product (Karatsuba / "zgemm3m"):
//              P1 = Ar Br^T, P2 = Ai Bi^T, P3 = (Ar + Ai)(Br + Bi)^T accumulate in three register sets (192 accumulator
//              registers for the 32 x 32 warp tile) and are combined in the epilogue -- 3 DMMAs per complex k4 step instead
//              of 4, i.e. 25 % less work on the pipe that bounds the kernel.  Default for ComplexF64; MPSK_FLAG_GEMM_4M
//              selects the conventional product.
template <bool CPLX, int WARPS_N_, int WN_ = 32, bool M3_ = false> struct GemmCfg {
  static constexpr bool M3 = CPLX && M3_;
  static constexpr int WARPS_M = CPLX ? 4 : 2;
  static constexpr int WARPS_N = WARPS_N_;
  static constexpr int WN = WN_;
  static constexpr int BM = 128, BN = WN_ * WARPS_N_;
  static constexpr int KSUB = CPLX ? 4 : 8, NSUB = 4;
  static constexpr int NTHREADS = 32 * WARPS_M * WARPS_N;
  static constexpr int CTAS_PER_SM = NTHREADS <= 128 ? (WN_ <= 16 ? 3 : 2) : 1;
  static constexpr int STAGES = CTAS_PER_SM >= 2 ? 2 : (CPLX ? 4 : 3);
};
constexpr int GEMM_WIDE_N_CPLX = 2, GEMM_NARROW_N_CPLX = 1, GEMM_WIDE_N_REAL = 4;
// tile configuration ids of a plan (GemmPlan::cfg); +4 = the 3M variant of a complex configuration
enum { TC_REAL = 0, TC_WIDE = 1, TC_NARROW = 2, TC_NARROW16 = 3, TC_3M = 4, TC_NCFG = 8 };

}  // namespace mpsk
