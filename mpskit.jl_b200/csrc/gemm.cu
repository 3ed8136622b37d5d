#include "gemm.cuh"
#include "internal.h"
#include <cstdlib>

namespace mpsk {

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// explicit shared-space 128-bit load: the ring base is re-aligned at run time, which makes the compiler lose the address
// space and emit generic LD.E.128 with 64-bit addresses; ld.shared gives LDS.128 on a 32-bit address
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------ kernel
// 8 DMMA warps, 256 threads, one persistent CTA per SM: CTA b handles tiles b, b+grid, ...
// Shared ring: per stage NSUB sub-tiles per operand, each sub-tile [rows][64 B] dense so that the
// LDS.128 fragment loads (lane (g,t): row g, 16-byte chunk t) are bank-conflict free.
// Lane 0 of warp 0 doubles as the TMA producer: before consuming k-block `it` it refills the stage
// that was consumed at `it-1` with k-block `it+STAGES-1` (the flat k-block sequence runs across tile
// boundaries, so the pipeline never drains between tiles).  A dedicated producer warp would cost a
// whole extra register-allocation granule (4 warps) and cap the DMMA warps at 168 registers.
template <bool CPLX>
struct LoadCursor {
  int tile, k0;
  const GemmProblemDev* p;
  int m0, n0, K;
  __device__ __forceinline__ void open(const GemmProblemDev* probs, const GemmTile* tiles, int ntiles) {
    if (tile < ntiles) {
      const GemmTile t = tiles[tile];
      p = &probs[t.prob]; m0 = t.m0; n0 = t.n0; K = p->K; k0 = 0;
      asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(&p->tmA) : "memory");
      asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(&p->tmB) : "memory");
    }
  }
};

template <bool CPLX, int WARPS_N, int WN_ = 32, bool M3 = false>
__global__ void __launch_bounds__(GemmCfg<CPLX, WARPS_N, WN_, M3>::NTHREADS, GemmCfg<CPLX, WARPS_N, WN_, M3>::CTAS_PER_SM)
gemm_dmma_kernel(const GemmProblemDev* __restrict__ probs, const GemmTile* __restrict__ tiles, int ntiles) {
  using Cfg = GemmCfg<CPLX, WARPS_N, WN_, M3>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, KSUB = Cfg::KSUB, NSUB = Cfg::NSUB, STAGES = Cfg::STAGES;
  constexpr int KSTAGE = KSUB * NSUB;                       // elements of K per stage
  constexpr int A_SUB_BYTES = BM * 64, B_SUB_BYTES = BN * 64;
  constexpr int STAGE_BYTES = NSUB * (A_SUB_BYTES + B_SUB_BYTES);
  constexpr int WM = CPLX ? 32 : 64, WN = Cfg::WN;          // warp tile
  constexpr int MI = WM / 8, NI = WN / 8;
  constexpr int WARPS_M = Cfg::WARPS_M;                     // 4 (cplx) / 2 (real)
  constexpr int NWARPS = Cfg::WARPS_M * Cfg::WARPS_N;
  constexpr int DPE = CPLX ? 2 : 1;                         // doubles per element

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const bool producer = (threadIdx.x == 0);
  LoadCursor<CPLX> cur;
  int pstage = 0; uint32_t pphase = 0;
  cur.tile = blockIdx.x; cur.k0 = 0; cur.p = nullptr; cur.m0 = cur.n0 = cur.K = 0;

  auto issue_one = [&]() {
    // producer only: load the next k-block of the flat sequence (skipping empty-K tiles) if any is left
    while (cur.tile < ntiles && cur.k0 >= cur.K) { cur.tile += gridDim.x; cur.open(probs, tiles, ntiles); }
    if (cur.tile >= ntiles) return;
    mbar_wait(&empty_bar[pstage], pphase ^ 1);
    mbar_expect_tx(&full_bar[pstage], STAGE_BYTES);
    unsigned char* sa = smem + pstage * STAGE_BYTES;
    unsigned char* sb = sa + NSUB * A_SUB_BYTES;
#pragma unroll
    for (int s = 0; s < NSUB; ++s) {
      tma_load_2d(sa + s * A_SUB_BYTES, &cur.p->tmA, &full_bar[pstage], (cur.k0 + s * KSUB) * DPE, cur.m0);
      tma_load_2d(sb + s * B_SUB_BYTES, &cur.p->tmB, &full_bar[pstage], (cur.k0 + s * KSUB) * DPE, cur.n0);
    }
    cur.k0 += KSTAGE;
    if (++pstage == STAGES) { pstage = 0; pphase ^= 1; }
  };

  if (producer) {
    cur.open(probs, tiles, ntiles);
#pragma unroll 1
    for (int s = 0; s < STAGES - 1; ++s) issue_one();
  }

  const int g = lane >> 2, t4 = lane & 3;
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  int stage = 0; uint32_t phase = 0;
  const uint32_t frag_a = smem_base + (wm * WM + g) * 64 + t4 * 16;
  const uint32_t frag_b = smem_base + NSUB * A_SUB_BYTES + (wn * WN + g) * 64 + t4 * 16;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const GemmTile t = tiles[tile];
    const GemmProblemDev* p = &probs[t.prob];
    const int K = p->K, M = p->M, N = p->N;

    double acc_re[MI][NI][2];
    double acc_im[CPLX ? MI : 1][CPLX ? NI : 1][2];
    double acc_p3[M3 ? MI : 1][M3 ? NI : 1][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        acc_re[i][j][0] = 0.0; acc_re[i][j][1] = 0.0;
        if (CPLX) { acc_im[i][j][0] = 0.0; acc_im[i][j][1] = 0.0; }
        if (M3) { acc_p3[i][j][0] = 0.0; acc_p3[i][j][1] = 0.0; }
      }

#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += KSTAGE) {
      if (producer) issue_one();
      __syncwarp();
      mbar_wait(&full_bar[stage], phase);
      const uint32_t sa = frag_a + stage * STAGE_BYTES, sb = frag_b + stage * STAGE_BYTES;
      // wide 3M: 192 accumulator registers leave no room to hold two k-steps of fragments (the second warp of the
      // scheduler covers the LDS latency); every other configuration double-buffers through the unrolled loop
#pragma unroll ((M3 && WN_ == 32) ? 1 : 2)
      for (int s = 0; s < NSUB; ++s) {
        double2 a[MI], b[NI];
#pragma unroll
        for (int i = 0; i < MI; ++i) a[i] = lds128(sa + s * A_SUB_BYTES + i * 8 * 64);
#pragma unroll
        for (int j = 0; j < NI; ++j) b[j] = lds128(sb + s * B_SUB_BYTES + j * 8 * 64);
        if (M3) {
          // 3M complex product (three real DMMAs instead of four, 25 % less tensor-pipe work):
          //   P1 += ar br ; P2 += ai bi ; P3 += (ar + ai)(br + bi)   ->   re = P1 - P2 ; im = P3 - P1 - P2  (epilogue)
          // The 8 operand sums per k-step share the FP64 pipe with the 48 DMMAs (tools/dmma_dadd_mix.cu: 5 % of the pipe);
          // the compiler overwrites the real parts with them once P1 has consumed those, so they cost no registers.
          double as[MI], bs[NI];
#pragma unroll
          for (int i = 0; i < MI; ++i) as[i] = a[i].x + a[i].y;
#pragma unroll
          for (int j = 0; j < NI; ++j) bs[j] = b[j].x + b[j].y;
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc_re[i][j][0], acc_re[i][j][1], a[i].x, b[j].x);
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc_im[i][j][0], acc_im[i][j][1], a[i].y, b[j].y);
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc_p3[i][j][0], acc_p3[i][j][1], as[i], bs[j]);
        } else if (CPLX) {
          // (ar + i ai)(br + i bi): re += ar br - ai bi ; im += ar bi + ai br
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc_re[i][j][0], acc_re[i][j][1], a[i].x, b[j].x);
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc_im[i][j][0], acc_im[i][j][1], a[i].x, b[j].y);
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc_im[i][j][0], acc_im[i][j][1], a[i].y, b[j].x);
#pragma unroll
          for (int i = 0; i < MI; ++i) {
            const double nai = -a[i].y;
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc_re[i][j][0], acc_re[i][j][1], nai, b[j].y);
          }
        } else {
          // lane holds k = 2t (x) and k = 2t+1 (y) of this 8-wide sub-tile: two k4 groups {0,2,4,6},{1,3,5,7}
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc_re[i][j][0], acc_re[i][j][1], a[i].x, b[j].x);
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc_re[i][j][0], acc_re[i][j][1], a[i].y, b[j].y);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }

    // ------------------------------------------------------------------ epilogue (registers -> global)
    double* C = (double*)p->C;
    const int64_t ldc = p->ldc;
    const bool accum = p->accumulate != 0;
    if (M3) {
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double p1 = acc_re[i][j][e], p2 = acc_im[i][j][e];
            acc_re[i][j][e] = p1 - p2;
            acc_im[i][j][e] = acc_p3[i][j][e] - (p1 + p2);
          }
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int row = t.m0 + wm * WM + i * 8 + g;
      if (row >= M) continue;
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int col = t.n0 + wn * WN + j * 8 + 2 * t4;
        if (col >= N) continue;
        if (CPLX) {
          double* dst = C + ((int64_t)row * ldc + col) * 2;
          double2 v0 = make_double2(acc_re[i][j][0], acc_im[i][j][0]);
          double2 v1 = make_double2(acc_re[i][j][1], acc_im[i][j][1]);
          if (accum) { double2 o = *(double2*)dst; v0.x += o.x; v0.y += o.y; }
          *(double2*)dst = v0;
          if (col + 1 < N) {
            if (accum) { double2 o = *(double2*)(dst + 2); v1.x += o.x; v1.y += o.y; }
            *(double2*)(dst + 2) = v1;
          }
        } else {
          double* dst = C + (int64_t)row * ldc + col;
          if (col + 1 < N) {
            double2 v = make_double2(acc_re[i][j][0], acc_re[i][j][1]);
            if (accum) { double2 o = *(double2*)dst; v.x += o.x; v.y += o.y; }
            *(double2*)dst = v;
          } else {
            double v = acc_re[i][j][0];
            if (accum) v += *dst;
            *dst = v;
          }
        }
      }
    }
  }
}

// Debug / cross-check kernel: one thread per output element (selected with MPSK_FLAG_DEBUG_GEMM).
template <bool CPLX>
__global__ void gemm_debug_kernel(const GemmProblemDev* __restrict__ probs, int nprob) {
  for (int pi = blockIdx.y; pi < nprob; pi += gridDim.y) {
    const GemmProblemDev p = probs[pi];
    const int64_t total = (int64_t)p.M * p.N;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
      const int m = (int)(idx / p.N), n = (int)(idx % p.N);
      if (CPLX) {
        const double2* a = (const double2*)p.A + m * p.lda;
        const double2* b = (const double2*)p.B + n * p.ldb;
        double re = 0, im = 0;
        for (int k = 0; k < p.K; ++k) { re += a[k].x * b[k].x - a[k].y * b[k].y; im += a[k].x * b[k].y + a[k].y * b[k].x; }
        double2* c = (double2*)p.C + m * p.ldc + n;
        if (p.accumulate) { re += c->x; im += c->y; }
        *c = make_double2(re, im);
      } else {
        const double* a = (const double*)p.A + m * p.lda;
        const double* b = (const double*)p.B + n * p.ldb;
        double s = 0;
        for (int k = 0; k < p.K; ++k) s += a[k] * b[k];
        double* c = (double*)p.C + m * p.ldc + n;
        if (p.accumulate) s += *c;
        *c = s;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return nullptr;
    fn = (PFN_encodeTiled)f;
  }
  return fn;
}

static bool encode_map(CUtensorMap* tm, const void* base, int64_t rows, int64_t K, int64_t ld, bool cplx, int boxrows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  const int dpe = cplx ? 2 : 1;
  cuuint64_t gdim[2] = {(cuuint64_t)(K * dpe), (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)(ld * dpe * 8)};
  cuuint32_t box[2] = {8u, (cuuint32_t)boxrows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <bool CPLX, int WARPS_N, int WN, bool M3> static int launch_cfg(Ctx* ctx, const GemmPlan* plan, int slot) {
  using Cfg = GemmCfg<CPLX, WARPS_N, WN, M3>;
  constexpr int smem = Cfg::STAGES * Cfg::NSUB * (Cfg::BM + Cfg::BN) * 64 + 2 * Cfg::STAGES * 8 + 128;
  auto kern = gemm_dmma_kernel<CPLX, WARPS_N, WN, M3>;
  const int grid = std::min(plan->ntiles, ctx->num_sms * Cfg::CTAS_PER_SM);
  if (!ctx->gemm_attr_set[slot]) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ctx->gemm_attr_set[slot] = true;
  }
  kern<<<grid, Cfg::NTHREADS, smem, ctx->stream>>>(plan->d_probs, plan->d_tiles, plan->ntiles);
  return MPSK_OK;
}

GemmPlan::~GemmPlan() {
  if (d_probs) ctx->free(d_probs);
  if (d_tiles) ctx->free(d_tiles);
  if (partials) ctx->free(partials);
  delete reduce;
}

// Build the device problem table (with TMA descriptors) and the tile list.
// Split-K: when the whole plan has fewer tiles than ~4 waves of the 148 persistent CTAs (small bond dimensions, or a
// row panel of a sharded matvec: e.g. 175 tiles of K ~ 18000 at 8 GPUs), long-K problems are cut into K-slices.
// Slice 0 writes C as before; slices s >= 1 write partial products to a scratch buffer that one streaming pass
// (the mix kernel, C += sum_s P_s) folds back.  Deterministic: no atomics, fixed summation order.
int gemm_plan_build(Ctx* ctx, bool cplx, const std::vector<GemmProblemHost>& probs_in, GemmPlan* plan) {
  plan->ctx = ctx; plan->cplx = cplx; plan->nprob = 0; plan->ntiles = 0; plan->flops = 0;
  const int BM = 128;
  const int KSTAGE = cplx ? 16 : 32;
  const size_t esz = cplx ? 16 : 8;
  // ---- tile configuration: narrow (128 x 32, complex only) when it cuts the quantised tile work by > 7 %
  int BN = cplx ? 32 * GEMM_WIDE_N_CPLX : 32 * GEMM_WIDE_N_REAL;
  plan->narrow = false; plan->narrow16 = false;
  plan->m3 = cplx && !(ctx->flags & MPSK_FLAG_GEMM_4M);
  if (cplx && !(ctx->flags & MPSK_FLAG_DEBUG_GEMM)) {
    double wide = 0, nar = 0, nar16 = 0;
    for (const GemmProblemHost& h : probs_in) {
      if (h.M <= 0 || h.N <= 0) continue;
      const double mt = (double)((h.M + BM - 1) / BM) * BM * h.K;
      wide += mt * ((h.N + 63) / 64) * 64; nar += mt * ((h.N + 31) / 32) * 32; nar16 += mt * ((h.N + 15) / 16) * 16;
    }
    if (nar < 0.93 * wide) { plan->narrow = true; BN = 32 * GEMM_NARROW_N_CPLX; }
    // 16-wide tiles cost ~1.5x the shared-memory operand traffic per DMMA: only when they remove > 6 % of the padded work
    if (plan->narrow && nar16 < 0.94 * nar && !getenv("MPSK_NO_NARROW16")) { plan->narrow16 = true; BN = 16; }
  }
  // ---- split-K decision
  int64_t base_tiles = 0; int64_t maxK = 0;
  for (const GemmProblemHost& h : probs_in) {
    if (h.M <= 0 || h.N <= 0) continue;
    base_tiles += (int64_t)((h.M + BM - 1) / BM) * ((h.N + BN - 1) / BN);
    maxK = std::max<int64_t>(maxK, h.K);
    plan->flops += (cplx ? 8.0 : 2.0) * (double)h.M * h.N * h.K;
  }
  const int64_t want_tiles = 4LL * ctx->num_sms * (plan->narrow16 ? 3 : (plan->narrow ? 2 : 1));
  int S = 1;
  if (base_tiles > 0 && base_tiles < want_tiles && maxK >= 512 && !(ctx->flags & MPSK_FLAG_DEBUG_GEMM))
    S = (int)std::min<int64_t>(8, (want_tiles + base_tiles - 1) / base_tiles);
  std::vector<GemmProblemHost> probs;
  std::vector<MixTask> rtasks; std::vector<MixTerm> rterms;
  if (S > 1) {
    // scratch for the partial products of slices >= 1
    int64_t pel = 0;
    std::vector<int> ns(probs_in.size(), 1);
    for (size_t i = 0; i < probs_in.size(); ++i) {
      const GemmProblemHost& h = probs_in[i];
      if (h.M <= 0 || h.N <= 0) continue;
      int si = (int)std::min<int64_t>(S, h.K / 256);          // keep every slice >= 256 deep
      ns[i] = std::max(si, 1);
      pel += (int64_t)(ns[i] - 1) * padg((int64_t)h.M * pad(h.N));
    }
    if (pel > 0) {
      plan->partials = ctx->alloc((size_t)pel * esz);
      if (!plan->partials) return ctx->fail(MPSK_ENOMEM, "split-K scratch");
    }
    int64_t poff = 0;
    for (size_t i = 0; i < probs_in.size(); ++i) {
      const GemmProblemHost& h = probs_in[i];
      if (h.M <= 0 || h.N <= 0) continue;
      if (ns[i] <= 1) { probs.push_back(h); continue; }
      const int64_t kchunk = ((h.K + ns[i] - 1) / ns[i] + KSTAGE - 1) / KSTAGE * KSTAGE;
      MixTask tk; tk.dst = h.C; tk.ldd = h.ldc; tk.rows = h.M; tk.cols = h.N; tk.term_begin = (int32_t)rterms.size();
      rterms.push_back(MixTerm{h.C, h.ldc, 1.0, 0.0});
      int s = 0;
      for (int64_t k0 = 0; k0 < h.K; k0 += kchunk, ++s) {
        GemmProblemHost q = h;
        q.A = (const char*)h.A + k0 * esz; q.B = (const char*)h.B + k0 * esz; q.K = (int32_t)std::min<int64_t>(kchunk, h.K - k0);
        if (s > 0) {
          q.C = (char*)plan->partials + poff * esz; q.ldc = pad(h.N); q.accumulate = 0;
          rterms.push_back(MixTerm{q.C, q.ldc, 1.0, 0.0});
          poff += padg((int64_t)h.M * pad(h.N));
        }
        probs.push_back(q);
      }
      tk.term_end = (int32_t)rterms.size();
      if (tk.term_end - tk.term_begin > 1) rtasks.push_back(tk);
    }
  } else {
    for (const GemmProblemHost& h : probs_in) if (h.M > 0 && h.N > 0) probs.push_back(h);
  }
  std::vector<GemmProblemDev> dev;
  std::vector<GemmTile> tiles;
  // order problems by decreasing K so that the long tiles are scheduled first (tail balance)
  std::vector<int> order;
  for (size_t i = 0; i < probs.size(); ++i) order.push_back((int)i);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return probs[a].K > probs[b].K; });
  for (int oi : order) {
    const GemmProblemHost& h = probs[oi];
    GemmProblemDev d;
    memset(&d, 0, sizeof(d));
    d.A = h.A; d.B = h.B; d.C = h.C; d.lda = h.lda; d.ldb = h.ldb; d.ldc = h.ldc;
    d.M = h.M; d.N = h.N; d.K = h.K; d.accumulate = h.accumulate;
    if (h.K > 0) {
      if (((uintptr_t)h.A % 16) || ((uintptr_t)h.B % 16) || ((h.lda * esz) % 16) || ((h.ldb * esz) % 16) ||
          ((uintptr_t)h.C % 16) || ((h.ldc * esz) % 16))
        return ctx->fail(MPSK_EINTERNAL, "gemm operand not 16-byte aligned");
      if (!(ctx->flags & MPSK_FLAG_DEBUG_GEMM)) {
        if (!encode_map(&d.tmA, h.A, h.M, h.K, h.lda, cplx, BM) || !encode_map(&d.tmB, h.B, h.N, h.K, h.ldb, cplx, BN))
          return ctx->fail(MPSK_ECUDA, "cuTensorMapEncodeTiled failed");
      }
    }
    const int pi = (int)dev.size();
    dev.push_back(d);
    // L2-aware raster: m-tiles are grouped into super-rows whose A panel (G x BM rows x K) stays L2-resident (~24 MB);
    // inside a super-row the n-panels are the outer loop, so the ~148 concurrently running CTAs share one or a few B
    // panels while the A super-row is re-used from L2 by every following n-panel instead of being streamed from DRAM
    // once per n-panel (round 1: 42.9 GB of DRAM reads for 1.2 GB of stage-A operands).
    const int64_t mt = (h.M + BM - 1) / BM;
    int64_t G = ((int64_t)24 << 20) / std::max<int64_t>((int64_t)BM * h.K * (int64_t)esz, 1);
    G = std::max<int64_t>(1, std::min<int64_t>(G, mt));
    for (int64_t ms = 0; ms < mt; ms += G)
      for (int n0 = 0; n0 < h.N; n0 += BN)
        for (int64_t mi = ms; mi < std::min(ms + G, mt); ++mi) tiles.push_back(GemmTile{pi, (int32_t)(mi * BM), n0, 0});
  }
  plan->nprob = (int)dev.size();
  plan->ntiles = (int)tiles.size();
  plan->splitk = S;
  if (plan->nprob == 0) return MPSK_OK;
  { void* d = nullptr; MPSK_TRY(upload_table(ctx, dev.data(), dev.size() * sizeof(GemmProblemDev), &d)); plan->d_probs = (GemmProblemDev*)d; }
  { void* d = nullptr; MPSK_TRY(upload_table(ctx, tiles.data(), tiles.size() * sizeof(GemmTile), &d)); plan->d_tiles = (GemmTile*)d; }
  if (!rtasks.empty()) {
    plan->reduce = new MixPlan;
    MPSK_TRY(mix_plan_build(ctx, cplx, rtasks, rterms, plan->reduce));
  }
  return MPSK_OK;
}

int gemm_plan_run(Ctx* ctx, const GemmPlan* plan) {
  if (plan->nprob == 0 || plan->ntiles == 0) return MPSK_OK;
  if (ctx->flags & MPSK_FLAG_DEBUG_GEMM) {
    dim3 grid(64, std::min(plan->nprob, 1024));
    if (plan->cplx) gemm_debug_kernel<true><<<grid, 256, 0, ctx->stream>>>(plan->d_probs, plan->nprob);
    else gemm_debug_kernel<false><<<grid, 256, 0, ctx->stream>>>(plan->d_probs, plan->nprob);
  } else {
    if (!plan->cplx) MPSK_TRY((launch_cfg<false, GEMM_WIDE_N_REAL, 32, false>(ctx, plan, TC_REAL)));
    else if (plan->m3) {
      if (plan->narrow16) MPSK_TRY((launch_cfg<true, 1, 16, true>(ctx, plan, TC_3M + TC_NARROW16)));
      else if (plan->narrow) MPSK_TRY((launch_cfg<true, GEMM_NARROW_N_CPLX, 32, true>(ctx, plan, TC_3M + TC_NARROW)));
      else MPSK_TRY((launch_cfg<true, GEMM_WIDE_N_CPLX, 32, true>(ctx, plan, TC_3M + TC_WIDE)));
    } else {
      if (plan->narrow16) MPSK_TRY((launch_cfg<true, 1, 16, false>(ctx, plan, TC_NARROW16)));
      else if (plan->narrow) MPSK_TRY((launch_cfg<true, GEMM_NARROW_N_CPLX, 32, false>(ctx, plan, TC_NARROW)));
      else MPSK_TRY((launch_cfg<true, GEMM_WIDE_N_CPLX, 32, false>(ctx, plan, TC_WIDE)));
    }
  }
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  if (plan->reduce) MPSK_TRY(mix_plan_run(ctx, plan->reduce));
  return MPSK_OK;
}

}  // namespace mpsk
