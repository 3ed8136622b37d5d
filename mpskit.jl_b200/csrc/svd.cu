// Post-eigensolve truncated SVD of the two-site tensor, on the device (dmrg.jl:96-103,112-119):
//   tsvd!(ac2; trunc, alg=SVD()) -> per coupled sector one gesvd (cuSOLVER; LAPACK-class library call, like the
//   reference's LAPACK gesvd), global truncation select over all sectors on the host (integer bookkeeping,
//   bit-exact with the oracle), normalize!(c), and al / c / ar written straight into GEMM-ready layouts.
#include "internal.h"
#include <cusolverDn.h>
#include <cmath>
#include <chrono>
#include <cstdlib>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace mpsk {

#define CUSOLVER_TRY(ctx, expr)                                                                  \
  do {                                                                                           \
    cusolverStatus_t _s = (expr);                                                                \
    if (_s != CUSOLVER_STATUS_SUCCESS) return (ctx)->fail(MPSK_ECUDA, "%s failed: cusolver status %d", #expr, (int)_s); \
  } while (0)

static int get_cusolver(Ctx* ctx, cusolverDnHandle_t* h) {
  if (!ctx->cusolver) {
    cusolverDnHandle_t hh;
    CUSOLVER_TRY(ctx, cusolverDnCreate(&hh));
    ctx->cusolver = hh;
  }
  *h = (cusolverDnHandle_t)ctx->cusolver;
  CUSOLVER_TRY(ctx, cusolverDnSetStream(*h, ctx->stream));
  return MPSK_OK;
}

__global__ void diag_kernel(double* dst, int64_t ld, int n, const double* s, double scale, int cplx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (cplx) { dst[2 * (i * ld + i)] = s[i] * scale; dst[2 * (i * ld + i) + 1] = 0.0; }
  else dst[i * ld + i] = s[i] * scale;
}

// ph[i] = R_ii / |R_ii| (1 if zero); Rout(i, j) = conj(ph_i) * A(i, j) for j >= i, 0 below; A column-major (lda)
// Rout element (i, j) is written at dst[i * si + j * sj], conjugated if conj_out (LQ: L = R^H)
template <bool CPLX>
__global__ void qr_phase_kernel(const double* __restrict__ A, int64_t lda, int k, int n, double* __restrict__ ph,
                                double* __restrict__ dst, int64_t si, int64_t sj, int conj_out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)k * n; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % k), j = (int)(e / k);
    double pr = 1.0, pi = 0.0;
    if (CPLX) {
      const double dr = A[2 * (i + (int64_t)i * lda)], di = A[2 * (i + (int64_t)i * lda) + 1];
      const double ab = sqrt(dr * dr + di * di);
      if (ab > 0) { pr = dr / ab; pi = di / ab; }
    } else {
      const double dr = A[i + (int64_t)i * lda];
      if (dr < 0) pr = -1.0;
    }
    if (j == 0) { if (CPLX) { ph[2 * i] = pr; ph[2 * i + 1] = pi; } else ph[i] = pr; }
    double vr = 0.0, vi = 0.0;
    if (j >= i) {
      if (CPLX) {
        const double ar = A[2 * (i + (int64_t)j * lda)], ai = A[2 * (i + (int64_t)j * lda) + 1];
        vr = pr * ar + pi * ai; vi = pr * ai - pi * ar;       // conj(ph) * a
      } else vr = pr * A[i + (int64_t)j * lda];
    }
    if (CPLX) { dst[2 * (i * si + j * sj)] = vr; dst[2 * (i * si + j * sj) + 1] = conj_out ? -vi : vi; }
    else dst[i * si + j * sj] = vr;
  }
}

// Q(:, i) *= ph[i]; Q column-major m x k
template <bool CPLX>
__global__ void q_phase_kernel(double* __restrict__ Q, int64_t ldq, int m, int k, const double* __restrict__ ph) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)m * k; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e % m), i = (int)(e / m);
    if (CPLX) {
      const double pr = ph[2 * i], pi = ph[2 * i + 1];
      const double qr = Q[2 * (r + (int64_t)i * ldq)], qi = Q[2 * (r + (int64_t)i * ldq) + 1];
      Q[2 * (r + (int64_t)i * ldq)] = qr * pr - qi * pi; Q[2 * (r + (int64_t)i * ldq) + 1] = qr * pi + qi * pr;
    } else Q[r + (int64_t)i * ldq] *= ph[i];
  }
}

// Gram route: emit the `head` singular triplets of one level.  lam ascending eigenvalues of M^H M (nl of them), the head
// is the index range [t, nl); Bn = M V and Rn = R V column-major.  Output position of eigen-index i: outpos + (nl-1-i)
// (descending).  The singular value is the NORM OF THE COLUMN b_i = M v_i (gram_colnorm_kernel), i.e. the Rayleigh
// quotient of the exact Gram matrix: its error is second order in the eigenvector error, whereas sqrt(lam_i) carries the
// first-order rounding error eps * lam_top of the explicitly formed Gram matrix.  U[:, o] = b_i / |b_i|,
// VT[o, :] = conj(Rn[:, i])^T, so U S V^H reproduces M to rounding before the final re-orthonormalisation.
template <bool CPLX>
__global__ void __launch_bounds__(256) gram_colnorm_kernel(const double* __restrict__ Bn, int64_t ldb, int mm, int t, int nl, int outpos,
                                                            double* __restrict__ S) {
  const int jj = blockIdx.x, i = nl - 1 - jj;
  if (i < t) return;
  const double* col = Bn + (CPLX ? 2 : 1) * (int64_t)i * ldb;
  double acc = 0.0;
  for (int r = threadIdx.x; r < (CPLX ? 2 : 1) * mm; r += blockDim.x) { const double v = col[r]; acc += v * v; }
  __shared__ double sm[256];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) S[outpos + jj] = sqrt(sm[0]);
}

template <bool CPLX>
__global__ void gram_emit_kernel(const double* __restrict__ Bn, int64_t ldb, const double* __restrict__ Rn, int64_t ldr,
                                 int mm, int nn, int t, int nl, int outpos,
                                 double* __restrict__ U, int64_t ldu, double* __restrict__ VT, int64_t ldvt, const double* __restrict__ S) {
  const int hc = nl - t;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = tid; e < (int64_t)mm * hc; e += nth) {
    const int r = (int)(e % mm), jj = (int)(e / mm), i = nl - 1 - jj, o = outpos + jj;
    const double sg = S[o], inv = sg > 0 ? 1.0 / sg : 0.0;
    if (CPLX) { U[2 * (r + (int64_t)o * ldu)] = Bn[2 * (r + (int64_t)i * ldb)] * inv; U[2 * (r + (int64_t)o * ldu) + 1] = Bn[2 * (r + (int64_t)i * ldb) + 1] * inv; }
    else U[r + (int64_t)o * ldu] = Bn[r + (int64_t)i * ldb] * inv;
  }
  for (int64_t e = tid; e < (int64_t)nn * hc; e += nth) {
    const int c = (int)(e % nn), jj = (int)(e / nn), i = nl - 1 - jj, o = outpos + jj;
    if (CPLX) { VT[2 * (o + (int64_t)c * ldvt)] = Rn[2 * (c + (int64_t)i * ldr)]; VT[2 * (o + (int64_t)c * ldvt) + 1] = -Rn[2 * (c + (int64_t)i * ldr) + 1]; }
    else VT[o + (int64_t)c * ldvt] = Rn[c + (int64_t)i * ldr];
  }
}

// ---- persistent pool of host threads driving independent cuSOLVER factorizations (one stream + handle per worker)
struct WorkerPool {
  std::vector<std::thread> th; std::mutex m; std::condition_variable cv_job, cv_done;
  std::function<void(int)> job; int gen = 0, pending = 0, nactive = 0; bool stop = false;
  void start(int n, int dev) {
    for (int t = (int)th.size(); t < n; ++t)
      th.emplace_back([this, t, dev]() {
        cudaSetDevice(dev);
        int seen = 0;
        for (;;) {
          std::function<void(int)> f; bool mine = false;
          {
            std::unique_lock<std::mutex> lk(m);
            cv_job.wait(lk, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen; mine = t < nactive; f = job;
          }
          if (mine) {
            f(t);
            std::unique_lock<std::mutex> lk(m);
            if (--pending == 0) cv_done.notify_all();
          }
        }
      });
  }
  void run(int nthreads, std::function<void(int)> f) {
    if (nthreads <= 0) return;
    std::unique_lock<std::mutex> lk(m);
    job = std::move(f); nactive = nthreads; pending = nthreads; ++gen;
    cv_job.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
  ~WorkerPool() {
    { std::unique_lock<std::mutex> lk(m); stop = true; }
    cv_job.notify_all();
    for (auto& t : th) t.join();
  }
};
constexpr int POOL_THREADS = 4;

void pool_destroy(Ctx* ctx) {
  delete (WorkerPool*)ctx->pool; ctx->pool = nullptr;
  for (int t = 0; t < POOL_THREADS; ++t) {
    if (ctx->pool_solver[t]) cusolverDnDestroy((cusolverDnHandle_t)ctx->pool_solver[t]);
    if (ctx->pool_stream[t]) cudaStreamDestroy((cudaStream_t)ctx->pool_stream[t]);
    if (ctx->pool_info[t]) cudaFree(ctx->pool_info[t]);
    ctx->pool_solver[t] = ctx->pool_stream[t] = nullptr; ctx->pool_info[t] = nullptr;
  }
  if (ctx->cusolver) { cusolverDnDestroy((cusolverDnHandle_t)ctx->cusolver); ctx->cusolver = nullptr; }
}

static int pool_get(Ctx* ctx, WorkerPool** out) {
  for (int t = 0; t < POOL_THREADS; ++t) {     // persistent per-worker stream + cuSOLVER handle (creation costs milliseconds)
    if (ctx->pool_solver[t]) continue;
    cudaStream_t st; cusolverDnHandle_t hs; int* di = nullptr;
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CUSOLVER_TRY(ctx, cusolverDnCreate(&hs));
    CUSOLVER_TRY(ctx, cusolverDnSetStream(hs, st));
    CUDA_TRY(ctx, cudaMalloc(&di, sizeof(int)));
    ctx->pool_stream[t] = st; ctx->pool_solver[t] = hs; ctx->pool_info[t] = di;
  }
  if (!ctx->pool) { WorkerPool* p = new WorkerPool; p->start(POOL_THREADS, ctx->device); ctx->pool = p; }
  *out = (WorkerPool*)ctx->pool;
  return MPSK_OK;
}

struct Coupled {
  int32_t Q;
  std::vector<std::pair<int, int>> rows, cols;    // (l-sector, u), (r-sector, t)
  std::vector<int64_t> rowoff, coloff;
  int64_t m = 0, n = 0, k = 0;
  bool flip = false;                              // factorise M^H instead (rows >= cols)
  bool mine = true, fast = false;                 // owned by this rank; Gram route
  int mm = 0, nn = 0;                             // factorised matrix is mm x nn, mm >= nn
  void* W = nullptr; void* U = nullptr; void* VT = nullptr; double* S = nullptr;
  int64_t ldw = 0, ldu = 0, ldvt = 0;
  int keep = 0;
};

// per-sector state of the multi-level Gram route
struct GramState {
  void* M = nullptr; void* R = nullptr; int nl = 0, outpos = 0, split = 0;     // current level: M (mm x nl), R (nn x nl, null = identity)
  int64_t ldm = 0, ldn = 0;
  void *Mconj = nullptr, *Mrow = nullptr, *Rrow = nullptr, *G = nullptr, *buf[2] = {nullptr, nullptr}, *rbuf[2] = {nullptr, nullptr};
  double* lam = nullptr; void* work = nullptr; int lwork = 0;
  void* tau = nullptr; double* ph = nullptr; void* Rtmp = nullptr; void* work2 = nullptr; int lwork2 = 0;
  std::vector<double> hlam; double lam0 = 0.0;
};

// Truncated SVD of the two-site tensor (dmrg.jl:96-97).  Per coupled sector:
//   * small sectors (min dim < 256): cuSOLVER gesvd (LAPACK-class, what the reference calls);
//   * large sectors: a multi-level ("deflated") Gram route built from our grouped DMMA GEMM + cuSOLVER heevd:
//       level l:  G = M^H M,  G = V L V^H,  B = M V,  R <- R V;   singular values within 1e-3 of the level's largest are
//       emitted (sigma = |b| = norm of the column of B, u = b / sigma, v = r): relative error <= eps * (sigma_top/sigma)^2 <= 1e-10.  The
//       remaining columns of B (norms <= 1e-3 of the level scale, orthogonal to the emitted ones to rounding) form the
//       matrix of the next level, so every level restores three decades of relative accuracy; at most 6 levels reach the
//       rounding floor eps * sigma_max, where gesvd's own absolute accuracy ends.  A final Householder QR with positive
//       diagonal makes U an exact isometry.  (Round 1 used a single level: singular values below 1e-8 sigma_max were
//       noise, see ADVICE.md.)
// Multi-GPU: the coupled sectors are distributed over the ranks (longest-processing-time rule on m n^2); the singular
// values are combined with one allreduce so that every rank takes the same global truncation decision, and the output
// tensors are assembled with one allreduce each (every element is written by exactly one rank: bit-identical results).
int tsvd(Ctx* ctx, const Tensor* x, int d1, const int32_t* p1ch, mpsk_trunc trunc, Tensor** Uo, Tensor** So, Tensor** Vo,
         double* eps_out) {
  const bool normalize = !(trunc.kind & 0x100);
  trunc.kind &= 0xff;
  if (d1 <= 0 || x->nm() % d1) return ctx->fail(MPSK_EINVAL, "tsvd: d1 does not divide the label count");
  const int d2 = x->nm() / d1;
  const bool cx = x->dtype == MPSK_C128;
  const size_t es = cx ? 16 : 8;
  std::vector<int32_t> s1(p1ch, p1ch + d1), s2(d2);
  for (int t = 0; t < d2; ++t) s2[t] = x->mch[t] - s1[0];
  for (int u = 0; u < d1; ++u) for (int t = 0; t < d2; ++t)
    if (x->mch[u * d2 + t] != s1[u] + s2[t]) return ctx->fail(MPSK_ESPACE, "tsvd: label charges are not of the form s1[p1] + s2[p2]");

  const Tensor* xk = x; bool own = false;
  MPSK_TRY(tensor_as_form(ctx, x, F_KL, false, &xk, &own));
  std::unique_ptr<const Tensor> xown(own ? xk : nullptr);
  const Space& Vl = x->Vl; const Space& Vr = x->Vr;
  // ---- coupled sectors (ascending charge)
  std::map<int64_t, Coupled> cs;
  for (int jl = 0; jl < Vl.nsec(); ++jl) for (int u = 0; u < d1; ++u) {
    const int64_t Q = Vl.norm((int64_t)Vl.q[jl] + s1[u]);
    Coupled& c = cs[Q]; c.Q = (int32_t)Q; c.rows.push_back({jl, u}); c.rowoff.push_back(c.m); c.m += Vl.n[jl];
  }
  for (int jr = 0; jr < Vr.nsec(); ++jr) for (int t = 0; t < d2; ++t) {
    const int64_t Q = Vl.norm((int64_t)Vr.q[jr] - s2[t]);
    auto it = cs.find(Q);
    if (it == cs.end()) continue;
    Coupled& c = it->second; c.cols.push_back({jr, t}); c.coloff.push_back(c.n); c.n += Vr.n[jr];
  }
  std::vector<Coupled*> secs;
  for (auto& kv : cs) if (kv.second.m > 0 && kv.second.n > 0) secs.push_back(&kv.second);
  const int ns = (int)secs.size();
  // every scratch buffer of this call comes out of ONE allocation (bump arena): a centre update at D = 4096 needs ~150
  // buffers whose sizes change from site to site, which defeats the caching allocator (cudaMalloc dominated the first
  // profile of this function)
  struct Arena {
    Ctx* ctx; char* base = nullptr; size_t off = 0; bool leased = false;
    void* take(size_t bytes) { off = (off + 255) & ~(size_t)255; void* p = base ? base + off : (void*)(uintptr_t)(off + 256); off += std::max<size_t>(bytes, 16); return p; }
    ~Arena() { if (base) ctx->unlease(Ctx::S_TSVD, base, leased); }
  } arena{ctx};
  // gesvd only for tiny sectors: cuSOLVER's QR-iteration SVD is latency-bound (tens of ms already at 200 x 200), the
  // multi-level Gram route is as accurate and an order of magnitude faster from ~32 columns on
  const int64_t fast_min = (ctx->flags & MPSK_FLAG_EXACT_SVD) ? (int64_t)1 << 60 : ((ctx->flags & MPSK_FLAG_FAST_SVD_ALL) ? 2 : 32);
  if (ctx->profile < 0) ctx->profile = getenv("MPSK_PROFILE") ? 1 : 0;
  auto tnow = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tp0 = ctx->profile ? tnow() : 0.0;
  double tp_prev = tp0;
  auto lap = [&](const char* what, int a = 0, int b = 0) {
    if (!ctx->profile) return;
    cudaStreamSynchronize(ctx->stream);
    const double t = tnow();
    fprintf(stderr, "[mpsk profile] tsvd %-14s %8.2f ms  (%d, %d)\n", what, (t - tp_prev) * 1e3, a, b);
    tp_prev = t;
  };
  // ---- ownership: sectors are distributed over the ranks (LPT on m n^2), deterministic on every rank
  const bool dist = ctx->nranks > 1 && !(ctx->flags & MPSK_FLAG_REPLICATE_LOCAL);
  std::vector<int> owner(ns, 0);
  {
    std::vector<int> ord(ns);
    for (int i = 0; i < ns; ++i) ord[i] = i;
    auto cost = [&](int i) { const double a = (double)std::max(secs[i]->m, secs[i]->n), b = (double)std::min(secs[i]->m, secs[i]->n); return a * b * b; };
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return cost(a) > cost(b); });
    std::vector<double> load(std::max(ctx->nranks, 1), 0.0);
    for (int i : ord) {
      int best = 0;
      if (dist) for (int r = 1; r < ctx->nranks; ++r) if (load[r] < load[best]) best = r;
      owner[i] = best; load[best] += cost(i);
    }
  }
  // ---- sizes, ownership, scratch layout
  int64_t ktot = 0; std::vector<int64_t> koff(ns);
  for (int si = 0; si < ns; ++si) {
    Coupled* c = secs[si];
    c->k = std::min(c->m, c->n); c->flip = c->m < c->n;
    c->mm = (int)(c->flip ? c->n : c->m); c->nn = (int)(c->flip ? c->m : c->n);
    c->mine = !dist || owner[si] == ctx->rank;
    c->fast = c->k >= fast_min;
    c->ldw = pad(c->mm); c->ldu = pad(c->mm); c->ldvt = pad(c->k);
    koff[si] = ktot; ktot += c->k;
  }
  cusolverDnHandle_t h0;
  MPSK_TRY(get_cusolver(ctx, &h0));
  struct SlowScratch { void* work = nullptr; double* rwork = nullptr; int lwork = 0; };
  std::vector<SlowScratch> ss(ns); std::vector<GramState> gs(ns);
  void* dummy = ctx->alloc(256);           // the size queries only want a valid device pointer
  if (!dummy) return MPSK_ENOMEM;
  struct DummyFree { Ctx* c; void* p; ~DummyFree() { c->free(p); } } dummy_free{ctx, dummy};
  for (int si = 0; si < ns; ++si) {        // cuSOLVER workspace sizes (host-side queries, no device work)
    Coupled* c = secs[si];
    if (!c->mine) continue;
    const int mm = c->mm, nn = c->nn;
    if (!c->fast) {
      if (cx) CUSOLVER_TRY(ctx, cusolverDnZgesvd_bufferSize(h0, mm, nn, &ss[si].lwork));
      else CUSOLVER_TRY(ctx, cusolverDnDgesvd_bufferSize(h0, mm, nn, &ss[si].lwork));
      continue;
    }
    GramState& g = gs[si];
    g.ldm = pad(mm); g.ldn = pad(nn);
    int l1 = 0, l2 = 0, l3 = 0;
    if (cx) { CUSOLVER_TRY(ctx, cusolverDnZheevd_bufferSize(h0, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, nn, (cuDoubleComplex*)dummy, (int)g.ldn, (double*)dummy, &l1));
              CUSOLVER_TRY(ctx, cusolverDnZgeqrf_bufferSize(h0, mm, nn, (cuDoubleComplex*)dummy, (int)c->ldu, &l2));
              CUSOLVER_TRY(ctx, cusolverDnZungqr_bufferSize(h0, mm, nn, nn, (cuDoubleComplex*)dummy, (int)c->ldu, (cuDoubleComplex*)dummy, &l3)); }
    else { CUSOLVER_TRY(ctx, cusolverDnDsyevd_bufferSize(h0, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, nn, (double*)dummy, (int)g.ldn, (double*)dummy, &l1));
           CUSOLVER_TRY(ctx, cusolverDnDgeqrf_bufferSize(h0, mm, nn, (double*)dummy, (int)c->ldu, &l2));
           CUSOLVER_TRY(ctx, cusolverDnDorgqr_bufferSize(h0, mm, nn, nn, (double*)dummy, (int)c->ldu, (double*)dummy, &l3)); }
    g.lwork = std::max(l1, 1); g.lwork2 = std::max(std::max(l2, l3), 1);
  }
  auto layout = [&]() {                    // assigns every buffer from the arena (first call only counts bytes)
    arena.off = 0;
    for (int si = 0; si < ns; ++si) {
      Coupled* c = secs[si];
      if (!c->mine) continue;
      const size_t mm = c->mm, nn = c->nn;
      c->W = arena.take(c->ldw * nn * es); c->U = arena.take(c->ldu * c->k * es); c->VT = arena.take(c->ldvt * nn * es);
      c->S = (double*)arena.take(c->k * sizeof(double));
      if (!c->fast) {
        ss[si].work = arena.take((size_t)std::max(ss[si].lwork, 1) * es);
        ss[si].rwork = (double*)arena.take(sizeof(double) * std::max<size_t>(nn, 1) * 5);
        continue;
      }
      GramState& g = gs[si];
      g.Mconj = arena.take(g.ldm * nn * es); g.Mrow = arena.take(mm * g.ldn * es); g.G = arena.take(g.ldn * nn * es);
      g.buf[0] = arena.take(g.ldm * nn * es); g.buf[1] = arena.take(g.ldm * nn * es);
      g.rbuf[0] = arena.take(g.ldn * nn * es); g.rbuf[1] = arena.take(g.ldn * nn * es); g.Rrow = arena.take(nn * g.ldn * es);
      g.lam = (double*)arena.take(sizeof(double) * nn); g.tau = arena.take(nn * es); g.ph = (double*)arena.take(nn * es);
      g.Rtmp = arena.take(nn * nn * es);
      g.work = arena.take((size_t)g.lwork * es); g.work2 = arena.take((size_t)g.lwork2 * es);
      g.M = c->W; g.R = nullptr; g.nl = (int)nn; g.outpos = 0;
    }
  };
  layout();
  arena.base = (char*)ctx->lease(Ctx::S_TSVD, arena.off + 512, &arena.leased);
  if (!arena.base) return MPSK_ENOMEM;
  layout();
  // ---- assemble the column-major matrices of the sectors this rank factorises
  std::vector<CopyTask> tasks;
  for (int si = 0; si < ns; ++si) {
    Coupled* c = secs[si];
    if (!c->mine) continue;
    const int64_t nn = c->nn;
    CUDA_TRY(ctx, cudaMemsetAsync(c->W, 0, c->ldw * nn * es, ctx->stream));
    for (size_t ri = 0; ri < c->rows.size(); ++ri) for (size_t ci = 0; ci < c->cols.size(); ++ci) {
      const int jl = c->rows[ri].first, u = c->rows[ri].second, jr = c->cols[ci].first, t = c->cols[ci].second;
      const Blk* b = xk->find(u * d2 + t, jl);
      if (!b || b->jr != jr) continue;
      CopyTask tk; tk.src = xk->ptr(b->off); tk.ssl = b->sl; tk.ssr = b->sr; tk.nl = (int32_t)Vl.n[jl]; tk.nr = (int32_t)Vr.n[jr];
      tk.scale_axis = 0; tk.scale = nullptr;
      if (!c->flip) { tk.dst = (char*)c->W + (c->rowoff[ri] + c->coloff[ci] * c->ldw) * es; tk.dsl = 1; tk.dsr = c->ldw; tk.conj = 0; }
      else { tk.dst = (char*)c->W + (c->coloff[ci] + c->rowoff[ri] * c->ldw) * es; tk.dsl = c->ldw; tk.dsr = 1; tk.conj = cx ? 1 : 0; }
      tasks.push_back(tk);
    }
  }
  MPSK_TRY(run_copy_tasks(ctx, tasks, cx));
  WorkerPool* pool = nullptr;
  MPSK_TRY(pool_get(ctx, &pool));
  std::vector<std::vector<double>> sv(ns);
  // work lists (largest first)
  std::vector<int> slow, fastl;
  for (int si = 0; si < ns; ++si) if (secs[si]->mine) (secs[si]->fast ? fastl : slow).push_back(si);
  auto bysize = [&](int a, int b) { return (double)secs[a]->mm * secs[a]->nn > (double)secs[b]->mm * secs[b]->nn; };
  std::sort(slow.begin(), slow.end(), bysize); std::sort(fastl.begin(), fastl.end(), bysize);
  std::vector<int> status(POOL_THREADS, 0);
  std::vector<std::string> errs(POOL_THREADS);
  // generic pool phase: `fn(si, stream, handle, d_info)` enqueues the factorisation of sector si and returns its status
  auto run_phase = [&](const std::vector<int>& list, const char* what,
                       const std::function<cusolverStatus_t(int, cudaStream_t, cusolverDnHandle_t, int*)>& fn,
                       const std::function<void(int, cudaStream_t)>& after) -> int {
    if (list.empty()) return MPSK_OK;
    const int nthreads = (int)std::min<size_t>(POOL_THREADS, list.size());
    std::atomic<int> next(0);
    pool->run(nthreads, [&](int tid) {
      cudaStream_t st = (cudaStream_t)ctx->pool_stream[tid]; cusolverDnHandle_t hs = (cusolverDnHandle_t)ctx->pool_solver[tid];
      int* d_info = ctx->pool_info[tid];
      for (;;) {
        const int oi = next.fetch_add(1);
        if (oi >= (int)list.size()) break;
        const int si = list[oi];
        cusolverStatus_t cs_ = fn(si, st, hs, d_info);
        int info = 0;
        cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (after) after(si, st);
        cudaError_t ce = cudaStreamSynchronize(st);
        if (cs_ != CUSOLVER_STATUS_SUCCESS || ce != cudaSuccess || info != 0) {
          status[tid] = MPSK_ECUDA;
          errs[tid] = std::string(what) + " failed in sector " + std::to_string(secs[si]->Q) + " (status " + std::to_string((int)cs_) + ", info " + std::to_string(info) + ")";
          break;
        }
      }
    });
    for (int t = 0; t < POOL_THREADS; ++t) if (status[t] != 0) return ctx->fail(status[t], "tsvd: %s", errs[t].c_str());
    return MPSK_OK;
  };
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));     // W is assembled: hand the sectors to the pool streams
  lap("assemble", ns, (int)fastl.size());
  // ---- small sectors: gesvd
  MPSK_TRY(run_phase(slow, "gesvd",
    [&](int si, cudaStream_t, cusolverDnHandle_t hs, int* d_info) {
      Coupled* c = secs[si]; SlowScratch& x = ss[si];
      if (cx) return cusolverDnZgesvd(hs, 'S', 'S', c->mm, c->nn, (cuDoubleComplex*)c->W, (int)c->ldw, c->S, (cuDoubleComplex*)c->U, (int)c->ldu,
                                      (cuDoubleComplex*)c->VT, (int)c->ldvt, (cuDoubleComplex*)x.work, x.lwork, x.rwork, d_info);
      return cusolverDnDgesvd(hs, 'S', 'S', c->mm, c->nn, (double*)c->W, (int)c->ldw, c->S, (double*)c->U, (int)c->ldu, (double*)c->VT,
                              (int)c->ldvt, (double*)x.work, x.lwork, x.rwork, d_info);
    },
    [&](int si, cudaStream_t st) { sv[si].resize(secs[si]->k); cudaMemcpyAsync(sv[si].data(), secs[si]->S, sizeof(double) * secs[si]->k, cudaMemcpyDeviceToHost, st); }));
  ctx->launches += (int64_t)slow.size();
  lap("gesvd", (int)slow.size(), slow.empty() ? 0 : secs[slow[0]]->nn);
  // ---- large sectors: multi-level Gram route
  // Two levels (MAXLEV = 1) already bound the ABSOLUTE error of every singular value by ~1e-11 sigma_max: level 1 works on
  // columns of norm <= 1e-3 sigma_max, so even a value lost in its rounding noise is off by at most sqrt(eps) * 1e-3 sigma_max.
  // That is what the parity targets need (spectra / truncation errors to 1e-8) and what the sweeps use.
  // MPSK_FLAG_SVD_RELATIVE continues down to the rounding floor (up to 6 levels) for full relative accuracy.
  const int MAXLEV = (ctx->flags & MPSK_FLAG_SVD_RELATIVE) ? 6 : 1;
  constexpr double SPLIT = 1e-3;       // singular values within this factor of the level's largest are emitted
  std::vector<int> active = fastl;
  for (int lev = 0; lev < MAXLEV + 1 && !active.empty(); ++lev) {
    std::vector<CopyTask> ct; std::vector<GemmProblemHost> g1;
    for (int si : active) {
      Coupled* c = secs[si]; GramState& g = gs[si];
      const int mm = c->mm, nl = g.nl;
      const int64_t ldr = pad(nl);
      ct.push_back(CopyTask{g.M, g.Mconj, 1, g.ldm, 1, g.ldm, mm, nl, cx ? 1 : 0, 0, nullptr});      // conj(M), same layout
      ct.push_back(CopyTask{g.M, g.Mrow, 1, g.ldm, ldr, 1, mm, nl, 0, 0, nullptr});                  // M row-major
      if (g.R) {
        ct.push_back(CopyTask{g.R, g.Rrow, 1, g.ldn, ldr, 1, c->nn, nl, 0, 0, nullptr});
      }
      // G (column-major, ld pad(nl)) = row-major buffer C[i][j] = sum_k M[k,i] conj(M[k,j])
      g1.push_back(GemmProblemHost{g.M, g.Mconj, g.G, g.ldm, g.ldm, ldr, nl, nl, mm, 0});
    }
    MPSK_TRY(run_copy_tasks(ctx, ct, cx));
    {
      GemmPlan p1;
      MPSK_TRY(gemm_plan_build(ctx, cx, g1, &p1));
      MPSK_TRY(gemm_plan_run(ctx, &p1));
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    lap("gram", lev, gs[active[0]].nl);
    MPSK_TRY(run_phase(active, "heevd",
      [&](int si, cudaStream_t, cusolverDnHandle_t hs, int* d_info) {
        GramState& g = gs[si]; const int nl = g.nl; const int ldg = (int)pad(nl);
        if (cx) return cusolverDnZheevd(hs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, nl, (cuDoubleComplex*)g.G, ldg, g.lam, (cuDoubleComplex*)g.work, g.lwork, d_info);
        return cusolverDnDsyevd(hs, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, nl, (double*)g.G, ldg, g.lam, (double*)g.work, g.lwork, d_info);
      },
      [&](int si, cudaStream_t st) { GramState& g = gs[si]; g.hlam.resize(g.nl); cudaMemcpyAsync(g.hlam.data(), g.lam, sizeof(double) * g.nl, cudaMemcpyDeviceToHost, st); }));
    ctx->launches += (int64_t)active.size();
    lap("heevd", lev, (int)active.size());
    // split decisions (host): emit everything within SPLIT of the level's largest singular value; at the rounding floor
    // of the original matrix (or at the last level) emit everything
    std::vector<GemmProblemHost> g2;
    for (int si : active) {
      Coupled* c = secs[si]; GramState& g = gs[si];
      const int nl = g.nl;
      const double ltop = std::max(g.hlam[nl - 1], 0.0);
      if (lev == 0) g.lam0 = ltop;
      const double floor2 = g.lam0 * (double)c->mm * 1e-30;          // (eps sqrt(m) sigma_max)^2 up to a small factor
      int t = 0;
      if (lev < MAXLEV && ltop > floor2) {
        const double cut = std::max(ltop * SPLIT * SPLIT, floor2);
        while (t < nl && g.hlam[t] < cut) ++t;
      }
      g.split = t;
      const int64_t ldr = pad(nl);
      void* Bn = g.buf[lev & 1];
      // Bn (column-major mm x nl, ld ldm) = row-major C[j][k] = sum_i Vt[j][i] Mrow[k][i]
      g2.push_back(GemmProblemHost{g.G, g.Mrow, Bn, ldr, ldr, g.ldm, nl, c->mm, nl, 0});
      if (g.R) {
        void* Rn = g.rbuf[lev & 1];
        g2.push_back(GemmProblemHost{g.G, g.Rrow, Rn, ldr, ldr, g.ldn, nl, c->nn, nl, 0});
      }
    }
    GemmPlan p2;
    MPSK_TRY(gemm_plan_build(ctx, cx, g2, &p2));
    MPSK_TRY(gemm_plan_run(ctx, &p2));
    std::vector<int> next_active;
    for (int si : active) {
      Coupled* c = secs[si]; GramState& g = gs[si];
      const int nl = g.nl, t = g.split;
      const int64_t ldr = pad(nl);
      void* Bn = g.buf[lev & 1];
      const void* Rn = g.R ? g.rbuf[lev & 1] : g.G;          // level 0: R V = V
      const int64_t ldrn = g.R ? g.ldn : ldr;
      const int64_t work_items = (int64_t)c->mm * (nl - t);
      const int grid = (int)std::min<int64_t>((work_items + 255) / 256 + 1, 8192);
      if (nl - t > 0) {
        if (cx) gram_colnorm_kernel<true><<<nl - t, 256, 0, ctx->stream>>>((const double*)Bn, g.ldm, c->mm, t, nl, g.outpos, c->S);
        else gram_colnorm_kernel<false><<<nl - t, 256, 0, ctx->stream>>>((const double*)Bn, g.ldm, c->mm, t, nl, g.outpos, c->S);
        if (cx) gram_emit_kernel<true><<<grid, 256, 0, ctx->stream>>>((const double*)Bn, g.ldm, (const double*)Rn, ldrn, c->mm, c->nn, t, nl, g.outpos,
                                                                        (double*)c->U, c->ldu, (double*)c->VT, c->ldvt, c->S);
        else gram_emit_kernel<false><<<grid, 256, 0, ctx->stream>>>((const double*)Bn, g.ldm, (const double*)Rn, ldrn, c->mm, c->nn, t, nl, g.outpos,
                                                                     (double*)c->U, c->ldu, (double*)c->VT, c->ldvt, c->S);
        ctx->launches++;
      }
      ctx->launches++;
      g.outpos += nl - t;
      if (t > 0) {
        if (!g.R) {      // level 0: the tail of V becomes the accumulated right factor (G is recycled by the next level)
          CUDA_TRY(ctx, cudaMemcpyAsync(g.rbuf[0], g.G, (size_t)ldr * t * es, cudaMemcpyDeviceToDevice, ctx->stream));
          g.R = g.rbuf[0];
          if (ldr != g.ldn) return ctx->fail(MPSK_EINTERNAL, "tsvd: leading dimension mismatch");
        } else g.R = g.rbuf[lev & 1];
        g.M = Bn; g.nl = t;
        next_active.push_back(si);
      }
    }
    CUDA_TRY(ctx, cudaGetLastError());
    active.swap(next_active);
    lap("project+emit", lev, (int)active.size());
  }
  if (!fastl.empty()) {
    for (int si : fastl) { sv[si].resize(secs[si]->k); CUDA_TRY(ctx, cudaMemcpyAsync(sv[si].data(), secs[si]->S, sizeof(double) * secs[si]->k, cudaMemcpyDeviceToHost, ctx->stream)); }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    // re-orthonormalise U: Householder QR with positive diagonal (exact isometry; changes U by O(eps sigma_max / sigma))
    MPSK_TRY(run_phase(fastl, "QR clean-up",
      [&](int si, cudaStream_t st, cusolverDnHandle_t hs, int* d_info) {
        Coupled* c = secs[si]; GramState& g = gs[si];
        const int mm = c->mm, nn = c->nn; const int ldu = (int)c->ldu;
        cusolverStatus_t cs_;
        if (cx) cs_ = cusolverDnZgeqrf(hs, mm, nn, (cuDoubleComplex*)c->U, ldu, (cuDoubleComplex*)g.tau, (cuDoubleComplex*)g.work2, g.lwork2, d_info);
        else cs_ = cusolverDnDgeqrf(hs, mm, nn, (double*)c->U, ldu, (double*)g.tau, (double*)g.work2, g.lwork2, d_info);
        const int grid = (int)std::min<int64_t>(((int64_t)nn * nn + 255) / 256, 4096);
        if (cx) qr_phase_kernel<true><<<grid, 256, 0, st>>>((const double*)c->U, ldu, nn, nn, g.ph, (double*)g.Rtmp, 1, nn, 0);
        else qr_phase_kernel<false><<<grid, 256, 0, st>>>((const double*)c->U, ldu, nn, nn, g.ph, (double*)g.Rtmp, 1, nn, 0);
        if (cs_ == CUSOLVER_STATUS_SUCCESS) {
          if (cx) cs_ = cusolverDnZungqr(hs, mm, nn, nn, (cuDoubleComplex*)c->U, ldu, (cuDoubleComplex*)g.tau, (cuDoubleComplex*)g.work2, g.lwork2, d_info);
          else cs_ = cusolverDnDorgqr(hs, mm, nn, nn, (double*)c->U, ldu, (double*)g.tau, (double*)g.work2, g.lwork2, d_info);
        }
        const int grid2 = (int)std::min<int64_t>(((int64_t)mm * nn + 255) / 256, 4096);
        if (cx) q_phase_kernel<true><<<grid2, 256, 0, st>>>((double*)c->U, ldu, mm, nn, g.ph);
        else q_phase_kernel<false><<<grid2, 256, 0, st>>>((double*)c->U, ldu, mm, nn, g.ph);
        return cs_;
      }, nullptr));
    ctx->launches += 4 * (int64_t)fastl.size();
    lap("qr clean-up", (int)fastl.size(), 0);
  }
  // ---- singular values of all sectors on every rank
  if (dist) {
    double* d_all = (double*)ctx->alloc(sizeof(double) * std::max<int64_t>(ktot, 1));
    if (!d_all) return MPSK_ENOMEM;
    CUDA_TRY(ctx, cudaMemsetAsync(d_all, 0, sizeof(double) * std::max<int64_t>(ktot, 1), ctx->stream));
    for (int si = 0; si < ns; ++si)
      if (secs[si]->mine) CUDA_TRY(ctx, cudaMemcpyAsync(d_all + koff[si], secs[si]->S, sizeof(double) * secs[si]->k, cudaMemcpyDeviceToDevice, ctx->stream));
    int st = comm_allreduce_sum(ctx, d_all, (size_t)ktot);
    std::vector<double> all((size_t)ktot);
    if (st == MPSK_OK && ktot > 0) {
      cudaMemcpyAsync(all.data(), d_all, sizeof(double) * ktot, cudaMemcpyDeviceToHost, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
    }
    ctx->free(d_all);
    MPSK_TRY(st);
    for (int si = 0; si < ns; ++si) sv[si].assign(all.begin() + koff[si], all.begin() + koff[si] + secs[si]->k);
  }
  // ---- global truncation select (same rule as oracle/linalg.py:tsvd): drop smallest first; ties: later sector, larger index first
  struct Ent { double s; int sec, idx; };
  std::vector<Ent> all;
  for (int si = 0; si < ns; ++si) { secs[si]->keep = (int)secs[si]->k; for (int i = 0; i < secs[si]->k; ++i) all.push_back(Ent{sv[si][i], si, i}); }
  std::stable_sort(all.begin(), all.end(), [](const Ent& a, const Ent& b) {
    if (a.s != b.s) return a.s < b.s;
    if (a.sec != b.sec) return a.sec > b.sec;
    return a.idx > b.idx;
  });
  double dropped2 = 0.0; size_t ndrop = 0;
  if (trunc.kind == MPSK_TRUNCDIM) {
    const int64_t D = (int64_t)trunc.value;
    ndrop = (int64_t)all.size() > D ? all.size() - (size_t)D : 0;
  } else if (trunc.kind == MPSK_TRUNCERR) {
    double acc = 0.0;
    for (const Ent& e : all) { if (std::sqrt(acc + e.s * e.s) <= trunc.value) { acc += e.s * e.s; ++ndrop; } else break; }
  } else if (trunc.kind == MPSK_TRUNCBELOW) {
    for (const Ent& e : all) { if (e.s < trunc.value) ++ndrop; else break; }
  } else if (trunc.kind != MPSK_NOTRUNC) return ctx->fail(MPSK_EINVAL, "unknown truncation scheme");
  for (size_t i = 0; i < ndrop; ++i) { secs[all[i].sec]->keep--; dropped2 += all[i].s * all[i].s; }
  if (eps_out) *eps_out = std::sqrt(dropped2);
  double kept2 = 0.0;
  for (int si = 0; si < ns; ++si) for (int i = 0; i < secs[si]->keep; ++i) kept2 += sv[si][i] * sv[si][i];
  if (kept2 <= 0.0) return ctx->fail(MPSK_EINVAL, "tsvd: truncation removed every singular value");
  const double snorm = normalize ? 1.0 / std::sqrt(kept2) : 1.0;
  // record the spectrum for mpsk_tsvd_spectrum
  TsvdSpectrum& sp = ctx->last_spectrum;
  sp.charge.clear(); sp.count.clear(); sp.kept.clear(); sp.values.clear();
  for (int si = 0; si < ns; ++si) {
    sp.charge.push_back(secs[si]->Q); sp.count.push_back((int32_t)secs[si]->k); sp.kept.push_back(secs[si]->keep);
    sp.values.insert(sp.values.end(), sv[si].begin(), sv[si].end());
  }
  // ---- new bond space and outputs
  Space Vn; Vn.mod = Vl.mod;
  for (Coupled* c : secs) if (c->keep > 0) { Vn.q.push_back(c->Q); Vn.n.push_back(c->keep); }
  Tensor *Ut = nullptr, *St = nullptr, *Vt = nullptr;
  MPSK_TRY(tensor_create(ctx, Vl, s1, Vn, x->dtype, F_KL, true, &Ut));
  std::unique_ptr<Tensor> ug(Ut);
  MPSK_TRY(tensor_create(ctx, Vn, std::vector<int32_t>{0}, Vn, x->dtype, F_KL, true, &St));
  std::unique_ptr<Tensor> sg(St);
  MPSK_TRY(tensor_create(ctx, Vn, s2, Vr, x->dtype, F_KL, true, &Vt));
  std::unique_ptr<Tensor> vg(Vt);
  tasks.clear();
  for (Coupled* c : secs) {
    if (c->keep <= 0 || !c->mine) continue;
    const int jn = Vn.find(c->Q);
    for (size_t ri = 0; ri < c->rows.size(); ++ri) {
      const int jl = c->rows[ri].first, u = c->rows[ri].second;
      const Blk* b = Ut->find(u, jl);
      if (!b || b->jr != jn) continue;
      CopyTask tk; tk.dst = Ut->ptr(b->off); tk.dsl = b->sl; tk.dsr = b->sr; tk.nl = (int32_t)Vl.n[jl]; tk.nr = c->keep; tk.scale_axis = 0; tk.scale = nullptr;
      if (!c->flip) { tk.src = (char*)c->U + c->rowoff[ri] * es; tk.ssl = 1; tk.ssr = c->ldu; tk.conj = 0; }
      else { tk.src = (char*)c->VT + (c->rowoff[ri] * c->ldvt) * es; tk.ssl = c->ldvt; tk.ssr = 1; tk.conj = cx ? 1 : 0; }
      tasks.push_back(tk);
    }
    for (size_t ci = 0; ci < c->cols.size(); ++ci) {
      const int jr = c->cols[ci].first, t = c->cols[ci].second;
      const Blk* b = Vt->find(t, jn);
      if (!b || b->jr != jr) continue;
      CopyTask tk; tk.dst = Vt->ptr(b->off); tk.dsl = b->sl; tk.dsr = b->sr; tk.nl = c->keep; tk.nr = (int32_t)Vr.n[jr]; tk.scale_axis = 0; tk.scale = nullptr;
      if (!c->flip) { tk.src = (char*)c->VT + (c->coloff[ci] * c->ldvt) * es; tk.ssl = 1; tk.ssr = c->ldvt; tk.conj = 0; }
      else { tk.src = (char*)c->U + c->coloff[ci] * es; tk.ssl = c->ldu; tk.ssr = 1; tk.conj = cx ? 1 : 0; }
      tasks.push_back(tk);
    }
    const Blk* sb = St->find(0, jn);
    diag_kernel<<<(c->keep + 127) / 128, 128, 0, ctx->stream>>>((double*)St->ptr(sb->off), St->groups[jn].ld,
                                                                 c->keep, c->S, snorm, cx ? 1 : 0);
    ctx->launches++;
  }
  MPSK_TRY(run_copy_tasks(ctx, tasks, cx));
  if (dist) {   // every element of the outputs was written by exactly one rank (zero elsewhere): a sum assembles them
    const size_t dpe = cx ? 2 : 1;
    MPSK_TRY(comm_allreduce_sum(ctx, Ut->d, (size_t)Ut->nelem * dpe));
    MPSK_TRY(comm_allreduce_sum(ctx, St->d, (size_t)St->nelem * dpe));
    MPSK_TRY(comm_allreduce_sum(ctx, Vt->d, (size_t)Vt->nelem * dpe));
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // the pool streams may reuse this memory next
  lap("outputs", 0, 0);
  if (ctx->profile) fprintf(stderr, "[mpsk profile] tsvd total %.2f ms\n", (tnow() - tp0) * 1e3);
  *Uo = ug.release(); *So = sg.release(); *Vo = vg.release();
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ inverse of a bond matrix
// inv(C) per sector (idmrg.jl:130-134,172-173: the edge updates of IDMRG2 divide by the bond matrix): LU with partial
// pivoting (cuSOLVER getrf / getrs on an identity right-hand side; LAPACK-class library call as in the reference).
__global__ void eye_kernel(double* dst, int64_t ld, int n, int cplx) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)n * n; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % n), j = (int)(e / n);
    const double v = i == j ? 1.0 : 0.0;
    if (cplx) { dst[2 * (i + j * ld)] = v; dst[2 * (i + j * ld) + 1] = 0.0; } else dst[i + j * ld] = v;
  }
}

int bond_inverse(Ctx* ctx, const Tensor* C, Tensor** out) {
  if (C->nm() != 1 || C->mch[0] != 0) return ctx->fail(MPSK_ESPACE, "bond_inverse: not a bond tensor");
  const bool cx = C->dtype == MPSK_C128;
  const size_t es = cx ? 16 : 8;
  Tensor* o = nullptr;
  MPSK_TRY(tensor_create(ctx, C->Vr, std::vector<int32_t>{0}, C->Vl, C->dtype, F_KL, true, &o));
  std::unique_ptr<Tensor> og(o);
  cusolverDnHandle_t hs;
  MPSK_TRY(get_cusolver(ctx, &hs));
  int* d_info = (int*)ctx->alloc(sizeof(int));
  if (!d_info) return MPSK_ENOMEM;
  struct Tmp { Ctx* c; std::vector<void*> p; ~Tmp() { for (void* q : p) c->free(q); } } tmp{ctx, {d_info}};
  for (const Blk& b : C->blks) {
    const int n = (int)C->Vl.n[b.jl];
    if ((int)C->Vr.n[b.jr] != n) return ctx->fail(MPSK_ESPACE, "bond_inverse: sector %d is not square", (int)C->Vl.q[b.jl]);
    const Blk* ob = o->find(0, b.jr);
    if (!ob || ob->jr != b.jl) return ctx->fail(MPSK_EINTERNAL, "bond_inverse: block mismatch");
    const int64_t ld = pad(n);
    void* W = ctx->alloc((size_t)ld * n * es); void* B = ctx->alloc((size_t)ld * n * es); int* piv = (int*)ctx->alloc(sizeof(int) * n);
    if (!W || !B || !piv) return MPSK_ENOMEM;
    tmp.p.push_back(W); tmp.p.push_back(B); tmp.p.push_back(piv);
    std::vector<CopyTask> in{CopyTask{C->ptr(b.off), W, b.sl, b.sr, 1, ld, n, n, 0, 0, nullptr}};      // W(l, r) column-major
    MPSK_TRY(run_copy_tasks(ctx, in, cx));
    eye_kernel<<<std::min((n * n + 255) / 256, 1024), 256, 0, ctx->stream>>>((double*)B, ld, n, cx ? 1 : 0);
    int lw = 0;
    if (cx) CUSOLVER_TRY(ctx, cusolverDnZgetrf_bufferSize(hs, n, n, (cuDoubleComplex*)W, (int)ld, &lw));
    else CUSOLVER_TRY(ctx, cusolverDnDgetrf_bufferSize(hs, n, n, (double*)W, (int)ld, &lw));
    void* work = ctx->alloc((size_t)std::max(lw, 1) * es);
    if (!work) return MPSK_ENOMEM;
    tmp.p.push_back(work);
    if (cx) { CUSOLVER_TRY(ctx, cusolverDnZgetrf(hs, n, n, (cuDoubleComplex*)W, (int)ld, (cuDoubleComplex*)work, piv, d_info));
              CUSOLVER_TRY(ctx, cusolverDnZgetrs(hs, CUBLAS_OP_N, n, n, (const cuDoubleComplex*)W, (int)ld, piv, (cuDoubleComplex*)B, (int)ld, d_info)); }
    else { CUSOLVER_TRY(ctx, cusolverDnDgetrf(hs, n, n, (double*)W, (int)ld, (double*)work, piv, d_info));
           CUSOLVER_TRY(ctx, cusolverDnDgetrs(hs, CUBLAS_OP_N, n, n, (const double*)W, (int)ld, piv, (double*)B, (int)ld, d_info)); }
    std::vector<CopyTask> outt{CopyTask{B, o->ptr(ob->off), 1, ld, ob->sl, ob->sr, n, n, 0, 0, nullptr}};
    MPSK_TRY(run_copy_tasks(ctx, outt, cx));
    ctx->launches += 3;
  }
  CUDA_TRY(ctx, cudaGetLastError());
  *out = og.release();
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ leftnull / rightnull
// leftnull(A) (left = true): isometry N[l m; k'] spanning the orthogonal complement of the range of A[l m; r];
// rightnull(A) (left = false): N[k' m; r] with orthonormal rows spanning the complement of the row space of A[l; m r]
// (changebonds/optimalexpand.jl:28-29,95-96).  Per coupled sector a full Householder QR (cuSOLVER geqrf + ungqr generating
// all M columns); coupled sectors of the fused (bond x label) space that are absent from the other bond are null entirely.
int nullspace(Ctx* ctx, const Tensor* A, bool left, Tensor** out) {
  const bool cx = A->dtype == MPSK_C128;
  const size_t es = cx ? 16 : 8;
  const Space& Vg = left ? A->Vr : A->Vl;       // the bond the factorisation removes (columns)
  const Space& Vo = left ? A->Vl : A->Vr;       // the bond that stays (fused with the labels into the rows)
  struct Row { int m, jo; int64_t off; };
  struct Sec { int64_t Q = 0; std::vector<Row> rows; int64_t M = 0, n = 0, k = 0; int jg = -1; void* W = nullptr; void* tau = nullptr; };
  std::map<int64_t, Sec> smap;
  for (int mm = 0; mm < A->nm(); ++mm)
    for (int jo = 0; jo < Vo.nsec(); ++jo) {
      const int64_t Q = Vo.norm(left ? (int64_t)Vo.q[jo] + A->mch[mm] : (int64_t)Vo.q[jo] - A->mch[mm]);
      Sec& sc = smap[Q]; sc.Q = Q;
      sc.rows.push_back(Row{mm, jo, sc.M}); sc.M += Vo.n[jo];
    }
  Space Vn; Vn.mod = Vg.mod;
  std::vector<Sec*> secs;
  for (auto& kv : smap) {
    Sec& sc = kv.second;
    sc.jg = Vg.find(sc.Q); sc.n = sc.jg >= 0 ? Vg.n[sc.jg] : 0; sc.k = std::min(sc.M, sc.n);
    if (sc.M - sc.k > 0) { Vn.q.push_back((int32_t)sc.Q); Vn.n.push_back(sc.M - sc.k); secs.push_back(&sc); }
  }
  Tensor* N = nullptr;
  if (left) MPSK_TRY(tensor_create(ctx, A->Vl, A->mch, Vn, A->dtype, F_KL, true, &N));
  else MPSK_TRY(tensor_create(ctx, Vn, A->mch, A->Vr, A->dtype, F_KL, true, &N));
  std::unique_ptr<Tensor> ng(N);
  struct Cleanup { Ctx* c; std::vector<Sec*>* v; ~Cleanup() { for (Sec* sc : *v) { c->free(sc->W); c->free(sc->tau); } } } cl{ctx, &secs};
  cusolverDnHandle_t hs;
  MPSK_TRY(get_cusolver(ctx, &hs));
  int* d_info = (int*)ctx->alloc(sizeof(int));
  if (!d_info) return MPSK_ENOMEM;
  struct InfoFree { Ctx* c; void* p; ~InfoFree() { c->free(p); } } inf{ctx, d_info};
  std::vector<CopyTask> outt;
  for (Sec* sp : secs) {
    Sec& sc = *sp;
    const int M = (int)sc.M, n = (int)sc.n, k = (int)sc.k;
    sc.W = ctx->alloc((size_t)M * M * es); sc.tau = ctx->alloc((size_t)std::max(k, 1) * es);
    if (!sc.W || !sc.tau) return MPSK_ENOMEM;
    if (n == 0) {
      eye_kernel<<<std::min((int)(((int64_t)M * M + 255) / 256), 4096), 256, 0, ctx->stream>>>((double*)sc.W, M, M, cx ? 1 : 0);
    } else {
      CUDA_TRY(ctx, cudaMemsetAsync(sc.W, 0, (size_t)M * M * es, ctx->stream));
      std::vector<CopyTask> in;
      for (const Row& rw : sc.rows) {
        const Blk* b = left ? A->find(rw.m, rw.jo) : A->find(rw.m, sc.jg);
        if (!b) return ctx->fail(MPSK_EINTERNAL, "nullspace: missing block");
        CopyTask tk; tk.src = A->ptr(b->off); tk.ssl = b->sl; tk.ssr = b->sr; tk.nl = (int32_t)A->Vl.n[b->jl]; tk.nr = (int32_t)A->Vr.n[b->jr];
        tk.scale_axis = 0; tk.scale = nullptr; tk.dst = (char*)sc.W + rw.off * es;
        if (left) { tk.dsl = 1; tk.dsr = M; tk.conj = 0; } else { tk.dsl = M; tk.dsr = 1; tk.conj = cx ? 1 : 0; }
        in.push_back(tk);
      }
      MPSK_TRY(run_copy_tasks(ctx, in, cx));
      int lw1 = 0, lw2 = 0;
      if (cx) { CUSOLVER_TRY(ctx, cusolverDnZgeqrf_bufferSize(hs, M, n, (cuDoubleComplex*)sc.W, M, &lw1));
                CUSOLVER_TRY(ctx, cusolverDnZungqr_bufferSize(hs, M, M, k, (cuDoubleComplex*)sc.W, M, (cuDoubleComplex*)sc.tau, &lw2)); }
      else { CUSOLVER_TRY(ctx, cusolverDnDgeqrf_bufferSize(hs, M, n, (double*)sc.W, M, &lw1));
             CUSOLVER_TRY(ctx, cusolverDnDorgqr_bufferSize(hs, M, M, k, (double*)sc.W, M, (double*)sc.tau, &lw2)); }
      const int lw = std::max(std::max(lw1, lw2), 1);
      void* work = ctx->alloc((size_t)lw * es);
      if (!work) return MPSK_ENOMEM;
      if (cx) { CUSOLVER_TRY(ctx, cusolverDnZgeqrf(hs, M, n, (cuDoubleComplex*)sc.W, M, (cuDoubleComplex*)sc.tau, (cuDoubleComplex*)work, lw, d_info));
                CUSOLVER_TRY(ctx, cusolverDnZungqr(hs, M, M, k, (cuDoubleComplex*)sc.W, M, (cuDoubleComplex*)sc.tau, (cuDoubleComplex*)work, lw, d_info)); }
      else { CUSOLVER_TRY(ctx, cusolverDnDgeqrf(hs, M, n, (double*)sc.W, M, (double*)sc.tau, (double*)work, lw, d_info));
             CUSOLVER_TRY(ctx, cusolverDnDorgqr(hs, M, M, k, (double*)sc.W, M, (double*)sc.tau, (double*)work, lw, d_info)); }
      ctx->free(work);
      ctx->launches += 2;
    }
    const int jn = Vn.find(sc.Q);
    const int nd = M - k;
    for (const Row& rw : sc.rows) {
      CopyTask tk; tk.scale_axis = 0; tk.scale = nullptr;
      tk.src = (char*)sc.W + (rw.off + (int64_t)k * M) * es;          // W(rw.off + i, k + c)
      if (left) {
        const Blk* b = N->find(rw.m, rw.jo);
        if (!b || b->jr != jn) return ctx->fail(MPSK_EINTERNAL, "leftnull: output block mismatch");
        tk.dst = N->ptr(b->off); tk.dsl = b->sl; tk.dsr = b->sr; tk.nl = (int32_t)Vo.n[rw.jo]; tk.nr = nd; tk.ssl = 1; tk.ssr = M; tk.conj = 0;
      } else {
        const Blk* b = N->find(rw.m, jn);
        if (!b || b->jr != rw.jo) return ctx->fail(MPSK_EINTERNAL, "rightnull: output block mismatch");
        tk.dst = N->ptr(b->off); tk.dsl = b->sl; tk.dsr = b->sr; tk.nl = nd; tk.nr = (int32_t)Vo.n[rw.jo]; tk.ssl = M; tk.ssr = 1; tk.conj = cx ? 1 : 0;
      }
      outt.push_back(tk);
    }
  }
  MPSK_TRY(run_copy_tasks(ctx, outt, cx));
  CUDA_TRY(ctx, cudaGetLastError());
  *out = ng.release();
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ QRpos / LQpos
// leftorth!(A; alg=QRpos()) and rightorth!(A; alg=LQpos()) (orthoview.jl:49-60, ortho.jl:127-138, finitemps.jl:150):
// per coupled sector one Householder QR (cuSOLVER geqrf/ungqr: LAPACK-class library call as in the reference), then
// the diagonal of R is made real non-negative by a phase kernel (unique factorisation -> elementwise parity).

struct QrSec { int jg; int64_t m, n, k; std::vector<std::pair<int, int64_t>> rows; void* W = nullptr; void* tau = nullptr; double* ph = nullptr; };

// left = true : A[l m; r] = Q[l m; k] R[k; r]     (groups = r sectors, rows (m, l))
// left = false: A[l; m r] = L[l; k] Q[k m; r]     (groups = l sectors, rows (m, r) of A^H)
int qrpos(Ctx* ctx, const Tensor* A, bool left, Tensor** Qo, Tensor** Ro) {
  const bool cx = A->dtype == MPSK_C128;
  const size_t es = cx ? 16 : 8;
  const Space& Vg = left ? A->Vr : A->Vl;       // grouping (column) space
  const Space& Vo = left ? A->Vl : A->Vr;       // the other bond
  const int ng = Vg.nsec(), nm = A->nm();
  std::vector<QrSec> secs;
  for (int g = 0; g < ng; ++g) {
    QrSec s; s.jg = g; s.m = 0; s.n = Vg.n[g];
    for (int mm = 0; mm < nm; ++mm) {
      const int jo = left ? Vo.find((int64_t)Vg.q[g] - A->mch[mm]) : Vo.find((int64_t)Vg.q[g] + A->mch[mm]);
      if (jo < 0) continue;
      s.rows.push_back({mm, s.m}); s.m += Vo.n[jo];
    }
    s.k = std::min(s.m, s.n);
    if (s.k > 0) secs.push_back(std::move(s));
  }
  struct Cleanup { Ctx* c; std::vector<QrSec>* v; ~Cleanup() { for (auto& s : *v) { c->free(s.W); c->free(s.tau); c->free(s.ph); } } } cl{ctx, &secs};
  std::vector<CopyTask> tasks;
  for (QrSec& s : secs) {
    s.W = ctx->alloc(s.m * s.n * es); s.tau = ctx->alloc(s.k * es); s.ph = (double*)ctx->alloc(s.k * es);
    if (!s.W || !s.tau || !s.ph) return MPSK_ENOMEM;
    for (auto& rw : s.rows) {
      const int mm = rw.first;
      const int jo = left ? Vo.find((int64_t)Vg.q[s.jg] - A->mch[mm]) : Vo.find((int64_t)Vg.q[s.jg] + A->mch[mm]);
      const Blk* b = left ? A->find(mm, jo) : A->find(mm, s.jg);
      if (!b) return ctx->fail(MPSK_EINTERNAL, "qrpos: missing block");
      CopyTask tk; tk.src = A->ptr(b->off); tk.ssl = b->sl; tk.ssr = b->sr; tk.nl = (int32_t)A->Vl.n[b->jl]; tk.nr = (int32_t)A->Vr.n[b->jr];
      tk.scale_axis = 0; tk.scale = nullptr;
      tk.dst = (char*)s.W + rw.second * es;
      if (left) { tk.dsl = 1; tk.dsr = s.m; tk.conj = 0; }            // W(row = off + il, col = ir)
      else { tk.dsl = s.m; tk.dsr = 1; tk.conj = cx ? 1 : 0; }         // W(row = off + ir, col = il) = conj(A)
      tasks.push_back(tk);
    }
  }
  MPSK_TRY(run_copy_tasks(ctx, tasks, cx));
  cusolverDnHandle_t hs;
  MPSK_TRY(get_cusolver(ctx, &hs));
  int* d_info = (int*)ctx->alloc(sizeof(int));
  // new bond space
  Space Vn; Vn.mod = Vg.mod;
  for (QrSec& s : secs) { Vn.q.push_back(Vg.q[s.jg]); Vn.n.push_back(s.k); }
  Tensor *Qt = nullptr, *Rt = nullptr;
  if (left) {
    MPSK_TRY(tensor_create(ctx, A->Vl, A->mch, Vn, A->dtype, F_KL, true, &Qt));
  } else {
    MPSK_TRY(tensor_create(ctx, Vn, A->mch, A->Vr, A->dtype, F_KL, true, &Qt));
  }
  std::unique_ptr<Tensor> qg(Qt);
  if (left) MPSK_TRY(tensor_create(ctx, Vn, std::vector<int32_t>{0}, A->Vr, A->dtype, F_KL, true, &Rt));
  else MPSK_TRY(tensor_create(ctx, A->Vl, std::vector<int32_t>{0}, Vn, A->dtype, F_KL, true, &Rt));
  std::unique_ptr<Tensor> rg(Rt);
  tasks.clear();
  for (QrSec& s : secs) {
    const int m = (int)s.m, n = (int)s.n, k = (int)s.k;
    int lw1 = 0, lw2 = 0;
    if (cx) { CUSOLVER_TRY(ctx, cusolverDnZgeqrf_bufferSize(hs, m, n, (cuDoubleComplex*)s.W, m, &lw1));
              CUSOLVER_TRY(ctx, cusolverDnZungqr_bufferSize(hs, m, k, k, (cuDoubleComplex*)s.W, m, (cuDoubleComplex*)s.tau, &lw2)); }
    else { CUSOLVER_TRY(ctx, cusolverDnDgeqrf_bufferSize(hs, m, n, (double*)s.W, m, &lw1));
           CUSOLVER_TRY(ctx, cusolverDnDorgqr_bufferSize(hs, m, k, k, (double*)s.W, m, (double*)s.tau, &lw2)); }
    const int lw = std::max(std::max(lw1, lw2), 1);
    void* work = ctx->alloc((size_t)lw * es);
    if (!work) return MPSK_ENOMEM;
    if (cx) CUSOLVER_TRY(ctx, cusolverDnZgeqrf(hs, m, n, (cuDoubleComplex*)s.W, m, (cuDoubleComplex*)s.tau, (cuDoubleComplex*)work, lw, d_info));
    else CUSOLVER_TRY(ctx, cusolverDnDgeqrf(hs, m, n, (double*)s.W, m, (double*)s.tau, (double*)work, lw, d_info));
    // R (k x n) with positive diagonal -> straight into the bond tensor
    const int jn = Vn.find(Vg.q[s.jg]);
    double* rdst; int64_t si, sj;
    if (left) { const Blk* rb = Rt->find(0, jn); rdst = (double*)Rt->ptr(rb->off); si = rb->sl; sj = rb->sr; }       // R[k; r]: l = k idx, r = col
    else { const Blk* rb = Rt->find(0, s.jg); rdst = (double*)Rt->ptr(rb->off); si = rb->sr; sj = rb->sl; }         // L[l; k] = conj(R'[k, l])
    const int grid = (int)std::min<int64_t>(((int64_t)k * n + 255) / 256, 4096);
    if (cx) qr_phase_kernel<true><<<grid, 256, 0, ctx->stream>>>((const double*)s.W, m, k, n, s.ph, rdst, si, sj, left ? 0 : 1);
    else qr_phase_kernel<false><<<grid, 256, 0, ctx->stream>>>((const double*)s.W, m, k, n, s.ph, rdst, si, sj, 0);
    if (cx) CUSOLVER_TRY(ctx, cusolverDnZungqr(hs, m, k, k, (cuDoubleComplex*)s.W, m, (cuDoubleComplex*)s.tau, (cuDoubleComplex*)work, lw, d_info));
    else CUSOLVER_TRY(ctx, cusolverDnDorgqr(hs, m, k, k, (double*)s.W, m, (double*)s.tau, (double*)work, lw, d_info));
    const int grid2 = (int)std::min<int64_t>(((int64_t)m * k + 255) / 256, 4096);
    if (cx) q_phase_kernel<true><<<grid2, 256, 0, ctx->stream>>>((double*)s.W, m, m, k, s.ph);
    else q_phase_kernel<false><<<grid2, 256, 0, ctx->stream>>>((double*)s.W, m, m, k, s.ph);
    ctx->launches += 4;
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->free(work);
    // Q blocks
    for (auto& rw : s.rows) {
      const int mm = rw.first;
      CopyTask tk; tk.scale_axis = 0; tk.scale = nullptr; tk.src = (char*)s.W + rw.second * es;
      if (left) {
        const int jo = Vo.find((int64_t)Vg.q[s.jg] - A->mch[mm]);
        const Blk* b = Qt->find(mm, jo);
        if (!b || b->jr != jn) return ctx->fail(MPSK_EINTERNAL, "qrpos: output block mismatch");
        tk.dst = Qt->ptr(b->off); tk.dsl = b->sl; tk.dsr = b->sr; tk.nl = (int32_t)Vo.n[jo]; tk.nr = k; tk.ssl = 1; tk.ssr = m; tk.conj = 0;
      } else {
        const Blk* b = Qt->find(mm, jn);
        if (!b) return ctx->fail(MPSK_EINTERNAL, "lqpos: output block mismatch");
        tk.dst = Qt->ptr(b->off); tk.dsl = b->sl; tk.dsr = b->sr; tk.nl = k; tk.nr = (int32_t)A->Vr.n[b->jr]; tk.ssl = m; tk.ssr = 1; tk.conj = cx ? 1 : 0;
      }
      tasks.push_back(tk);
    }
  }
  ctx->free(d_info);
  MPSK_TRY(run_copy_tasks(ctx, tasks, cx));
  *Qo = qg.release(); *Ro = rg.release();
  return MPSK_OK;
}

}  // namespace mpsk
