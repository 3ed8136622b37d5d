// Device-resident Krylov solvers behind fixedpoint (src/algorithms/fixedpoint.jl:9-30 -> KrylovKit eigsolve /
// schursolve), integrate (src/algorithms/timestep/integrators.jl:20-25 -> KrylovKit exponentiate) and the
// environment linear systems (src/environments/mpohaminfenv.jl:95,113,146,164 -> KrylovKit linsolve, GMRES).
// The Krylov basis lives in HBM; every step is one operator application (LinOp::apply_k) + a fused CGS2
// orthogonalisation (multi-dot / multi-axpy streaming kernels, coefficients never leave the device) + ONE small host
// transfer of the step's scalars; the projected (<= krylovdim x krylovdim) problem is solved on the host, like KrylovKit
// does.  For a row-panel (sharded) effective Hamiltonian the basis vectors are this rank's panels: inner products are
// local partial sums followed by one fused NCCL allreduce of the step's scalars, so every rank sees bit-identical
// Krylov scalars and takes identical decisions.
//   * Lanczos: full re-orthogonalisation, eager convergence test, thick restart (KrylovKit keeps 3/5 of the Ritz vectors)
//   * Arnoldi: Krylov-Schur restarts on the complex Schur form of the projected matrix (Defaults.eigsolver, defaults.jl:33)
//   * GMRES:   restarted, (a0 + a1 f) x = b
#include "internal.h"
#include <cmath>
#include <chrono>
#include <cstdlib>

namespace mpsk {

// ------------------------------------------------------------------------------------------ small dense host algebra
// cyclic Jacobi eigensolver for small real symmetric matrices (projected Lanczos problems)
void jacobi_eigh(int n, std::vector<double>& A, std::vector<double>& evals, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) { diag += A[(size_t)i * n + i] * A[(size_t)i * n + i]; for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j]; }
    if (off <= 1e-32 * (diag + off) || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq; A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk; A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq; V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  evals.resize(n);
  for (int i = 0; i < n; ++i) evals[i] = A[(size_t)i * n + i];
}

// Givens rotation  [ c  s ; -conj(s)  c ] [f; g] = [r; 0]   (c real)
static void givens(cplx f, cplx g, double& c, cplx& s, cplx& r) {
  if (g == cplx(0.0)) { c = 1.0; s = 0.0; r = f; return; }
  if (f == cplx(0.0)) { c = 0.0; s = std::conj(g) / std::abs(g); r = std::abs(g); return; }
  const double af = std::abs(f), ag = std::abs(g), nrm = std::hypot(af, ag);
  const cplx ph = f / af;
  c = af / nrm; s = ph * std::conj(g) / nrm; r = ph * nrm;
}

// Complex Schur decomposition A = Z T Z^H of a general n x n matrix (column-major; overwritten by the upper triangular
// T): Householder reduction to Hessenberg form followed by the implicit single-shift QR algorithm (complex arithmetic
// needs no double shifts).  Used for the projected problems of the Arnoldi / Krylov-Schur eigensolver (n <= krylovdim).
int schur_complex(int n, std::vector<cplx>& H, std::vector<cplx>& Z) {
  auto h = [&](int i, int j) -> cplx& { return H[(size_t)i + (size_t)j * n]; };
  auto z = [&](int i, int j) -> cplx& { return Z[(size_t)i + (size_t)j * n]; };
  Z.assign((size_t)n * n, cplx(0.0));
  for (int i = 0; i < n; ++i) z(i, i) = 1.0;
  // ---- Hessenberg reduction
  std::vector<cplx> v(n);
  for (int k = 0; k + 2 < n; ++k) {
    double xn = 0.0;
    for (int i = k + 1; i < n; ++i) xn += std::norm(h(i, k));
    xn = std::sqrt(xn);
    double tail = 0.0;
    for (int i = k + 2; i < n; ++i) tail += std::norm(h(i, k));
    if (tail == 0.0) continue;
    const cplx x0 = h(k + 1, k);
    const cplx ph = x0 == cplx(0.0) ? cplx(1.0) : x0 / std::abs(x0);
    const cplx alpha = -ph * xn;
    for (int i = k + 1; i < n; ++i) v[i] = h(i, k);
    v[k + 1] -= alpha;
    double vn = 0.0;
    for (int i = k + 1; i < n; ++i) vn += std::norm(v[i]);
    vn = std::sqrt(vn);
    if (vn == 0.0) continue;
    for (int i = k + 1; i < n; ++i) v[i] /= vn;
    for (int j = k; j < n; ++j) {            // H <- (I - 2 v v^H) H
      cplx s = 0.0;
      for (int i = k + 1; i < n; ++i) s += std::conj(v[i]) * h(i, j);
      for (int i = k + 1; i < n; ++i) h(i, j) -= 2.0 * v[i] * s;
    }
    for (int i = 0; i < n; ++i) {            // H <- H (I - 2 v v^H),  Z <- Z (I - 2 v v^H)
      cplx s = 0.0, sz = 0.0;
      for (int j = k + 1; j < n; ++j) { s += h(i, j) * v[j]; sz += z(i, j) * v[j]; }
      for (int j = k + 1; j < n; ++j) { h(i, j) -= 2.0 * s * std::conj(v[j]); z(i, j) -= 2.0 * sz * std::conj(v[j]); }
    }
    for (int i = k + 2; i < n; ++i) h(i, k) = 0.0;
  }
  // ---- shifted QR iterations
  const double eps = 2.220446049250313e-16;
  double anorm = 0.0;
  for (size_t i = 0; i < H.size(); ++i) anorm = std::max(anorm, std::abs(H[i]));
  int hi = n - 1, iter = 0, total = 0;
  while (hi > 0) {
    int l = hi;
    while (l > 0) {
      double s = std::abs(h(l - 1, l - 1)) + std::abs(h(l, l));
      if (s == 0.0) s = anorm;
      if (std::abs(h(l, l - 1)) <= eps * s) { h(l, l - 1) = 0.0; break; }
      --l;
    }
    if (l == hi) { --hi; iter = 0; continue; }
    if (++total > 60 * n + 200) return MPSK_EINTERNAL;
    cplx mu;
    if (iter == 10 || iter == 20) {
      mu = h(hi, hi) + cplx(std::abs(h(hi, hi - 1)) + (hi >= 2 ? std::abs(h(hi - 1, hi - 2)) : 0.0));   // exceptional shift
    } else {
      const cplx a = h(hi - 1, hi - 1), b = h(hi - 1, hi), c = h(hi, hi - 1), d = h(hi, hi);
      const cplx tr = 0.5 * (a + d), det = a * d - b * c;
      const cplx disc = std::sqrt(tr * tr - det);
      const cplx m1 = tr + disc, m2 = tr - disc;
      mu = std::abs(m1 - d) < std::abs(m2 - d) ? m1 : m2;
    }
    ++iter;
    for (int k = l; k < hi; ++k) {
      cplx f, g;
      if (k == l) { f = h(k, k) - mu; g = h(k + 1, k); }
      else { f = h(k, k - 1); g = h(k + 1, k - 1); }
      double c; cplx s, r;
      givens(f, g, c, s, r);
      if (k > l) { h(k, k - 1) = r; h(k + 1, k - 1) = 0.0; }
      for (int j = k; j < n; ++j) {
        const cplx t1 = h(k, j), t2 = h(k + 1, j);
        h(k, j) = c * t1 + s * t2; h(k + 1, j) = -std::conj(s) * t1 + c * t2;
      }
      const int imax = std::min(k + 2, hi);
      for (int i = 0; i <= imax; ++i) {
        const cplx t1 = h(i, k), t2 = h(i, k + 1);
        h(i, k) = c * t1 + std::conj(s) * t2; h(i, k + 1) = -s * t1 + c * t2;
      }
      for (int i = 0; i < n; ++i) {
        const cplx t1 = z(i, k), t2 = z(i, k + 1);
        z(i, k) = c * t1 + std::conj(s) * t2; z(i, k + 1) = -s * t1 + c * t2;
      }
    }
  }
  for (int j = 0; j < n; ++j) for (int i = j + 1; i < n; ++i) h(i, j) = 0.0;
  return MPSK_OK;
}

// move the eigenvalue at diagonal position `from` of the Schur form to position `to` (to <= from) by adjacent swaps
// (one Givens rotation per swap, as LAPACK ztrexc)
void schur_move(int n, std::vector<cplx>& T, std::vector<cplx>& Z, int from, int to) {
  auto t = [&](int i, int j) -> cplx& { return T[(size_t)i + (size_t)j * n]; };
  auto z = [&](int i, int j) -> cplx& { return Z[(size_t)i + (size_t)j * n]; };
  for (int k = from - 1; k >= to; --k) {
    const cplx t11 = t(k, k), t22 = t(k + 1, k + 1);
    double c; cplx s, r;
    givens(t(k, k + 1), t22 - t11, c, s, r);
    for (int j = k + 2; j < n; ++j) {
      const cplx a = t(k, j), b = t(k + 1, j);
      t(k, j) = c * a + s * b; t(k + 1, j) = c * b - std::conj(s) * a;
    }
    for (int i = 0; i < k; ++i) {
      const cplx a = t(i, k), b = t(i, k + 1);
      t(i, k) = c * a + std::conj(s) * b; t(i, k + 1) = c * b - s * a;
    }
    t(k, k) = t22; t(k + 1, k + 1) = t11;
    for (int i = 0; i < n; ++i) {
      const cplx a = z(i, k), b = z(i, k + 1);
      z(i, k) = c * a + std::conj(s) * b; z(i, k + 1) = c * b - s * a;
    }
  }
}

static bool more_wanted(cplx a, cplx b, int which) {
  if (which == MPSK_SR) return a.real() < b.real();
  if (which == MPSK_LR) return a.real() > b.real();
  return std::abs(a) > std::abs(b);
}

// ------------------------------------------------------------------------------------------ Krylov space on the device
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct KSpace {
  Ctx* ctx = nullptr; LinOp* op = nullptr; bool cx = true; int kd = 0; int64_t n = 0;
  std::vector<Tensor*> V;       // kd + 1 basis vectors (structure of op->local_like()): views into one leased slab
  std::vector<Tensor*> spare;   // targets of basis rotations (thick restarts), allocated on first use (second slab)
  double* d_s = nullptr;
  void* slab = nullptr; bool slab_leased = false; void* slab2 = nullptr; bool slab2_leased = false; int nspare_cap = 0;
  ~KSpace() {
    for (Tensor* t : V) delete t;
    for (Tensor* t : spare) delete t;
    if (d_s) ctx->free(d_s);
    if (ctx) { ctx->unlease(Ctx::S_KRYLOV, slab, slab_leased); ctx->unlease(Ctx::S_KRYLOV_SPARE, slab2, slab2_leased); }
  }
  // a vector with the structure of `like` whose storage is `data` (not owned)
  Tensor* view(void* data) {
    const Tensor* like = op->local_like();
    Tensor* t = new Tensor;
    t->ctx = ctx; t->Vl = like->Vl; t->Vr = like->Vr; t->mch = like->mch; t->dtype = like->dtype; t->form = F_KL;
    tensor_layout_only(t);
    t->d = data; t->owns = false;
    return t;
  }

  double* d_h1() const { return d_s; }
  double* d_h2() const { return d_s + 2 * (kd + 1); }
  double* d_n() const { return d_s + 4 * (kd + 1); }
  double* d_c() const { return d_s + 4 * (kd + 1) + 4; }

  int newvec(Tensor** out) {
    const Tensor* like = op->local_like();
    return tensor_create(ctx, like->Vl, like->mch, like->Vr, like->dtype, F_KL, false, out);
  }
  int init(Ctx* c, LinOp* o, int kd_) {
    ctx = c; op = o; kd = kd_; cx = o->local_like()->dtype == MPSK_C128; n = o->local_like()->nelem;
    const size_t vb = o->local_like()->bytes();
    slab = ctx->lease(Ctx::S_KRYLOV, vb * (size_t)(kd + 1), &slab_leased);
    if (!slab) return MPSK_ENOMEM;
    for (int i = 0; i <= kd; ++i) V.push_back(view((char*)slab + (size_t)i * vb));
    d_s = (double*)ctx->alloc(sizeof(double) * (4 * (kd + 1) + 4 + 2 * (kd + 2)));
    if (!d_s) return MPSK_ENOMEM;
    if (!ctx->h_scratch) CUDA_TRY(ctx, cudaMallocHost(&ctx->h_scratch, sizeof(double) * 2048));
    if (ctx->profile < 0) ctx->profile = getenv("MPSK_PROFILE") ? 1 : 0;
    return MPSK_OK;
  }
  int reduce(double* d, int count) { return op->sharded() ? comm_allreduce_sum(ctx, d, (size_t)count) : MPSK_OK; }

  // V[k+1] <- A V[k], orthogonalised against V[0..k] (CGS2); h[j] = <V[j], A V[k]> (both passes summed), beta = norm of
  // the remainder.  The new vector is left un-normalised (its squared norm stays in d_n() for scale_last()).
  int expand(int k, std::vector<cplx>* h, double* beta, bool apply = true) {
    double t0 = 0, t1 = 0, t2 = 0;
    if (ctx->profile) { cudaStreamSynchronize(ctx->stream); t0 = now_s(); }
    if (apply) MPSK_TRY(op->apply_k(V[k], V[k + 1]));
    if (ctx->profile) { cudaStreamSynchronize(ctx->stream); t1 = now_s(); }
    std::vector<const void*> ptrs(k + 1);
    for (int j = 0; j <= k; ++j) ptrs[j] = V[j]->d;
    void* wv = V[k + 1]->d;
    MPSK_TRY(vec_multidot(ctx, cx, k + 1, ptrs.data(), wv, n, d_h1()));
    MPSK_TRY(reduce(d_h1(), 2 * (k + 1)));
    MPSK_TRY(vec_multiaxpy(ctx, cx, k + 1, ptrs.data(), d_h1(), -1.0, wv, n));
    MPSK_TRY(vec_multidot(ctx, cx, k + 1, ptrs.data(), wv, n, d_h2()));
    MPSK_TRY(reduce(d_h2(), 2 * (k + 1)));
    MPSK_TRY(vec_multiaxpy(ctx, cx, k + 1, ptrs.data(), d_h2(), -1.0, wv, n));
    MPSK_TRY(vec_norm2(ctx, cx, wv, n, d_n()));
    MPSK_TRY(reduce(d_n(), 2));
    if (ctx->profile) { cudaStreamSynchronize(ctx->stream); t2 = now_s(); }
    double* hs = ctx->h_scratch;
    CUDA_TRY(ctx, cudaMemcpyAsync(hs, d_s, sizeof(double) * (4 * (kd + 1) + 2), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    h->resize(k + 1);
    for (int j = 0; j <= k; ++j) (*h)[j] = cplx(hs[2 * j] + hs[2 * (kd + 1) + 2 * j], hs[2 * j + 1] + hs[2 * (kd + 1) + 2 * j + 1]);
    const double n2 = hs[4 * (kd + 1)];
    *beta = n2 > 0 ? std::sqrt(n2) : 0.0;
    if (ctx->profile) {
      const double t3 = now_s(); ctx->t_apply += t1 - t0; ctx->t_orth += t2 - t1; ctx->t_sync += t3 - t2;
      if (k % 10 == 9) fprintf(stderr, "[mpsk profile] krylov k=%d cumulative: apply %.3f s, orthogonalise %.3f s, scalars %.3f s\n", k, ctx->t_apply, ctx->t_orth, ctx->t_sync);
    }
    return MPSK_OK;
  }
  // V[idx] /= sqrt(*d_n())  (the squared norm left by expand / norm2)
  int scale_last(int idx) { return vec_scale_rsqrt(ctx, cx, V[idx]->d, n, d_n()); }

  int norm(const Tensor* t, double* out) {
    MPSK_TRY(vec_norm2(ctx, cx, t->d, n, d_n()));
    MPSK_TRY(reduce(d_n(), 2));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scratch, d_n(), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = std::sqrt(std::max(ctx->h_scratch[0], 0.0));
    return MPSK_OK;
  }
  // dst <- src / |src| ; returns |src| if requested
  int normalise_into(const Tensor* src, Tensor* dst, double* nrm_out) {
    if (dst->d != src->d) CUDA_TRY(ctx, cudaMemcpyAsync(dst->d, src->d, src->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
    MPSK_TRY(vec_norm2(ctx, cx, dst->d, n, d_n()));
    MPSK_TRY(reduce(d_n(), 2));
    if (nrm_out) {
      CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scratch, d_n(), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      *nrm_out = std::sqrt(std::max(ctx->h_scratch[0], 0.0));
      if (*nrm_out == 0.0) return MPSK_OK;
    }
    return vec_scale_rsqrt(ctx, cx, dst->d, n, d_n());
  }
  // out (+)= sum_{j<m} coef[j] V[j]   (coefficients from the host)
  int combine(int m, const cplx* coef, Tensor* out, bool accumulate = false) {
    for (int j0 = 0; j0 < m; j0 += kd + 2) {      // coefficient scratch holds kd + 2 entries
      const int cnt = std::min(kd + 2, m - j0);
      double* hp = (double*)ctx->arena_take(sizeof(double) * 2 * cnt);
      if (!hp) return MPSK_ENOMEM;
      for (int j = 0; j < cnt; ++j) { hp[2 * j] = coef[j0 + j].real(); hp[2 * j + 1] = coef[j0 + j].imag(); }
      CUDA_TRY(ctx, cudaMemcpyAsync(d_c(), hp, sizeof(double) * 2 * cnt, cudaMemcpyHostToDevice, ctx->stream));
      if (!accumulate && j0 == 0) CUDA_TRY(ctx, cudaMemsetAsync(out->d, 0, out->bytes(), ctx->stream));
      std::vector<const void*> ptrs(cnt);
      for (int j = 0; j < cnt; ++j) ptrs[j] = V[j0 + j]->d;
      MPSK_TRY(vec_multiaxpy(ctx, cx, cnt, ptrs.data(), d_c(), 1.0, out->d, out->nelem));
    }
    return MPSK_OK;
  }
  // basis rotation of a restart: V[j] <- sum_{i<m} V[i] Y(i, j) for j < keep (Y column-major m x *, leading dim m),
  // V[keep] <- V[m].  New vectors are written to spare buffers and swapped in (no copies of basis vectors).
  int rotate(int m, int keep, const cplx* Y) {
    if (!slab2) {      // spare vectors for every possible restart of this solve (keep <= kd)
      const size_t vb = op->local_like()->bytes();
      nspare_cap = (3 * kd) / 5 + 1;          // restarts keep at most 3/5 of the Ritz / Schur vectors
      slab2 = ctx->lease(Ctx::S_KRYLOV_SPARE, vb * (size_t)nspare_cap, &slab2_leased);
      if (!slab2) return MPSK_ENOMEM;
    }
    if (keep > nspare_cap) return ctx->fail(MPSK_EINTERNAL, "krylov: restart keeps more vectors than reserved");
    while ((int)spare.size() < keep) spare.push_back(view((char*)slab2 + spare.size() * op->local_like()->bytes()));
    for (int j = 0; j < keep; ++j) MPSK_TRY(combine(m, Y + (size_t)j * m, spare[j]));
    for (int j = 0; j < keep; ++j) std::swap(V[j], spare[j]);
    std::swap(V[keep], V[m]);
    return MPSK_OK;
  }
};

static int check_vectors(Ctx* ctx, LinOp* op, const Tensor* a, const Tensor* b, int krylovdim) {
  if (krylovdim < 1 || krylovdim > 120) return ctx->fail(MPSK_EINVAL, "krylovdim out of range");
  if (!a->same_structure(*op->like()) || !b->same_structure(*op->like())) return ctx->fail(MPSK_ESPACE, "vector does not match the operator");
  if (a->form != F_KL || b->form != F_KL) return ctx->fail(MPSK_EINTERNAL, "Krylov vectors must be in KL form");
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ Lanczos eigensolver
int eigsolve_lanczos(Ctx* ctx, LinOp* op, const Tensor* x0, int which, int krylovdim, int maxiter, double tol,
                     cplx* val, Tensor* vec, mpsk_info* info) {
  MPSK_TRY(check_vectors(ctx, op, x0, vec, krylovdim));
  const int kd = krylovdim;
  KSpace K;
  MPSK_TRY(K.init(ctx, op, kd));
  MPSK_TRY(op->to_local(x0, K.V[0]));
  double nrm0 = 0;
  MPSK_TRY(K.normalise_into(K.V[0], K.V[0], &nrm0));
  if (nrm0 == 0.0) return ctx->fail(MPSK_EINVAL, "eigsolve: zero starting vector");
  const int ld = kd + 1;
  std::vector<double> T((size_t)ld * ld, 0.0);          // projected matrix (real symmetric), row-major ld x ld
  int numops = 0, m0 = 0, it = 0, m = 0; double theta = 0, res = 0; bool conv = false;
  std::vector<double> ev, U; std::vector<cplx> h;
  int best = 0;
  for (it = 1; it <= std::max(maxiter, 1); ++it) {
    double beta = 0;
    for (int k = m0; k < kd; ++k) {
      MPSK_TRY(K.expand(k, &h, &beta));
      ++numops; m = k + 1;
      T[(size_t)k * ld + k] = h[k].real();
      T[(size_t)k * ld + k + 1] = T[(size_t)(k + 1) * ld + k] = beta;
      std::vector<double> Tm((size_t)m * m);
      for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) Tm[(size_t)i * m + j] = T[(size_t)i * ld + j];
      jacobi_eigh(m, Tm, ev, U);
      best = 0;
      for (int i = 1; i < m; ++i) if (more_wanted(ev[i], ev[best], which)) best = i;
      theta = ev[best];
      res = std::fabs(beta * U[(size_t)(m - 1) * m + best]);
      if (res < tol || beta < 1e-14) { conv = true; break; }
      if (k + 1 < kd || it < maxiter) MPSK_TRY(K.scale_last(k + 1));
    }
    if (conv || it >= maxiter) break;
    // ---- thick restart: keep the `keep` most wanted Ritz vectors + the residual direction
    int keep = std::min(m - 1, std::max(1, (3 * kd) / 5));
    if (keep < 1) { m0 = 0; std::swap(K.V[0], K.V[m]); std::fill(T.begin(), T.end(), 0.0); continue; }   // krylovdim 1: plain restart
    std::vector<int> order(m);
    for (int i = 0; i < m; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return more_wanted(ev[a], ev[b], which); });
    std::vector<cplx> Y((size_t)m * keep);
    for (int j = 0; j < keep; ++j) for (int i = 0; i < m; ++i) Y[(size_t)j * m + i] = U[(size_t)i * m + order[j]];
    MPSK_TRY(K.rotate(m, keep, Y.data()));
    std::fill(T.begin(), T.end(), 0.0);
    for (int j = 0; j < keep; ++j) {
      T[(size_t)j * ld + j] = ev[order[j]];
      T[(size_t)j * ld + keep] = T[(size_t)keep * ld + j] = beta * U[(size_t)(m - 1) * m + order[j]];
    }
    m0 = keep;
  }
  Tensor* xr = nullptr;
  MPSK_TRY(K.newvec(&xr));
  std::unique_ptr<Tensor> xr_guard(xr);
  std::vector<cplx> c(m);
  for (int i = 0; i < m; ++i) c[i] = U[(size_t)i * m + best];
  MPSK_TRY(K.combine(m, c.data(), xr));
  MPSK_TRY(K.normalise_into(xr, xr, nullptr));
  MPSK_TRY(op->from_local(xr, vec));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *val = theta;
  if (info) { info->converged = conv ? 1 : 0; info->numiter = std::min(it, std::max(maxiter, 1)); info->numops = numops; info->normres = res; }
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ Arnoldi / Krylov-Schur
// howmany > 1 (KrylovKit eigsolve(f, x0, howmany, which): transfer_spectrum, toolbox.jl:42-58): the leading `howmany`
// positions of the Schur form hold the most wanted Ritz values in order of wantedness; the solve has converged when the
// residuals |b^H z_j| of all of them are below tol.  `vals` receives *nvals = min(howmany, m) values, `nconv` how many leading ones
// have converged; `vec` (first Schur vector = eigenvector of the most wanted value) may be NULL.
int eigsolve_arnoldi(Ctx* ctx, LinOp* op, const Tensor* x0, int which, int krylovdim, int maxiter, double tol,
                     cplx* val, Tensor* vec, mpsk_info* info, int howmany, cplx* vals, int* nvals, int* nconv) {
  MPSK_TRY(check_vectors(ctx, op, x0, vec ? vec : x0, krylovdim));
  const int kd = krylovdim;
  if (howmany < 1 || howmany > std::max(1, (3 * kd) / 5) || (howmany > 1 && howmany > kd - 1))
    return ctx->fail(MPSK_EINVAL, "eigsolve: howmany must be between 1 and 3/5 of krylovdim (the restart keeps that many Schur vectors)");
  KSpace K;
  MPSK_TRY(K.init(ctx, op, kd));
  MPSK_TRY(op->to_local(x0, K.V[0]));
  double nrm0 = 0;
  MPSK_TRY(K.normalise_into(K.V[0], K.V[0], &nrm0));
  if (nrm0 == 0.0) return ctx->fail(MPSK_EINVAL, "eigsolve: zero starting vector");
  const int ld = kd + 1;
  std::vector<cplx> B((size_t)ld * ld, cplx(0.0));      // Krylov decomposition A V_m = V_m B_m + v_{m+1} b^H, column-major
  auto b = [&](int i, int j) -> cplx& { return B[(size_t)i + (size_t)j * ld]; };
  int numops = 0, m0 = 0, it = 0, m = 0, want = 1, nc = 0; double res = 0; bool conv = false; cplx theta = 0;
  std::vector<cplx> S, Z, h;
  for (it = 1; it <= std::max(maxiter, 1); ++it) {
    double beta = 0;
    for (int k = m0; k < kd; ++k) {
      MPSK_TRY(K.expand(k, &h, &beta));
      ++numops; m = k + 1;
      for (int i = 0; i <= k; ++i) b(i, k) = h[i];
      b(k + 1, k) = beta;
      S.assign((size_t)m * m, cplx(0.0));
      for (int j = 0; j < m; ++j) for (int i = 0; i < m; ++i) S[(size_t)i + (size_t)j * m] = b(i, j);
      if (schur_complex(m, S, Z) != MPSK_OK) return ctx->fail(MPSK_EINTERNAL, "Schur decomposition of the projected matrix did not converge");
      want = std::min(howmany, m);
      for (int j = 0; j < want; ++j) {                       // selection sort of the leading block by wantedness
        int p = j;
        for (int i = j + 1; i < m; ++i) if (more_wanted(S[(size_t)i + (size_t)i * m], S[(size_t)p + (size_t)p * m], which)) p = i;
        schur_move(m, S, Z, p, j);
      }
      theta = S[0];
      res = 0.0; nc = 0;
      for (int j = 0; j < want; ++j) {                       // |b^H z_j| with b = beta e_m
        const double rj = std::abs(beta * Z[(size_t)(m - 1) + (size_t)j * m]);
        res = std::max(res, rj);
        if (rj < tol && nc == j) ++nc;
      }
      if ((m >= howmany && nc >= howmany) || beta < 1e-14) { conv = true; if (beta < 1e-14) nc = want; break; }
      if (k + 1 < kd || it < maxiter) MPSK_TRY(K.scale_last(k + 1));
    }
    if (conv || it >= maxiter) break;
    // ---- Krylov-Schur restart: order the Schur form by wantedness, keep the leading block
    int keep = std::min(m - 1, std::max(1, (3 * kd) / 5));
    if (keep < 1) { m0 = 0; std::swap(K.V[0], K.V[m]); std::fill(B.begin(), B.end(), cplx(0.0)); continue; }
    for (int j = want; j < keep; ++j) {
      int p = j;
      for (int i = j + 1; i < m; ++i) if (more_wanted(S[(size_t)i + (size_t)i * m], S[(size_t)p + (size_t)p * m], which)) p = i;
      schur_move(m, S, Z, p, j);
    }
    MPSK_TRY(K.rotate(m, keep, Z.data()));
    std::fill(B.begin(), B.end(), cplx(0.0));
    for (int j = 0; j < keep; ++j) {
      for (int i = 0; i <= j; ++i) b(i, j) = S[(size_t)i + (size_t)j * m];
      b(keep, j) = beta * Z[(size_t)(m - 1) + (size_t)j * m];
    }
    m0 = keep;
  }
  if (vec) {
    Tensor* xr = nullptr;
    MPSK_TRY(K.newvec(&xr));
    std::unique_ptr<Tensor> xr_guard(xr);
    MPSK_TRY(K.combine(m, Z.data(), xr));                   // first Schur vector = eigenvector of the wanted Ritz value
    MPSK_TRY(K.normalise_into(xr, xr, nullptr));
    MPSK_TRY(op->from_local(xr, vec));
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (val) *val = theta;
  if (vals) for (int j = 0; j < std::min(howmany, m); ++j) vals[j] = S[(size_t)j + (size_t)j * m];
  if (nvals) *nvals = std::min(howmany, m);
  if (nconv) *nconv = std::min(nc, std::min(howmany, m));
  if (info) { info->converged = conv ? 1 : 0; info->numiter = std::min(it, std::max(maxiter, 1)); info->numops = numops; info->normres = res; }
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ GMRES
// (a0 + a1 f) x = b, restarted GMRES(krylovdim) with Givens rotations on the projected Hessenberg matrix
int linsolve_gmres(Ctx* ctx, LinOp* op, const Tensor* bvec, const Tensor* x0, cplx a0, cplx a1, int krylovdim, int maxiter,
                   double tol, Tensor* xout, mpsk_info* info) {
  MPSK_TRY(check_vectors(ctx, op, bvec, x0, krylovdim));
  if (!xout->same_structure(*op->like()) || xout->form != F_KL) return ctx->fail(MPSK_ESPACE, "solution vector does not match the operator");
  if (op->sharded()) return ctx->fail(MPSK_EINVAL, "linsolve on a row-panel operator is not supported");
  const int kd = krylovdim;
  KSpace K;
  MPSK_TRY(K.init(ctx, op, kd));
  const bool cx = K.cx;
  if (!cx && (a0.imag() != 0.0 || a1.imag() != 0.0)) return ctx->fail(MPSK_EINVAL, "complex coefficients need a ComplexF64 vector");
  Tensor *x = nullptr, *bl = nullptr;
  MPSK_TRY(K.newvec(&x)); std::unique_ptr<Tensor> xg(x);
  MPSK_TRY(K.newvec(&bl)); std::unique_ptr<Tensor> bg(bl);
  MPSK_TRY(op->to_local(x0, x));
  MPSK_TRY(op->to_local(bvec, bl));
  // w <- a0 v + a1 f(v)
  auto shifted = [&](const Tensor* v, Tensor* w) -> int {
    MPSK_TRY(op->apply_k(v, w));
    if (a1 != cplx(1.0)) MPSK_TRY(vec_scale(ctx, cx, w->d, w->nelem, a1));
    if (a0 != cplx(0.0)) MPSK_TRY(vec_axpy(ctx, cx, a0, v->d, w->d, w->nelem));
    return MPSK_OK;
  };
  int numops = 0, it = 0; double beta = 0; bool conv = false;
  std::vector<cplx> Hc((size_t)(kd + 1) * kd), g(kd + 1), sn(kd), h; std::vector<double> cs(kd);
  for (it = 0; it <= maxiter; ++it) {
    // true residual r = b - (a0 + a1 f) x  -> V[0]
    MPSK_TRY(shifted(x, K.V[0])); ++numops;
    MPSK_TRY(vec_scale(ctx, cx, K.V[0]->d, K.n, cplx(-1.0)));
    MPSK_TRY(vec_axpy(ctx, cx, cplx(1.0), bl->d, K.V[0]->d, K.n));
    MPSK_TRY(K.normalise_into(K.V[0], K.V[0], &beta));
    if (beta < tol) { conv = true; break; }
    if (it == maxiter) break;
    std::fill(g.begin(), g.end(), cplx(0.0)); g[0] = beta;
    int m = 0; double resk = beta;
    for (int k = 0; k < kd; ++k) {
      MPSK_TRY(shifted(K.V[k], K.V[k + 1])); ++numops;
      double hn = 0;
      MPSK_TRY(K.expand(k, &h, &hn, /*apply=*/false));
      auto Hk = [&](int i) -> cplx& { return Hc[(size_t)i + (size_t)k * (kd + 1)]; };
      for (int i = 0; i <= k; ++i) Hk(i) = h[i];
      Hk(k + 1) = hn;
      for (int i = 0; i < k; ++i) {             // previous rotations
        const cplx t1 = Hk(i), t2 = Hk(i + 1);
        Hk(i) = cs[i] * t1 + sn[i] * t2; Hk(i + 1) = -std::conj(sn[i]) * t1 + cs[i] * t2;
      }
      cplx r;
      givens(Hk(k), Hk(k + 1), cs[k], sn[k], r);
      Hk(k) = r; Hk(k + 1) = 0.0;
      const cplx g1 = g[k];
      g[k] = cs[k] * g1; g[k + 1] = -std::conj(sn[k]) * g1;
      resk = std::abs(g[k + 1]);
      m = k + 1;
      if (resk < tol || hn < 1e-14) break;
      if (k + 1 < kd) MPSK_TRY(K.scale_last(k + 1));
    }
    // back substitution R y = g[0:m], x += V_m y
    std::vector<cplx> y(m);
    for (int i = m - 1; i >= 0; --i) {
      cplx s = g[i];
      for (int j = i + 1; j < m; ++j) s -= Hc[(size_t)i + (size_t)j * (kd + 1)] * y[j];
      y[i] = s / Hc[(size_t)i + (size_t)i * (kd + 1)];
    }
    MPSK_TRY(K.combine(m, y.data(), x, /*accumulate=*/true));
  }
  MPSK_TRY(op->from_local(x, xout));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (info) { info->converged = conv ? 1 : 0; info->numiter = it; info->numops = numops; info->normres = beta; }
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ exponentiate
// y = exp(tau H) y0, H Hermitian.  Lanczos with sub-stepping until the a-posteriori estimate
// |beta_m [exp(tau dt T)]_{m,1}| is below tol (same rule as oracle/krylov.py:exponentiate_lanczos).
int expv_lanczos(Ctx* ctx, LinOp* op, cplx tau, const Tensor* y0, int krylovdim, int maxiter, double tol, Tensor* yout,
                 mpsk_info* info) {
  MPSK_TRY(check_vectors(ctx, op, y0, yout, krylovdim));
  const int kd = krylovdim;
  KSpace K;
  MPSK_TRY(K.init(ctx, op, kd));
  const bool cx = K.cx;
  if (!cx && tau.imag() != 0.0) return ctx->fail(MPSK_EINVAL, "complex time step needs a ComplexF64 vector");
  Tensor* ycur = nullptr;
  MPSK_TRY(K.newvec(&ycur));
  std::unique_ptr<Tensor> yguard(ycur);
  MPSK_TRY(op->to_local(y0, ycur));
  double t_done = 0.0; int numops = 0; bool conv = false; int it = 0;
  std::vector<double> alpha(kd), beta(kd); std::vector<cplx> h;
  for (it = 0; it < maxiter; ++it) {
    double nrm = 0;
    MPSK_TRY(K.normalise_into(ycur, K.V[0], &nrm));
    if (nrm == 0.0) { conv = true; break; }
    int m = 0; bool full = false;
    for (int k = 0; k < kd; ++k) {
      MPSK_TRY(K.expand(k, &h, &beta[k]));
      alpha[k] = h[k].real();
      ++numops; m = k + 1;
      if (beta[k] < 1e-14) break;
      if (k + 1 < kd) MPSK_TRY(K.scale_last(k + 1));
      else full = true;
    }
    std::vector<double> Tm((size_t)m * m, 0.0), ev, U;
    for (int i = 0; i < m; ++i) { Tm[(size_t)i * m + i] = alpha[i]; if (i + 1 < m) Tm[(size_t)i * m + i + 1] = Tm[(size_t)(i + 1) * m + i] = beta[i]; }
    jacobi_eigh(m, Tm, ev, U);
    double dt = 1.0 - t_done;
    std::vector<cplx> E(m);
    while (true) {
      for (int i = 0; i < m; ++i) {
        cplx s = 0;
        for (int j = 0; j < m; ++j) s += U[(size_t)i * m + j] * std::exp(tau * dt * ev[j]) * U[(size_t)0 * m + j];
        E[i] = s;
      }
      const double err = full ? std::fabs(beta[m - 1]) * std::abs(E[m - 1]) * std::abs(tau * dt) : 0.0;
      if (err <= tol * std::max(dt, 1e-3) || dt < 1e-12) break;
      dt *= 0.5;
    }
    for (int i = 0; i < m; ++i) E[i] *= nrm;
    MPSK_TRY(K.combine(m, E.data(), ycur));
    t_done += dt;
    if (t_done >= 1.0 - 1e-15) { conv = true; ++it; break; }
  }
  MPSK_TRY(op->from_local(ycur, yout));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (info) { info->converged = conv ? 1 : 0; info->numiter = it; info->numops = numops; info->normres = 0.0; }
  return MPSK_OK;
}

}  // namespace mpsk
