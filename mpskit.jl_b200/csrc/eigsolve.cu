// Device-resident Krylov solvers behind fixedpoint (src/algorithms/fixedpoint.jl:9-30 -> KrylovKit eigsolve)
// and integrate (src/algorithms/timestep/integrators.jl:20-25 -> KrylovKit exponentiate).
// The Krylov basis lives in HBM; every step is one matvec (heff_apply) + a fused CGS2 orthogonalisation
// (multi-dot / multi-axpy streaming kernels) + one small host transfer of (alpha, beta); the projected
// (<= krylovdim x krylovdim) problem is solved on the host, like KrylovKit does.
#include "internal.h"
#include <cmath>
#include <chrono>
#include <cstdlib>

namespace mpsk {

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

struct KrylovWork {
  Ctx* ctx; std::vector<Tensor*> V; double* d_s = nullptr; int kd;
  ~KrylovWork() { for (Tensor* t : V) delete t; if (d_s) ctx->free(d_s); }
};

static int work_init(Ctx* ctx, const Tensor* like, int kd, KrylovWork* w) {
  w->ctx = ctx; w->kd = kd;
  for (int i = 0; i <= kd; ++i) {
    Tensor* t = nullptr;
    MPSK_TRY(tensor_create(ctx, like->Vl, like->mch, like->Vr, like->dtype, F_KL, false, &t));
    w->V.push_back(t);
  }
  w->d_s = (double*)ctx->alloc(sizeof(double) * (4 * (kd + 1) + 4 + 2 * (kd + 1)));
  if (!w->d_s) return MPSK_ENOMEM;
  if (!ctx->h_scratch) CUDA_TRY(ctx, cudaMallocHost(&ctx->h_scratch, sizeof(double) * 1024));
  return MPSK_OK;
}

// One Lanczos expansion step: V[k+1] <- orthonormalised H V[k]; returns alpha_k and beta_k (host).
// If normalise is false the new vector is left un-normalised (caller stops here).
static double g_t_apply = 0, g_t_orth = 0, g_t_sync = 0; static int g_prof = -1;
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int lanczos_step(Ctx* ctx, HeffOp* op, KrylovWork& w, int k, double* alpha, double* beta, bool cx) {
  if (g_prof < 0) g_prof = getenv("MPSK_PROFILE") ? 1 : 0;
  double t0 = 0, t1 = 0, t2 = 0;
  if (g_prof) { cudaStreamSynchronize(ctx->stream); t0 = now_s(); }
  const int64_t n = w.V[0]->nelem;
  const int kd = w.kd;
  double* d_h1 = w.d_s; double* d_h2 = w.d_s + 2 * (kd + 1); double* d_n = w.d_s + 4 * (kd + 1);
  MPSK_TRY(heff_apply(ctx, op, w.V[k], w.V[k + 1]));
  if (g_prof) { cudaStreamSynchronize(ctx->stream); t1 = now_s(); }
  std::vector<const void*> ptrs(k + 1);
  for (int j = 0; j <= k; ++j) ptrs[j] = w.V[j]->d;
  void* wv = w.V[k + 1]->d;
  MPSK_TRY(vec_multidot(ctx, cx, k + 1, ptrs.data(), wv, n, d_h1));
  MPSK_TRY(vec_multiaxpy(ctx, cx, k + 1, ptrs.data(), d_h1, -1.0, wv, n));
  MPSK_TRY(vec_multidot(ctx, cx, k + 1, ptrs.data(), wv, n, d_h2));
  MPSK_TRY(vec_multiaxpy(ctx, cx, k + 1, ptrs.data(), d_h2, -1.0, wv, n));
  MPSK_TRY(vec_norm2(ctx, cx, wv, n, d_n));
  if (g_prof) { cudaStreamSynchronize(ctx->stream); t2 = now_s(); }
  double* h = ctx->h_scratch;
  CUDA_TRY(ctx, cudaMemcpyAsync(h, w.d_s, sizeof(double) * (4 * (kd + 1) + 2), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *alpha = h[2 * k] + h[2 * (kd + 1) + 2 * k];
  const double n2 = h[4 * (kd + 1)];
  *beta = n2 > 0 ? std::sqrt(n2) : 0.0;
  if (g_prof) { const double t3 = now_s(); g_t_apply += t1 - t0; g_t_orth += t2 - t1; g_t_sync += t3 - t2;
    if (k % 10 == 9) fprintf(stderr, "[mpsk profile] lanczos k=%d cumulative: apply %.3f s, orthogonalise %.3f s, scalars %.3f s\n", k, g_t_apply, g_t_orth, g_t_sync); }
  return MPSK_OK;
}

static int normalise_into(Ctx* ctx, KrylovWork& w, const Tensor* src, Tensor* dst, bool cx, double* nrm_out) {
  double* d_n = w.d_s + 4 * (w.kd + 1);
  if (dst->d != src->d) CUDA_TRY(ctx, cudaMemcpyAsync(dst->d, src->d, src->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  MPSK_TRY(vec_norm2(ctx, cx, dst->d, dst->nelem, d_n));
  if (nrm_out) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scratch, d_n, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *nrm_out = std::sqrt(ctx->h_scratch[0]);
    if (*nrm_out == 0.0) return MPSK_OK;
  }
  MPSK_TRY(vec_scale_rsqrt(ctx, cx, dst->d, dst->nelem, d_n));
  return MPSK_OK;
}

// out = sum_j coef[j] V[j]  (coefficients from the host)
static int combine(Ctx* ctx, KrylovWork& w, int m, const cplx* coef, Tensor* out, bool cx) {
  double* d_c = w.d_s + 4 * (w.kd + 1) + 4;
  double* h = ctx->h_scratch + 512;
  for (int j = 0; j < m; ++j) { h[2 * j] = coef[j].real(); h[2 * j + 1] = coef[j].imag(); }
  CUDA_TRY(ctx, cudaMemcpyAsync(d_c, h, sizeof(double) * 2 * m, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemsetAsync(out->d, 0, out->bytes(), ctx->stream));
  std::vector<const void*> ptrs(m);
  for (int j = 0; j < m; ++j) ptrs[j] = w.V[j]->d;
  MPSK_TRY(vec_multiaxpy(ctx, cx, m, ptrs.data(), d_c, 1.0, out->d, out->nelem));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // h_scratch is reused by the caller
  return MPSK_OK;
}

int eigsolve_lanczos(Ctx* ctx, HeffOp* op, const Tensor* x0, int which, int krylovdim, int maxiter, double tol,
                     cplx* val, Tensor* vec, mpsk_info* info) {
  if (krylovdim < 1 || krylovdim > 120) return ctx->fail(MPSK_EINVAL, "krylovdim out of range");
  if (!x0->same_structure(*op->xs) || !vec->same_structure(*op->xs)) return ctx->fail(MPSK_ESPACE, "x0 does not match the operator");
  const bool cx = x0->dtype == MPSK_C128;
  KrylovWork w;
  MPSK_TRY(work_init(ctx, x0, krylovdim, &w));
  double nrm0 = 0;
  MPSK_TRY(normalise_into(ctx, w, x0, w.V[0], cx, &nrm0));
  if (nrm0 == 0.0) return ctx->fail(MPSK_EINVAL, "eigsolve: zero starting vector");
  std::vector<double> alpha(krylovdim), beta(krylovdim);
  int numops = 0; double theta = 0, res = 0; bool conv = false; int it = 0;
  Tensor* xr = nullptr;
  MPSK_TRY(tensor_create(ctx, x0->Vl, x0->mch, x0->Vr, x0->dtype, F_KL, false, &xr));
  std::unique_ptr<Tensor> xr_guard(xr);
  for (it = 1; it <= maxiter; ++it) {
    int m = 0; std::vector<double> y;
    for (int k = 0; k < krylovdim; ++k) {
      MPSK_TRY(lanczos_step(ctx, op, w, k, &alpha[k], &beta[k], cx));
      ++numops; m = k + 1;
      std::vector<double> Tm((size_t)m * m, 0.0), ev, U;
      for (int i = 0; i < m; ++i) { Tm[(size_t)i * m + i] = alpha[i]; if (i + 1 < m) Tm[(size_t)i * m + i + 1] = Tm[(size_t)(i + 1) * m + i] = beta[i]; }
      jacobi_eigh(m, Tm, ev, U);
      int idx = 0;
      for (int i = 1; i < m; ++i) {
        if (which == MPSK_SR && ev[i] < ev[idx]) idx = i;
        if (which == MPSK_LR && ev[i] > ev[idx]) idx = i;
        if (which == MPSK_LM && std::fabs(ev[i]) > std::fabs(ev[idx])) idx = i;
      }
      theta = ev[idx];
      y.resize(m);
      for (int i = 0; i < m; ++i) y[i] = U[(size_t)i * m + idx];
      res = std::fabs(beta[k] * y[k]);
      if (res < tol || beta[k] < 1e-14) { conv = true; break; }
      if (k + 1 < krylovdim) MPSK_TRY(vec_scale_rsqrt(ctx, cx, w.V[k + 1]->d, w.V[k + 1]->nelem, w.d_s + 4 * (krylovdim + 1)));
    }
    std::vector<cplx> c(m);
    for (int i = 0; i < m; ++i) c[i] = y[i];
    MPSK_TRY(combine(ctx, w, m, c.data(), xr, cx));
    MPSK_TRY(normalise_into(ctx, w, xr, w.V[0], cx, nullptr));
    if (conv) break;
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(vec->d, w.V[0]->d, vec->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *val = theta;
  if (info) { info->converged = conv ? 1 : 0; info->numiter = std::min(it, maxiter); info->numops = numops; info->normres = res; }
  return MPSK_OK;
}

int eigsolve_arnoldi(Ctx* ctx, HeffOp*, const Tensor*, int, int, int, double, cplx*, Tensor*, mpsk_info*) {
  return ctx->fail(MPSK_EINVAL, "non-Hermitian Arnoldi eigsolve is not available in this build (SURVEY.md 8f item 2)");
}

// y = exp(tau H) y0, H Hermitian.  Lanczos with sub-stepping until the a-posteriori estimate
// |beta_m [exp(tau dt T)]_{m,1}| is below tol (same rule as oracle/krylov.py:exponentiate_lanczos).
int expv_lanczos(Ctx* ctx, HeffOp* op, cplx tau, const Tensor* y0, int krylovdim, int maxiter, double tol, Tensor* yout,
                 mpsk_info* info) {
  if (krylovdim < 1 || krylovdim > 120) return ctx->fail(MPSK_EINVAL, "krylovdim out of range");
  if (!y0->same_structure(*op->xs) || !yout->same_structure(*op->xs)) return ctx->fail(MPSK_ESPACE, "y0 does not match the operator");
  const bool cx = y0->dtype == MPSK_C128;
  if (!cx && tau.imag() != 0.0) return ctx->fail(MPSK_EINVAL, "complex time step needs a ComplexF64 vector");
  KrylovWork w;
  MPSK_TRY(work_init(ctx, y0, krylovdim, &w));
  Tensor* ycur = nullptr;
  MPSK_TRY(tensor_create(ctx, y0->Vl, y0->mch, y0->Vr, y0->dtype, F_KL, false, &ycur));
  std::unique_ptr<Tensor> yguard(ycur);
  CUDA_TRY(ctx, cudaMemcpyAsync(ycur->d, y0->d, y0->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  double t_done = 0.0; int numops = 0; bool conv = false; int it = 0;
  std::vector<double> alpha(krylovdim), beta(krylovdim);
  for (it = 0; it < maxiter; ++it) {
    double nrm = 0;
    MPSK_TRY(normalise_into(ctx, w, ycur, w.V[0], cx, &nrm));
    if (nrm == 0.0) { conv = true; break; }
    int m = 0; bool full = false;
    for (int k = 0; k < krylovdim; ++k) {
      MPSK_TRY(lanczos_step(ctx, op, w, k, &alpha[k], &beta[k], cx));
      ++numops; m = k + 1;
      if (beta[k] < 1e-14) break;
      if (k + 1 < krylovdim) MPSK_TRY(vec_scale_rsqrt(ctx, cx, w.V[k + 1]->d, w.V[k + 1]->nelem, w.d_s + 4 * (krylovdim + 1)));
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
    MPSK_TRY(combine(ctx, w, m, E.data(), ycur, cx));
    t_done += dt;
    if (t_done >= 1.0 - 1e-15) { conv = true; ++it; break; }
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(yout->d, ycur->d, yout->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (info) { info->converged = conv ? 1 : 0; info->numiter = it; info->numops = numops; info->normres = 0.0; }
  return MPSK_OK;
}

}  // namespace mpsk
