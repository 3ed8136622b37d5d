// Effective Hamiltonians dC / dAC / dAC2 (derivatives.jl:77-197) as one generic three-launch pipeline
//   stage A : T[(i,a),(m,c)]   = GL_i[a,b] x[b,m,c]          grouped GEMM, MPO levels stacked along M
//   mix     : U_k[a,m',c]      = sum_{i,m} W[i,k][m',m] T_i  memory-bound rectangle gather (scalar and operator
//                                                            MPO entries folded in as coefficients)
//   stage C : y[a,m',c']       = sum_{k,c} U_k[a,m',c] GR_k[c,c']   grouped GEMM, levels stacked along K
// which is the optimal-order contraction whose flop count SURVEY.md 8d defines as the algorithmic work.
// The reference evaluates the same sum entry by entry (derivatives.jl:125-151); results agree to rounding.
#include "internal.h"
#include <chrono>
#include <cstdlib>

namespace mpsk {

HeffOp::~HeffOp() {
  if (ctx) { ctx->unlease(Ctx::S_HEFF_T, T, T_leased); ctx->unlease(Ctx::S_HEFF_U, U, U_leased); ctx->free(gather); }
}

static int build_W(Ctx* ctx, int kind, const MpoSlice* h1, const MpoSlice* h2, const Tensor* x_like, int nlev,
                   std::vector<WEntry>* W) {
  W->clear();
  if (kind == MPSK_OP_DC) {
    if (x_like->nm() != 1) return ctx->fail(MPSK_ESPACE, "dC acts on bond tensors (one label)");
    for (int i = 0; i < nlev; ++i) W->push_back(WEntry{i, i, 0, 0, cplx(1.0)});
    return MPSK_OK;
  }
  if (!h1) {
    // dAC(x, nothing, GL, GR) / dAC2(x, nothing, nothing, GL, GR) (derivatives.jl:112-114,159-161): overlap environments,
    // one level, identity on the physical labels
    if (nlev != 1) return ctx->fail(MPSK_ESPACE, "MPO-free derivative needs single-level (bond) environments");
    for (int mm = 0; mm < x_like->nm(); ++mm) W->push_back(WEntry{0, 0, mm, mm, cplx(1.0)});
    return MPSK_OK;
  }
  if (kind == MPSK_OP_DAC) {
    if (x_like->nm() != h1->d) return ctx->fail(MPSK_ESPACE, "dAC: vector has %d labels, slice has d=%d", x_like->nm(), h1->d);
    for (int i = 0; i < h1->odim; ++i)
      for (int k = 0; k < h1->odim; ++k) {
        if (!h1->kind[i * h1->odim + k]) continue;
        for (int po = 0; po < h1->d; ++po)
          for (int pi = 0; pi < h1->d; ++pi) {
            cplx c = h1->elem(i, k, po, pi);
            if (c != cplx(0.0)) W->push_back(WEntry{i, k, po, pi, c});
          }
      }
    return MPSK_OK;
  }
  if (!h2) return ctx->fail(MPSK_EINVAL, "dAC2 needs two MPO slices");
  if (h1->odim != h2->odim) return ctx->fail(MPSK_ESPACE, "slices with different odim");
  const int d1 = h1->d, d2 = h2->d, od = h1->odim;
  if (x_like->nm() != d1 * d2) return ctx->fail(MPSK_ESPACE, "dAC2: vector has %d labels, expected %d", x_like->nm(), d1 * d2);
  std::map<std::tuple<int, int, int, int>, cplx> acc;
  for (int i = 0; i < od; ++i)
    for (int j = 0; j < od; ++j) {
      if (!h1->kind[i * od + j]) continue;
      for (int k = 0; k < od; ++k) {
        if (!h2->kind[j * od + k]) continue;
        for (int s = 0; s < d1; ++s)
          for (int u = 0; u < d1; ++u) {
            const cplx a = h1->elem(i, j, s, u);
            if (a == cplx(0.0)) continue;
            for (int T = 0; T < d2; ++T)
              for (int t = 0; t < d2; ++t) {
                const cplx b = h2->elem(j, k, T, t);
                if (b == cplx(0.0)) continue;
                acc[std::make_tuple(i, k, s * d2 + T, u * d2 + t)] += a * b;
              }
          }
      }
    }
  for (auto& kv : acc)
    if (kv.second != cplx(0.0))
      W->push_back(WEntry{std::get<0>(kv.first), std::get<1>(kv.first), std::get<2>(kv.first), std::get<3>(kv.first), kv.second});
  return MPSK_OK;
}

int heff_create(Ctx* ctx, int kind, const MpoSlice* h1, const MpoSlice* h2, const Tensor* GL, const Tensor* GR,
                const Tensor* x_like, HeffOp** out) {
  if (!GL || !GR || !x_like) return ctx->fail(MPSK_EINVAL, "null argument");
  if (ctx->profile < 0) ctx->profile = getenv("MPSK_PROFILE") ? 1 : 0;
  auto wall = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tp0 = 0, tp1 = 0, tp2 = 0;
  if (ctx->profile) { cudaStreamSynchronize(ctx->stream); tp0 = wall(); }
  if (GL->dtype != x_like->dtype || GR->dtype != x_like->dtype) return ctx->fail(MPSK_EINVAL, "mixed scalar types");
  // ket legs of the environments must match x; the bra legs (GL->Vl, GR->Vr) fix the structure of y.  Square operators
  // (bra = ket) are what the eigensolvers iterate; a row panel of the bra-left space is the multi-GPU shard; general
  // rectangular operators project a vector of one state onto the tangent space of another (ac_proj / ac2_proj of
  // `approximate`, derivatives.jl:209-232).
  if (!(GL->Vr == x_like->Vl)) return ctx->fail(MPSK_ESPACE, "left environment does not match the left bond of x");
  if (!(GR->Vl == x_like->Vr)) return ctx->fail(MPSK_ESPACE, "right environment does not match the right bond of x");
  if (GL->Vl.mod != x_like->Vl.mod || GR->Vr.mod != x_like->Vr.mod) return ctx->fail(MPSK_ESPACE, "environments and vector belong to different groups");
  if (GL->nm() != GR->nm()) return ctx->fail(MPSK_ESPACE, "environments have different numbers of MPO levels");
  const int nlev = GL->nm();
  if (h1 && h1->odim != nlev) return ctx->fail(MPSK_ESPACE, "MPO slice odim %d != environment levels %d", h1->odim, nlev);
  std::unique_ptr<HeffOp> op(new HeffOp);
  // sharded: GL is this rank's row panel and apply must allgather the panels -- an explicit flag, because on small spaces
  // a rank's panel can coincide with the full space (or be empty) and every rank has to take the same collective path
  const bool sharded = (kind & MPSK_OP_SHARDED) != 0;
  kind &= 0xff;
  op->ctx = ctx; op->kind = kind; op->dtype = x_like->dtype;
  op->rectangular = sharded;
  const bool cx = op->dtype == MPSK_C128;
  MPSK_TRY(build_W(ctx, kind, h1, h2, x_like, nlev, &op->W));
  const Space& Vl = x_like->Vl;      // ket (contracted) left bond
  const Space& Vb = GL->Vl;          // bra left bond (== Vl unless this op computes a row panel)
  // charge rule  c_k = c_i + ch(m_out) - ch(m_in)  with c_i = -GL.mch[i], c_k = +GR.mch[k]
  for (const WEntry& w : op->W) {
    const int64_t lhs = Vl.norm((int64_t)GR->mch[w.k]);
    const int64_t rhs = Vl.norm(-(int64_t)GL->mch[w.i] + x_like->mch[w.mout] - x_like->mch[w.min]);
    if (lhs != rhs) return ctx->fail(MPSK_ESPACE, "MPO entry (%d,%d) element (%d,%d) violates charge conservation", w.i, w.k, w.mout, w.min);
  }
  // operands in their GEMM forms
  { Tensor* t = nullptr; MPSK_TRY(tensor_repack(ctx, GL, F_KR, false, &t)); op->GLk.reset(t); }
  { Tensor* t = nullptr; MPSK_TRY(tensor_repack(ctx, GR, F_SL, false, &t)); op->GRs.reset(t); }
  { Tensor* t = nullptr; MPSK_TRY(tensor_create(ctx, x_like->Vl, x_like->mch, x_like->Vr, x_like->dtype, F_KL, true, &t)); op->xs.reset(t); }
  if (ctx->profile) { cudaStreamSynchronize(ctx->stream); tp1 = wall(); }
  const Tensor& xs = *op->xs; const Tensor& GLk = *op->GLk; const Tensor& GRs = *op->GRs;
  Tensor* ybuf = nullptr;
  MPSK_TRY(tensor_create(ctx, Vb, x_like->mch, GR->Vr, x_like->dtype, F_KL, true, &ybuf));
  std::unique_ptr<Tensor> ys(ybuf);
  const int ngl = Vl.nsec();
  // ---- T workspace: per l-sector jb a (M_A x N_A) matrix
  std::vector<int64_t> Toff(ngl), Tld(ngl);
  int64_t tot = 0;
  for (int jb = 0; jb < ngl; ++jb) {
    Tld[jb] = pad(xs.groups[jb].rows);
    Toff[jb] = tot;
    tot += padg(GLk.groups[jb].rows * Tld[jb]);
  }
  op->T_elems = std::max<int64_t>(tot, GROUP_ALIGN);
  // ---- U workspace: per y block (m_out, ja) an (n_a x Kz) matrix
  std::vector<int64_t> Uoff(ys->blks.size());
  tot = 0;
  for (size_t b = 0; b < ys->blks.size(); ++b) {
    const Blk& bk = ys->blks[b];
    Uoff[b] = tot;
    tot += padg(Vb.n[bk.jl] * GRs.groups[bk.jr].ld);
  }
  op->U_elems = std::max<int64_t>(tot, GROUP_ALIGN);
  const size_t es = cx ? 16 : 8;
  op->T = ctx->lease(Ctx::S_HEFF_T, op->T_elems * es, &op->T_leased);
  op->U = ctx->lease(Ctx::S_HEFF_U, op->U_elems * es, &op->U_leased);
  if (!op->T || !op->U) return MPSK_ENOMEM;
  CUDA_TRY(ctx, cudaMemsetAsync(op->T, 0, op->T_elems * es, ctx->stream));
  CUDA_TRY(ctx, cudaMemsetAsync(op->U, 0, op->U_elems * es, ctx->stream));
  auto Tptr = [&](int64_t off) { return (char*)op->T + off * es; };
  auto Uptr = [&](int64_t off) { return (char*)op->U + off * es; };
  const double kappa = cx ? 8.0 : 2.0;
  op->flops_ = 0;
  // ---- stage A
  std::vector<GemmProblemHost> pa;
  for (int jb = 0; jb < ngl; ++jb) {
    const Group& ga = GLk.groups[jb]; const Group& gx = xs.groups[jb];
    if (ga.rows == 0 || gx.rows == 0) continue;
    pa.push_back(GemmProblemHost{GLk.ptr(ga.off), xs.ptr(gx.off), Tptr(Toff[jb]), ga.ld, gx.ld, Tld[jb], (int32_t)ga.rows,
                                 (int32_t)gx.rows, (int32_t)ga.cols, 0});
    op->flops_ += kappa * (double)ga.rows * gx.rows * ga.cols;
  }
  // ---- mix + stage C
  std::vector<MixTask> tasks; std::vector<MixTerm> terms; std::vector<GemmProblemHost> pc;
  const int ngr = GRs.Vr.nsec();
  for (size_t b = 0; b < ys->blks.size(); ++b) {
    const Blk& bk = ys->blks[b];
    const int ja = bk.jl, jz = bk.jr;
    const Group& gz = GRs.groups[jz];
    const int64_t ldU = gz.ld;
    double klog = 0;
    for (int k = 0; k < nlev; ++k) {
      const int64_t segk = GRs.seg[(size_t)k * ngr + jz];
      if (segk < 0) continue;
      const int jc = GRs.Vl.find((int64_t)GRs.Vr.q[jz] - GRs.mch[k]);
      klog += (double)GRs.Vl.n[jc];
      MixTask tk; tk.dst = Uptr(Uoff[b] + segk); tk.ldd = ldU; tk.rows = (int32_t)Vb.n[ja]; tk.cols = (int32_t)GRs.Vl.n[jc];
      tk.term_begin = (int32_t)terms.size();
      for (const WEntry& w : op->W) {
        if (w.k != k || w.mout != bk.m) continue;
        const int jb = Vl.find((int64_t)Vb.q[ja] + GLk.mch[w.i]);
        if (jb < 0) continue;
        const Blk* xb = xs.find(w.min, jb);
        if (!xb) continue;
        if (xb->jr != jc) return ctx->fail(MPSK_EINTERNAL, "sector bookkeeping mismatch in mix plan");
        const int64_t rowoff = GLk.seg[(size_t)w.i * ngl + jb];
        const int64_t coloff = xs.seg[(size_t)w.min * ngl + jb];
        if (rowoff < 0 || coloff < 0) return ctx->fail(MPSK_EINTERNAL, "missing segment in mix plan");
        terms.push_back(MixTerm{Tptr(Toff[jb] + rowoff * Tld[jb] + coloff), Tld[jb], w.c.real(), w.c.imag()});
      }
      tk.term_end = (int32_t)terms.size();
      if (tk.term_end > tk.term_begin) tasks.push_back(tk);
    }
    if (gz.cols == 0) continue;
    pc.push_back(GemmProblemHost{GRs.ptr(gz.off), Uptr(Uoff[b]), ys->ptr(bk.off), gz.ld, ldU, ys->groups[ja].ld, (int32_t)gz.rows,
                                 (int32_t)Vb.n[ja], (int32_t)gz.cols, 0});
    op->flops_ += kappa * (double)gz.rows * Vb.n[ja] * klog;
  }
  if (ctx->profile) { cudaStreamSynchronize(ctx->stream); tp2 = wall(); }
  op->gemmA.reset(new GemmPlan); op->gemmC.reset(new GemmPlan); op->mix.reset(new MixPlan);
  MPSK_TRY(gemm_plan_build(ctx, cx, pa, op->gemmA.get()));
  MPSK_TRY(mix_plan_build(ctx, cx, tasks, terms, op->mix.get()));
  MPSK_TRY(gemm_plan_build(ctx, cx, pc, op->gemmC.get()));
  op->ybuf.reset(ys.release());
  if (ctx->profile) { cudaStreamSynchronize(ctx->stream); fprintf(stderr, "[mpsk profile] heff_create: repack %.2f ms, workspaces+problem lists %.2f ms, plans %.2f ms\n", (tp1 - tp0) * 1e3, (tp2 - tp1) * 1e3, (wall() - tp2) * 1e3); }
  *out = op.release();
  return MPSK_OK;
}

// y_panel = (row panel of H_eff) x : the rank-local part of a sharded matvec (also the whole matvec when square)
int heff_apply_local(Ctx* ctx, HeffOp* op, const Tensor* x, Tensor* y) {
  if (!x->same_structure(*op->xs) || (y && !y->same_structure(*op->ybuf))) return ctx->fail(MPSK_ESPACE, "vector structure does not match the operator");
  if (x->form != F_KL || (y && y->form != F_KL)) return ctx->fail(MPSK_EINTERNAL, "Krylov vectors must be in KL form");
  if (x->d != op->xs->d)
    CUDA_TRY(ctx, cudaMemcpyAsync(op->xs->d, x->d, x->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  MPSK_TRY(gemm_plan_run(ctx, op->gemmA.get()));
  MPSK_TRY(mix_plan_run(ctx, op->mix.get()));
  MPSK_TRY(gemm_plan_run(ctx, op->gemmC.get()));
  if (y && y->d != op->ybuf->d)
    CUDA_TRY(ctx, cudaMemcpyAsync(y->d, op->ybuf->d, y->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  return MPSK_OK;
}

// allgather buffer (nranks slots + one send slot of the largest panel) and the two persistent scatter plans:
// gathered panels -> yfull (full-vector apply) and gathered panels -> xs (Krylov apply on panel vectors)
int HeffOp::ensure_gather() {
  if (gather) return MPSK_OK;
  if (ctx->nranks <= 1) return ctx->fail(MPSK_EINVAL, "row-panel operator used without a communicator");
  const Tensor* like_ = xs.get();
  const size_t es = like_->esize();
  int64_t mx = 0;
  for (int r = 0; r < ctx->nranks; ++r) mx = std::max(mx, panel_nelem(like_, r, ctx->nranks));
  if (panel_nelem(like_, ctx->rank, ctx->nranks) != ybuf->nelem) return ctx->fail(MPSK_ESPACE, "left environment is not this rank's row panel");
  gather_elems = mx;
  gather = ctx->alloc((size_t)mx * (ctx->nranks + 1) * es);
  if (!gather) return MPSK_ENOMEM;
  CUDA_TRY(ctx, cudaMemsetAsync(gather, 0, (size_t)mx * (ctx->nranks + 1) * es, ctx->stream));
  Tensor* yf = nullptr;
  MPSK_TRY(tensor_create(ctx, like_->Vl, like_->mch, like_->Vr, like_->dtype, F_KL, true, &yf));
  yfull.reset(yf);
  std::vector<CopyTask> ty, tx;
  for (int r = 0; r < ctx->nranks; ++r) {
    MPSK_TRY(scatter_panel_tasks(ctx, (char*)gather + (size_t)r * mx * es, r, ctx->nranks, yf, &ty));
    MPSK_TRY(scatter_panel_tasks(ctx, (char*)gather + (size_t)r * mx * es, r, ctx->nranks, xs.get(), &tx));
  }
  scatter.reset(new CopyPlan); scatter_x.reset(new CopyPlan);
  MPSK_TRY(copy_plan_build(ctx, ty, like_->dtype == MPSK_C128, scatter.get()));
  MPSK_TRY(copy_plan_build(ctx, tx, like_->dtype == MPSK_C128, scatter_x.get()));
  return MPSK_OK;
}

// y = H_eff x.  Square operator: local.  Row-panel operator on a multi-rank context: every rank computes its panel,
// one NCCL allgather over NVLink assembles the full y on all ranks (bitwise identical everywhere).
int heff_apply(Ctx* ctx, HeffOp* op, const Tensor* x, Tensor* y) {
  if (!op->rectangular) return heff_apply_local(ctx, op, x, y);
  if (!x->same_structure(*op->xs) || !y->same_structure(*op->xs)) return ctx->fail(MPSK_ESPACE, "vector structure does not match the operator");
  MPSK_TRY(op->ensure_gather());
  MPSK_TRY(heff_apply_local(ctx, op, x, nullptr));
  const size_t es = y->esize();
  char* send = (char*)op->gather + (size_t)op->gather_elems * ctx->nranks * es;
  CUDA_TRY(ctx, cudaMemcpyAsync(send, op->ybuf->d, op->ybuf->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  MPSK_TRY(comm_allgather(ctx, send, op->gather, (size_t)op->gather_elems * es));
  MPSK_TRY(copy_plan_run(ctx, op->scatter.get()));
  CUDA_TRY(ctx, cudaMemcpyAsync(y->d, op->yfull->d, y->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  return MPSK_OK;
}

// Panel in -> panel out (what a sharded Krylov solver iterates): allgather the panels of x straight into the operator's
// input buffer (NCCL over NVLink + one scatter pass), then the local three-launch pipeline.  Krylov basis vectors, their
// inner products and axpys stay 1/N per rank; no full-size vector other than the operator's input is ever materialised.
int heff_apply_panel(Ctx* ctx, HeffOp* op, const Tensor* xp, Tensor* yp) {
  if (!op->rectangular) return heff_apply_local(ctx, op, xp, yp);
  if (!xp->same_structure(*op->ybuf) || !yp->same_structure(*op->ybuf)) return ctx->fail(MPSK_ESPACE, "panel structure does not match the operator");
  MPSK_TRY(op->ensure_gather());
  const size_t es = xp->esize();
  char* send = (char*)op->gather + (size_t)op->gather_elems * ctx->nranks * es;
  CUDA_TRY(ctx, cudaMemcpyAsync(send, xp->d, xp->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  MPSK_TRY(comm_allgather(ctx, send, op->gather, (size_t)op->gather_elems * es));
  MPSK_TRY(copy_plan_run(ctx, op->scatter_x.get()));
  MPSK_TRY(gemm_plan_run(ctx, op->gemmA.get()));
  MPSK_TRY(mix_plan_run(ctx, op->mix.get()));
  MPSK_TRY(gemm_plan_run(ctx, op->gemmC.get()));
  if (yp->d != op->ybuf->d)
    CUDA_TRY(ctx, cudaMemcpyAsync(yp->d, op->ybuf->d, yp->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  return MPSK_OK;
}

int HeffOp::apply(const Tensor* x, Tensor* y) { return heff_apply(ctx, this, x, y); }
int HeffOp::apply_k(const Tensor* v, Tensor* w) { return heff_apply_panel(ctx, this, v, w); }

int HeffOp::to_local(const Tensor* x, Tensor* loc) {
  if (!rectangular) return LinOp::to_local(x, loc);
  if (!x->same_structure(*xs) || !loc->same_structure(*ybuf)) return ctx->fail(MPSK_ESPACE, "to_local: structure mismatch");
  Tensor* p = nullptr;
  MPSK_TRY(tensor_row_panel(ctx, x, ctx->rank, ctx->nranks, &p));
  std::unique_ptr<Tensor> pg(p);
  if (p->nelem != loc->nelem) return ctx->fail(MPSK_EINTERNAL, "to_local: panel size mismatch");
  CUDA_TRY(ctx, cudaMemcpyAsync(loc->d, p->d, p->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  return MPSK_OK;
}

int HeffOp::from_local(const Tensor* loc, Tensor* x) {
  if (!rectangular) return LinOp::from_local(loc, x);
  if (!x->same_structure(*xs) || !loc->same_structure(*ybuf)) return ctx->fail(MPSK_ESPACE, "from_local: structure mismatch");
  MPSK_TRY(ensure_gather());
  const size_t es = x->esize();
  char* send = (char*)gather + (size_t)gather_elems * ctx->nranks * es;
  CUDA_TRY(ctx, cudaMemcpyAsync(send, loc->d, loc->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  MPSK_TRY(comm_allgather(ctx, send, gather, (size_t)gather_elems * es));
  MPSK_TRY(copy_plan_run(ctx, scatter.get()));
  CUDA_TRY(ctx, cudaMemcpyAsync(x->d, yfull->d, x->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  return MPSK_OK;
}

int LinOp::to_local(const Tensor* x, Tensor* loc) {
  if (!x->same_structure(*loc) || x->form != loc->form) return ctx->fail(MPSK_ESPACE, "to_local: structure mismatch");
  if (x->d != loc->d) CUDA_TRY(ctx, cudaMemcpyAsync(loc->d, x->d, x->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  return MPSK_OK;
}
int LinOp::from_local(const Tensor* loc, Tensor* x) {
  if (!x->same_structure(*loc) || x->form != loc->form) return ctx->fail(MPSK_ESPACE, "from_local: structure mismatch");
  if (x->d != loc->d) CUDA_TRY(ctx, cudaMemcpyAsync(x->d, loc->d, x->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  return MPSK_OK;
}

}  // namespace mpsk
