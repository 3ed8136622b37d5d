// Environment updates (transfer.jl:137-259 through FinEnv.jl:114-145) and the local tensor algebra of the
// sweep bodies (dmrg.jl:92,108; orthoview.jl:95-106), all expressed as grouped NT GEMMs + one mix pass.
#include "internal.h"
#include <chrono>
#include <cstdlib>

namespace mpsk {

struct Work {
  Ctx* ctx; void* p = nullptr; size_t es; int64_t elems = 0;
  ~Work() { if (p) ctx->free(p); }
  int init(Ctx* c, int64_t n, size_t e) {
    ctx = c; es = e; elems = std::max<int64_t>(n, GROUP_ALIGN);
    p = ctx->alloc(elems * es);
    if (!p) return MPSK_ENOMEM;
    cudaError_t err = cudaMemsetAsync(p, 0, elems * es, ctx->stream);
    return err == cudaSuccess ? MPSK_OK : ctx->fail(MPSK_ECUDA, "memset failed");
  }
  char* at(int64_t off) const { return (char*)p + off * es; }
};

struct Owned {
  const Tensor* t = nullptr; bool own = false;
  ~Owned() { if (own) delete t; }
};

// ------------------------------------------------------------------------------------------ persistent transfer plans
TransferPlan::~TransferPlan() { if (ctx) { ctx->unlease(Ctx::S_TR_T, T, T_leased); ctx->unlease(Ctx::S_TR_Q, Q, Q_leased); ctx->free(gather); } }

// dst = (conj) src for two tensors of the same structure in (possibly) different forms; asynchronous
static int tensor_copy_into(Ctx* ctx, const Tensor* src, Tensor* dst, bool conj) {
  if (!src->same_structure(*dst)) return ctx->fail(MPSK_ESPACE, "tensor structures differ");
  const bool cj = conj && src->dtype == MPSK_C128;
  if (src->form == dst->form && !cj) {
    if (src->d != dst->d) CUDA_TRY(ctx, cudaMemcpyAsync(dst->d, src->d, src->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
    return MPSK_OK;
  }
  std::vector<CopyTask> tasks;
  tasks.reserve(src->blks.size());
  for (size_t b = 0; b < src->blks.size(); ++b) {
    const Blk& s = src->blks[b]; const Blk& d = dst->blks[b];
    tasks.push_back(CopyTask{src->ptr(s.off), dst->ptr(d.off), s.sl, s.sr, d.sl, d.sr, (int32_t)src->Vl.n[s.jl], (int32_t)src->Vr.n[s.jr], cj ? 1 : 0, 0, nullptr});
  }
  return run_copy_tasks(ctx, tasks, src->dtype == MPSK_C128);
}

static int alloc_work(Ctx* ctx, int slot, int64_t elems, size_t es, void** out, bool* leased) {
  elems = std::max<int64_t>(elems, GROUP_ALIGN);
  *out = ctx->lease(slot, (size_t)elems * es, leased);
  if (!*out) return MPSK_ENOMEM;
  CUDA_TRY(ctx, cudaMemsetAsync(*out, 0, (size_t)elems * es, ctx->stream));
  return MPSK_OK;
}

// GL'[a' k; b'] = sum GL[a j; b] A[b s; b'] O[j,k][s',s] conj(Abar[a s'; a'])        (transfer.jl:105-107,166-211)
static int build_left(Ctx* ctx, TransferPlan* P, const MpoSlice* h, const Tensor* A, const Tensor* Abar, const Tensor* GL) {
  if (A->dtype != GL->dtype || Abar->dtype != GL->dtype) return ctx->fail(MPSK_EINVAL, "mixed scalar types");
  if (!(GL->Vr == A->Vl)) return ctx->fail(MPSK_ESPACE, "transfer_left: ket space of GL != left space of A");
  if (!(GL->Vl == Abar->Vl)) return ctx->fail(MPSK_ESPACE, "transfer_left: bra space of GL != left space of Abar");
  if (h->odim != GL->nm() || h->d != A->nm() || h->d != Abar->nm()) return ctx->fail(MPSK_ESPACE, "transfer_left: MPO slice does not match");
  const bool cx = GL->dtype == MPSK_C128;
  const size_t es = cx ? 16 : 8;
  const double kappa = cx ? 8.0 : 2.0;
  const int od = h->odim, d = h->d;
  for (int j = 0; j < od; ++j) if (A->Vl.norm(-(int64_t)GL->mch[j]) != A->Vl.norm(h->cl[j])) return ctx->fail(MPSK_ESPACE, "transfer_left: level charges of GL and slice differ");
  { Tensor* t = nullptr; MPSK_TRY(tensor_repack(ctx, A, F_KL, false, &t)); P->Ak.reset(t); }
  { Tensor* t = nullptr; MPSK_TRY(tensor_repack(ctx, Abar, F_SL, true, &t)); P->Bs.reset(t); }
  { Tensor* t = nullptr; MPSK_TRY(tensor_create(ctx, GL->Vl, GL->mch, GL->Vr, GL->dtype, F_KR, true, &t)); P->vin.reset(t); }
  const Tensor* Ak = P->Ak.get(); const Tensor* Gk = P->vin.get(); const Tensor* Bs = P->Bs.get();
  std::vector<int32_t> mo(od);
  for (int k = 0; k < od; ++k) mo[k] = -h->cr[k];
  Tensor* o = nullptr;
  MPSK_TRY(tensor_create(ctx, Abar->Vr, mo, A->Vr, GL->dtype, F_KR, true, &o));
  P->vout.reset(o);
  { Tensor* t = nullptr; MPSK_TRY(tensor_create(ctx, Abar->Vr, mo, A->Vr, GL->dtype, F_KL, true, &t)); P->out_kl.reset(t); }
  const Space& Vb = A->Vl;
  const int nb = Vb.nsec();
  // step 1: T^T_jb[(s,b') x (j,a)] = A_KL(jb) * GL_KR(jb)^T
  std::vector<int64_t> Toff(nb), Tld(nb); int64_t tot = 0;
  for (int jb = 0; jb < nb; ++jb) { Tld[jb] = pad(Gk->groups[jb].rows); Toff[jb] = tot; tot += padg(Ak->groups[jb].rows * Tld[jb]); }
  MPSK_TRY(alloc_work(ctx, Ctx::S_TR_T, tot, es, &P->T, &P->T_leased));
  auto Tat = [&](int64_t off) { return (char*)P->T + off * es; };
  std::vector<GemmProblemHost> p1;
  for (int jb = 0; jb < nb; ++jb) {
    const Group& ga = Ak->groups[jb]; const Group& gg = Gk->groups[jb];
    if (!ga.rows || !gg.rows) continue;
    p1.push_back(GemmProblemHost{Ak->ptr(ga.off), Gk->ptr(gg.off), Tat(Toff[jb]), ga.ld, gg.ld, Tld[jb], (int32_t)ga.rows, (int32_t)gg.rows, (int32_t)ga.cols, 0});
    P->flops += kappa * (double)ga.rows * gg.rows * ga.cols;
  }
  // step 2/3
  std::vector<int64_t> Qoff(o->blks.size()); tot = 0;
  for (size_t b = 0; b < o->blks.size(); ++b) { Qoff[b] = tot; tot += padg(o->Vr.n[o->blks[b].jr] * Bs->groups[o->blks[b].jl].ld); }
  MPSK_TRY(alloc_work(ctx, Ctx::S_TR_Q, tot, es, &P->Q, &P->Q_leased));
  auto Qat = [&](int64_t off) { return (char*)P->Q + off * es; };
  std::vector<MixTask> tasks; std::vector<MixTerm> terms; std::vector<GemmProblemHost> p3;
  const int ngB = Bs->Vr.nsec();   // SL groups = r sectors of Abar (a')
  for (size_t b = 0; b < o->blks.size(); ++b) {
    const Blk& bk = o->blks[b];
    const int k = bk.m, jap = bk.jl, jbp = bk.jr;
    const Group& gB = Bs->groups[jap];
    for (int sp = 0; sp < d; ++sp) {
      const int64_t seg = Bs->seg[(size_t)sp * ngB + jap];
      if (seg < 0) continue;
      const int ja = Bs->Vl.find((int64_t)Bs->Vr.q[jap] - Bs->mch[sp]);
      MixTask tk; tk.dst = Qat(Qoff[b] + seg); tk.ldd = gB.ld; tk.rows = (int32_t)o->Vr.n[jbp]; tk.cols = (int32_t)Bs->Vl.n[ja];
      tk.term_begin = (int32_t)terms.size();
      for (int j = 0; j < od; ++j) {
        if (!h->kind[j * od + k]) continue;
        for (int s_ = 0; s_ < d; ++s_) {
          const cplx c = h->elem(j, k, sp, s_);
          if (c == cplx(0.0)) continue;
          const int jb = Vb.find((int64_t)o->Vr.q[jbp] - A->mch[s_]);
          if (jb < 0) continue;
          const Blk* ab = Ak->find(s_, jb);
          if (!ab || Ak->Vr.q[ab->jr] != o->Vr.q[jbp]) continue;
          const int jbra = Gk->Vl.find((int64_t)Vb.q[jb] - Gk->mch[j]);
          if (jbra < 0) continue;
          if (jbra != ja) return ctx->fail(MPSK_ESPACE, "transfer_left: MPO entry (%d,%d) violates charge conservation", j, k);
          const int64_t rowoff = Ak->seg[(size_t)s_ * nb + jb], coloff = Gk->seg[(size_t)j * nb + jb];
          if (rowoff < 0 || coloff < 0) return ctx->fail(MPSK_EINTERNAL, "transfer_left: missing segment");
          terms.push_back(MixTerm{Tat(Toff[jb] + rowoff * Tld[jb] + coloff), Tld[jb], c.real(), c.imag()});
        }
      }
      tk.term_end = (int32_t)terms.size();
      if (tk.term_end > tk.term_begin) tasks.push_back(tk);
    }
    if (!gB.cols) continue;
    p3.push_back(GemmProblemHost{Bs->ptr(gB.off), Qat(Qoff[b]), o->ptr(bk.off), gB.ld, gB.ld, o->groups[jbp].ld, (int32_t)gB.rows, (int32_t)o->Vr.n[jbp], (int32_t)gB.cols, 0});
    P->flops += kappa * (double)gB.rows * o->Vr.n[jbp] * gB.cols;
  }
  MPSK_TRY(gemm_plan_build(ctx, cx, p1, &P->g1));
  MPSK_TRY(mix_plan_build(ctx, cx, tasks, terms, &P->mx));
  MPSK_TRY(gemm_plan_build(ctx, cx, p3, &P->g3));
  return MPSK_OK;
}

// GR'[a j; b] = sum A[a s; c] O[j,k][s',s] conj(Abar[b s'; c']) GR[c k; c']            (transfer.jl:108-110,212-259)
static int build_right(Ctx* ctx, TransferPlan* P, const MpoSlice* h, const Tensor* A, const Tensor* Abar, const Tensor* GR) {
  if (A->dtype != GR->dtype || Abar->dtype != GR->dtype) return ctx->fail(MPSK_EINVAL, "mixed scalar types");
  if (!(GR->Vl == A->Vr)) return ctx->fail(MPSK_ESPACE, "transfer_right: ket space of GR != right space of A");
  if (!(GR->Vr == Abar->Vr)) return ctx->fail(MPSK_ESPACE, "transfer_right: bra space of GR != right space of Abar");
  if (h->odim != GR->nm() || h->d != A->nm() || h->d != Abar->nm()) return ctx->fail(MPSK_ESPACE, "transfer_right: MPO slice does not match");
  const bool cx = GR->dtype == MPSK_C128;
  const size_t es = cx ? 16 : 8;
  const double kappa = cx ? 8.0 : 2.0;
  const int od = h->odim, d = h->d;
  for (int k = 0; k < od; ++k) if (A->Vl.norm((int64_t)GR->mch[k]) != A->Vl.norm(h->cr[k])) return ctx->fail(MPSK_ESPACE, "transfer_right: level charges of GR and slice differ");
  { Tensor* t = nullptr; MPSK_TRY(tensor_repack(ctx, A, F_KR, false, &t)); P->Ak.reset(t); }
  { Tensor* t = nullptr; MPSK_TRY(tensor_repack(ctx, Abar, F_SR, true, &t)); P->Bs.reset(t); }
  { Tensor* t = nullptr; MPSK_TRY(tensor_create(ctx, GR->Vl, GR->mch, GR->Vr, GR->dtype, F_KL, true, &t)); P->vin.reset(t); }
  const Tensor* Ak = P->Ak.get(); const Tensor* Gk = P->vin.get(); const Tensor* Bs = P->Bs.get();
  std::vector<int32_t> mo(od);
  for (int j = 0; j < od; ++j) mo[j] = h->cl[j];
  Tensor* o = nullptr;
  MPSK_TRY(tensor_create(ctx, A->Vl, mo, Abar->Vl, GR->dtype, F_KL, true, &o));
  P->vout.reset(o);
  const Space& Vc = A->Vr;
  const int nc = Vc.nsec();
  // step 1: T^T_jc[(s,a) x (k,c')] = A_KR(jc) * GR_KL(jc)^T
  std::vector<int64_t> Toff(nc), Tld(nc); int64_t tot = 0;
  for (int jc = 0; jc < nc; ++jc) { Tld[jc] = pad(Gk->groups[jc].rows); Toff[jc] = tot; tot += padg(Ak->groups[jc].rows * Tld[jc]); }
  MPSK_TRY(alloc_work(ctx, Ctx::S_TR_T, tot, es, &P->T, &P->T_leased));
  auto Tat = [&](int64_t off) { return (char*)P->T + off * es; };
  std::vector<GemmProblemHost> p1;
  for (int jc = 0; jc < nc; ++jc) {
    const Group& ga = Ak->groups[jc]; const Group& gg = Gk->groups[jc];
    if (!ga.rows || !gg.rows) continue;
    p1.push_back(GemmProblemHost{Ak->ptr(ga.off), Gk->ptr(gg.off), Tat(Toff[jc]), ga.ld, gg.ld, Tld[jc], (int32_t)ga.rows, (int32_t)gg.rows, (int32_t)ga.cols, 0});
    P->flops += kappa * (double)ga.rows * gg.rows * ga.cols;
  }
  std::vector<int64_t> Qoff(o->blks.size()); tot = 0;
  for (size_t b = 0; b < o->blks.size(); ++b) { Qoff[b] = tot; tot += padg(o->Vl.n[o->blks[b].jl] * Bs->groups[o->blks[b].jr].ld); }
  MPSK_TRY(alloc_work(ctx, Ctx::S_TR_Q, tot, es, &P->Q, &P->Q_leased));
  auto Qat = [&](int64_t off) { return (char*)P->Q + off * es; };
  std::vector<MixTask> tasks; std::vector<MixTerm> terms; std::vector<GemmProblemHost> p3;
  const int ngB = Bs->Vl.nsec();   // SR groups = l sectors of Abar (b)
  for (size_t b = 0; b < o->blks.size(); ++b) {
    const Blk& bk = o->blks[b];
    const int j = bk.m, ja = bk.jl, jb = bk.jr;
    const Group& gB = Bs->groups[jb];
    for (int sp = 0; sp < d; ++sp) {
      const int64_t seg = Bs->seg[(size_t)sp * ngB + jb];
      if (seg < 0) continue;
      const int jcp = Bs->Vr.find((int64_t)Bs->Vl.q[jb] + Bs->mch[sp]);
      MixTask tk; tk.dst = Qat(Qoff[b] + seg); tk.ldd = gB.ld; tk.rows = (int32_t)o->Vl.n[ja]; tk.cols = (int32_t)Bs->Vr.n[jcp];
      tk.term_begin = (int32_t)terms.size();
      for (int k = 0; k < od; ++k) {
        if (!h->kind[j * od + k]) continue;
        for (int s_ = 0; s_ < d; ++s_) {
          const cplx c = h->elem(j, k, sp, s_);
          if (c == cplx(0.0)) continue;
          const Blk* ab = Ak->find(s_, ja);
          if (!ab) continue;
          const int jc = ab->jr;
          const int jbra = Gk->Vr.find((int64_t)Vc.q[jc] + Gk->mch[k]);
          if (jbra < 0) continue;
          if (jbra != jcp) return ctx->fail(MPSK_ESPACE, "transfer_right: MPO entry (%d,%d) violates charge conservation", j, k);
          const int64_t rowoff = Ak->seg[(size_t)s_ * nc + jc], coloff = Gk->seg[(size_t)k * nc + jc];
          if (rowoff < 0 || coloff < 0) return ctx->fail(MPSK_EINTERNAL, "transfer_right: missing segment");
          terms.push_back(MixTerm{Tat(Toff[jc] + rowoff * Tld[jc] + coloff), Tld[jc], c.real(), c.imag()});
        }
      }
      tk.term_end = (int32_t)terms.size();
      if (tk.term_end > tk.term_begin) tasks.push_back(tk);
    }
    if (!gB.cols) continue;
    p3.push_back(GemmProblemHost{Bs->ptr(gB.off), Qat(Qoff[b]), o->ptr(bk.off), gB.ld, gB.ld, o->groups[ja].ld, (int32_t)gB.rows, (int32_t)o->Vl.n[ja], (int32_t)gB.cols, 0});
    P->flops += kappa * (double)gB.rows * o->Vl.n[ja] * gB.cols;
  }
  MPSK_TRY(gemm_plan_build(ctx, cx, p1, &P->g1));
  MPSK_TRY(mix_plan_build(ctx, cx, tasks, terms, &P->mx));
  MPSK_TRY(gemm_plan_build(ctx, cx, p3, &P->g3));
  return MPSK_OK;
}

// Multi-GPU: the update is sharded over the OUTPUT KET index (b' of GL', a of GR'), i.e. over the free bond index of the
// ket tensor A.  Both GEMM stages then shrink by 1/N (a split over the bra index would leave stage 1 replicated), there
// are no partial sums, and one allgather assembles the full environment on every rank (bit-identical everywhere).
int transfer_plan_create(Ctx* ctx, bool left, const MpoSlice* h, const Tensor* A, const Tensor* Abar, const Tensor* v_like,
                         bool sharded, TransferPlan** out) {
  std::unique_ptr<TransferPlan> P(new TransferPlan);
  P->ctx = ctx; P->left = left; P->cx = v_like->dtype == MPSK_C128;
  sharded = sharded && ctx->nranks > 1;
  P->sharded = sharded;
  const Tensor* Ause = A; std::unique_ptr<Tensor> Apan;
  if (sharded) {
    Tensor* t = nullptr;
    MPSK_TRY(tensor_panel(ctx, A, left ? 1 : 0, ctx->rank, ctx->nranks, &t));
    Apan.reset(t); Ause = t;
  }
  MPSK_TRY(left ? build_left(ctx, P.get(), h, Ause, Abar, v_like) : build_right(ctx, P.get(), h, Ause, Abar, v_like));
  if (sharded) {
    const Tensor* res = left ? P->out_kl.get() : P->vout.get();
    Tensor* full = nullptr;
    if (left) MPSK_TRY(tensor_create(ctx, res->Vl, res->mch, A->Vr, res->dtype, F_KL, true, &full));
    else MPSK_TRY(tensor_create(ctx, A->Vl, res->mch, res->Vr, res->dtype, F_KL, true, &full));
    P->out_full.reset(full);
    const int axis = left ? 1 : 0;
    int64_t mx = 0;
    for (int r = 0; r < ctx->nranks; ++r) mx = std::max(mx, panel_nelem_axis(full, axis, r, ctx->nranks));
    if (panel_nelem_axis(full, axis, ctx->rank, ctx->nranks) != res->nelem) return ctx->fail(MPSK_EINTERNAL, "transfer: panel layout mismatch");
    const size_t es = res->esize();
    P->gather_elems = mx;
    P->gather = ctx->alloc((size_t)mx * (ctx->nranks + 1) * es);
    if (!P->gather) return MPSK_ENOMEM;
    CUDA_TRY(ctx, cudaMemsetAsync(P->gather, 0, (size_t)mx * (ctx->nranks + 1) * es, ctx->stream));
    std::vector<CopyTask> tasks;
    for (int r = 0; r < ctx->nranks; ++r)
      MPSK_TRY(scatter_panel_tasks_axis(ctx, (char*)P->gather + (size_t)r * mx * es, axis, r, ctx->nranks, full, &tasks));
    P->scatter.reset(new CopyPlan);
    MPSK_TRY(copy_plan_build(ctx, tasks, P->cx, P->scatter.get()));
  }
  *out = P.release();
  return MPSK_OK;
}

int transfer_plan_run(Ctx* ctx, TransferPlan* P, const Tensor* v) {
  MPSK_TRY(tensor_copy_into(ctx, v, P->vin.get(), false));
  MPSK_TRY(gemm_plan_run(ctx, &P->g1));
  MPSK_TRY(mix_plan_run(ctx, &P->mx));
  MPSK_TRY(gemm_plan_run(ctx, &P->g3));
  if (P->left) MPSK_TRY(tensor_copy_into(ctx, P->vout.get(), P->out_kl.get(), false));
  if (P->sharded) {
    const Tensor* res = P->left ? P->out_kl.get() : P->vout.get();
    const size_t es = res->esize();
    char* send = (char*)P->gather + (size_t)P->gather_elems * ctx->nranks * es;
    CUDA_TRY(ctx, cudaMemcpyAsync(send, res->d, res->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
    MPSK_TRY(comm_allgather(ctx, send, P->gather, (size_t)P->gather_elems * es));
    MPSK_TRY(copy_plan_run(ctx, P->scatter.get()));
  }
  return MPSK_OK;
}

static double wall_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int transfer_once(Ctx* ctx, bool left, const MpoSlice* h, const Tensor* A, const Tensor* Abar, const Tensor* v, Tensor** out) {
  if (ctx->profile < 0) ctx->profile = getenv("MPSK_PROFILE") ? 1 : 0;
  double t0 = 0, t1 = 0;
  if (ctx->profile) { cudaStreamSynchronize(ctx->stream); t0 = wall_s(); }
  TransferPlan* P = nullptr;
  MPSK_TRY(transfer_plan_create(ctx, left, h, A, Abar, v, /*sharded=*/!(ctx->flags & MPSK_FLAG_REPLICATE_LOCAL), &P));
  std::unique_ptr<TransferPlan> pg(P);
  if (ctx->profile) { cudaStreamSynchronize(ctx->stream); t1 = wall_s(); }
  MPSK_TRY(transfer_plan_run(ctx, P, v));
  if (ctx->profile) { cudaStreamSynchronize(ctx->stream); fprintf(stderr, "[mpsk profile] transfer_%s: plan %.2f ms, run %.2f ms (sharded %d)\n", left ? "left" : "right", (t1 - t0) * 1e3, (wall_s() - t1) * 1e3, (int)P->sharded); }
  // hand the result tensor over to the caller (the plan and its workspaces are released in stream order)
  if (P->sharded) *out = P->out_full.release();
  else if (left) *out = P->out_kl.release();
  else *out = P->vout.release();
  return MPSK_OK;
}

int transfer_left(Ctx* ctx, const MpoSlice* h, const Tensor* A, const Tensor* Abar, const Tensor* GL, Tensor** out) {
  return transfer_once(ctx, true, h, A, Abar, GL, out);
}
int transfer_right(Ctx* ctx, const MpoSlice* h, const Tensor* A, const Tensor* Abar, const Tensor* GR, Tensor** out) {
  return transfer_once(ctx, false, h, A, Abar, GR, out);
}

// ------------------------------------------------------------------------------------------ transfer matrix as a linear operator
TmOp::~TmOp() { if (ctx && d_coef) ctx->free(d_coef); }

int TmOp::apply(const Tensor* x, Tensor* y) {
  if (!x->same_structure(*vlike) || !y->same_structure(*vlike)) return ctx->fail(MPSK_ESPACE, "transfer matrix: vector structure mismatch");
  const Tensor* cur = x;
  for (auto& sp : sites) {
    MPSK_TRY(transfer_plan_run(ctx, sp.get(), cur));
    cur = sp->result();
  }
  if (!cur->same_structure(*y)) return ctx->fail(MPSK_ESPACE, "transfer matrix does not map the space onto itself");
  CUDA_TRY(ctx, cudaMemcpyAsync(y->d, cur->d, y->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  if (ladj) {   // regularisation (transfermatrix.jl:68-76): y -= <l, y> r, the coefficient never leaves the device
    const bool cx = y->dtype == MPSK_C128;
    const void* V[1] = {ladj->d};
    MPSK_TRY(vec_multidot(ctx, cx, 1, V, y->d, y->nelem, d_coef));
    const void* R[1] = {rvec->d};
    MPSK_TRY(vec_multiaxpy(ctx, cx, 1, R, d_coef, -1.0, y->d, y->nelem));
  }
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ multi-row operators
// copy `n` consecutive labels (all their blocks) between two tensors with the same bond spaces; asynchronous
static int copy_labels(Ctx* ctx, const Tensor* src, int s0, Tensor* dst, int d0, int n) {
  if (!(src->Vl == dst->Vl) || !(src->Vr == dst->Vr) || src->dtype != dst->dtype) return ctx->fail(MPSK_ESPACE, "copy_labels: bond spaces differ");
  std::vector<CopyTask> tasks;
  for (int m = 0; m < n; ++m) {
    if (src->mch[s0 + m] != dst->mch[d0 + m]) return ctx->fail(MPSK_ESPACE, "copy_labels: label charges differ");
    for (int jl = 0; jl < src->Vl.nsec(); ++jl) {
      const Blk* a = src->find(s0 + m, jl); const Blk* b = dst->find(d0 + m, jl);
      if (!a || !b) continue;
      tasks.push_back(CopyTask{src->ptr(a->off), dst->ptr(b->off), a->sl, a->sr, b->sl, b->sr, (int32_t)src->Vl.n[a->jl], (int32_t)src->Vr.n[a->jr], 0, 0, nullptr});
    }
  }
  return run_copy_tasks(ctx, tasks, src->dtype == MPSK_C128);
}

int multiop_create(Ctx* ctx, const std::vector<LinOp*>& ops, int shift, MultiOp** out) {
  if (ops.empty()) return ctx->fail(MPSK_EINVAL, "multiop: no operators");
  std::unique_ptr<MultiOp> M(new MultiOp);
  M->ctx = ctx; M->ops = ops; M->shift = shift; M->dtype = ops[0]->dtype;
  const Tensor* l0 = ops[0]->like();
  std::vector<int32_t> mch;
  for (LinOp* o : ops) {
    const Tensor* l = o->like();
    if (o->sharded()) return ctx->fail(MPSK_EINVAL, "multiop: row-panel operators are not supported");
    if (!(l->Vl == l0->Vl) || !(l->Vr == l0->Vr) || l->dtype != l0->dtype || l->mch != l0->mch)
      return ctx->fail(MPSK_ESPACE, "multiop: all rows must act on vectors of the same structure");
    mch.insert(mch.end(), l->mch.begin(), l->mch.end());
    Tensor* a = nullptr; Tensor* b = nullptr;
    MPSK_TRY(tensor_create(ctx, l->Vl, l->mch, l->Vr, l->dtype, F_KL, true, &a)); M->xin.emplace_back(a);
    MPSK_TRY(tensor_create(ctx, l->Vl, l->mch, l->Vr, l->dtype, F_KL, true, &b)); M->yout.emplace_back(b);
  }
  Tensor* v = nullptr;
  MPSK_TRY(tensor_create(ctx, l0->Vl, mch, l0->Vr, l0->dtype, F_KL, true, &v));
  M->vlike.reset(v);
  *out = M.release();
  return MPSK_OK;
}

int MultiOp::apply(const Tensor* x, Tensor* y) {
  if (!x->same_structure(*vlike) || !y->same_structure(*vlike)) return ctx->fail(MPSK_ESPACE, "multiop: vector structure mismatch");
  const int n = (int)ops.size(), nm = ops[0]->like()->nm();
  for (int r = 0; r < n; ++r) {
    MPSK_TRY(copy_labels(ctx, x, r * nm, xin[r].get(), 0, nm));
    MPSK_TRY(ops[r]->apply(xin[r].get(), yout[r].get()));
    const int tr = ((r + shift) % n + n) % n;
    MPSK_TRY(copy_labels(ctx, yout[r].get(), 0, y, tr * nm, nm));
  }
  return MPSK_OK;
}

// ac2[l (p1 p2) r] = sum_k A1[l p1 k] A2[k p2 r]     (dmrg.jl:92,108)
int contract_ac2(Ctx* ctx, const Tensor* A1, const Tensor* A2, Tensor** out) {
  if (A1->dtype != A2->dtype) return ctx->fail(MPSK_EINVAL, "mixed scalar types");
  if (!(A1->Vr == A2->Vl)) return ctx->fail(MPSK_ESPACE, "contract_ac2: bond spaces differ");
  const bool cx = A1->dtype == MPSK_C128;
  const int d1 = A1->nm(), d2 = A2->nm();
  std::vector<int32_t> mch(d1 * d2);
  for (int u = 0; u < d1; ++u) for (int t = 0; t < d2; ++t) mch[u * d2 + t] = A1->mch[u] + A2->mch[t];
  Owned a1, a2;
  MPSK_TRY(tensor_as_form(ctx, A1, F_SR, false, &a1.t, &a1.own));
  MPSK_TRY(tensor_as_form(ctx, A2, F_SL, false, &a2.t, &a2.own));
  Tensor* x = nullptr;
  MPSK_TRY(tensor_create(ctx, A1->Vl, mch, A2->Vr, A1->dtype, F_KL, true, &x));
  std::unique_ptr<Tensor> xg(x);
  std::vector<GemmProblemHost> pr;
  for (const Blk& bk : x->blks) {
    const int u = bk.m / d2, t = bk.m % d2, jb = bk.jl, jc = bk.jr;
    const Blk* b1 = a1.t->find(u, jb);
    if (!b1) continue;
    const int jk = b1->jr;
    const Blk* b2 = a2.t->find(t, jk);
    if (!b2 || b2->jr != jc) continue;
    pr.push_back(GemmProblemHost{a2.t->ptr(b2->off), a1.t->ptr(b1->off), x->ptr(bk.off), a2.t->groups[jc].ld, a1.t->groups[jb].ld,
                                 x->groups[jb].ld, (int32_t)x->Vr.n[jc], (int32_t)x->Vl.n[jb], (int32_t)A1->Vr.n[jk], 0});
  }
  GemmPlan g;
  MPSK_TRY(gemm_plan_build(ctx, cx, pr, &g));
  MPSK_TRY(gemm_plan_run(ctx, &g));
  *out = xg.release();
  return MPSK_OK;
}

// B[k' p2; r] = sum_{l, p1} conj(N[l p1; k']) x2[l (p1 p2); r]: the adjoint of contract_ac2 in its first factor
// (adjoint(VL) * AC2 of changebonds/optimalexpand.jl:30,98).  One K-stacked GEMM per output block: N (conjugated) and the
// p2-slice of x2 in SL form share the K layout (segments (p1, l-sector)).
int contract_adjoint_left(Ctx* ctx, const Tensor* N, const Tensor* x2, Tensor** out) {
  if (N->dtype != x2->dtype) return ctx->fail(MPSK_EINVAL, "mixed scalar types");
  if (!(N->Vl == x2->Vl)) return ctx->fail(MPSK_ESPACE, "contract_adjoint_left: left bonds differ");
  const int d1 = N->nm();
  if (d1 <= 0 || x2->nm() % d1) return ctx->fail(MPSK_ESPACE, "contract_adjoint_left: label counts do not match");
  const int d2 = x2->nm() / d1;
  const bool cx = N->dtype == MPSK_C128;
  std::vector<int32_t> s2(d2);
  for (int t = 0; t < d2; ++t) s2[t] = x2->mch[t] - N->mch[0];
  for (int u = 0; u < d1; ++u) for (int t = 0; t < d2; ++t)
    if (x2->mch[u * d2 + t] != N->mch[u] + s2[t]) return ctx->fail(MPSK_ESPACE, "contract_adjoint_left: label charges do not factorise");
  Tensor* o = nullptr;
  MPSK_TRY(tensor_create(ctx, N->Vr, s2, x2->Vr, N->dtype, F_KL, true, &o));
  std::unique_ptr<Tensor> og(o);
  Tensor* Ns = nullptr;
  MPSK_TRY(tensor_repack(ctx, N, F_SL, true, &Ns));
  std::unique_ptr<Tensor> nsg(Ns);
  std::vector<std::unique_ptr<Tensor>> keep;
  std::vector<GemmProblemHost> pr;
  for (int t = 0; t < d2; ++t) {
    std::vector<int32_t> lab(d1);
    for (int u = 0; u < d1; ++u) lab[u] = u * d2 + t;
    Tensor* Xp = nullptr; Tensor* Xs = nullptr;
    MPSK_TRY(tensor_select_labels(ctx, x2, lab.data(), d1, &Xp));
    std::unique_ptr<Tensor> xpg(Xp);
    MPSK_TRY(tensor_repack(ctx, Xp, F_SL, false, &Xs));
    keep.emplace_back(Xs);
    for (int jk = 0; jk < o->Vl.nsec(); ++jk) {
      const Blk* b = o->find(t, jk);
      if (!b) continue;
      const Group& gN = Ns->groups[jk]; const Group& gX = Xs->groups[b->jr];
      if (gN.cols == 0 || gX.cols == 0) continue;
      if (gN.cols != gX.cols) return ctx->fail(MPSK_EINTERNAL, "contract_adjoint_left: K layouts differ");
      pr.push_back(GemmProblemHost{Xs->ptr(gX.off), Ns->ptr(gN.off), o->ptr(b->off), gX.ld, gN.ld, o->groups[jk].ld,
                                   (int32_t)gX.rows, (int32_t)gN.rows, (int32_t)gN.cols, 0});
    }
  }
  GemmPlan g;
  MPSK_TRY(gemm_plan_build(ctx, cx, pr, &g));
  MPSK_TRY(gemm_plan_run(ctx, &g));
  *out = og.release();
  return MPSK_OK;
}

// direct sum along the left (axis 0) or right (axis 1) bond: catcodomain / catdomain of TensorKit
// (changebonds/optimalexpand.jl:103-104, changebonds.jl:17-32).  B may be NULL with `extra` giving a space to pad with zeros.
int tensor_cat(Ctx* ctx, const Tensor* A, const Tensor* B, int axis, Tensor** out) {
  if (A->dtype != B->dtype || A->mch != B->mch) return ctx->fail(MPSK_ESPACE, "cat: labels / scalar types differ");
  if (axis == 0 ? !(A->Vr == B->Vr) : !(A->Vl == B->Vl)) return ctx->fail(MPSK_ESPACE, "cat: the other bond differs");
  const Space& Sa = axis == 0 ? A->Vl : A->Vr; const Space& Sb = axis == 0 ? B->Vl : B->Vr;
  if (Sa.mod != Sb.mod) return ctx->fail(MPSK_ESPACE, "cat: different groups");
  std::map<int32_t, int64_t> dims;
  for (int j = 0; j < Sa.nsec(); ++j) dims[Sa.q[j]] += Sa.n[j];
  for (int j = 0; j < Sb.nsec(); ++j) dims[Sb.q[j]] += Sb.n[j];
  Space Sn; Sn.mod = Sa.mod;
  for (auto& kv : dims) { Sn.q.push_back(kv.first); Sn.n.push_back(kv.second); }
  Tensor* o = nullptr;
  if (axis == 0) MPSK_TRY(tensor_create(ctx, Sn, A->mch, A->Vr, A->dtype, F_KL, true, &o));
  else MPSK_TRY(tensor_create(ctx, A->Vl, A->mch, Sn, A->dtype, F_KL, true, &o));
  std::unique_ptr<Tensor> og(o);
  std::vector<CopyTask> tasks;
  for (int which = 0; which < 2; ++which) {
    const Tensor* T = which ? B : A;
    const Space& St = axis == 0 ? T->Vl : T->Vr;
    for (const Blk& b : T->blks) {
      const int js = axis == 0 ? b.jl : b.jr;
      const int jn = Sn.find(St.q[js]);
      int64_t shift = 0;
      if (which) { const int ja = Sa.find(St.q[js]); shift = ja >= 0 ? Sa.n[ja] : 0; }
      const Blk* d = axis == 0 ? o->find(b.m, jn) : o->find(b.m, b.jl);
      if (!d || (axis == 1 && d->jr != jn)) return ctx->fail(MPSK_EINTERNAL, "cat: output block mismatch");
      char* dst = o->ptr(d->off + shift * (axis == 0 ? d->sl : d->sr));
      tasks.push_back(CopyTask{T->ptr(b.off), dst, b.sl, b.sr, d->sl, d->sr, (int32_t)T->Vl.n[b.jl], (int32_t)T->Vr.n[b.jr], 0, 0, nullptr});
    }
  }
  MPSK_TRY(run_copy_tasks(ctx, tasks, A->dtype == MPSK_C128));
  *out = og.release();
  return MPSK_OK;
}

// A'[l p r'] = sum_r A[l p r] C[r r']
int mul_bond_right(Ctx* ctx, const Tensor* A, const Tensor* C, Tensor** out) {
  if (A->dtype != C->dtype) return ctx->fail(MPSK_EINVAL, "mixed scalar types");
  if (!(A->Vr == C->Vl) || C->nm() != 1 || C->mch[0] != 0) return ctx->fail(MPSK_ESPACE, "mul_bond_right: spaces differ");
  const bool cx = A->dtype == MPSK_C128;
  Owned a, c;
  MPSK_TRY(tensor_as_form(ctx, A, F_SR, false, &a.t, &a.own));
  MPSK_TRY(tensor_as_form(ctx, C, F_SL, false, &c.t, &c.own));
  Tensor* o = nullptr;
  MPSK_TRY(tensor_create(ctx, A->Vl, A->mch, C->Vr, A->dtype, F_KL, true, &o));
  std::unique_ptr<Tensor> og(o);
  std::vector<GemmProblemHost> pr;
  for (const Blk& bk : o->blks) {
    const Blk* ab = a.t->find(bk.m, bk.jl);
    if (!ab) continue;
    const Blk* cb = c.t->find(0, ab->jr);
    if (!cb || cb->jr != bk.jr) continue;
    pr.push_back(GemmProblemHost{c.t->ptr(cb->off), a.t->ptr(ab->off), o->ptr(bk.off), c.t->groups[cb->jr].ld, a.t->groups[bk.jl].ld,
                                 o->groups[bk.jl].ld, (int32_t)o->Vr.n[bk.jr], (int32_t)o->Vl.n[bk.jl], (int32_t)A->Vr.n[ab->jr], 0});
  }
  GemmPlan g;
  MPSK_TRY(gemm_plan_build(ctx, cx, pr, &g));
  MPSK_TRY(gemm_plan_run(ctx, &g));
  *out = og.release();
  return MPSK_OK;
}

// A'[l' p r] = sum_l C[l' l] A[l p r]
int mul_bond_left(Ctx* ctx, const Tensor* C, const Tensor* A, Tensor** out) {
  if (A->dtype != C->dtype) return ctx->fail(MPSK_EINVAL, "mixed scalar types");
  if (!(C->Vr == A->Vl) || C->nm() != 1 || C->mch[0] != 0) return ctx->fail(MPSK_ESPACE, "mul_bond_left: spaces differ");
  const bool cx = A->dtype == MPSK_C128;
  Owned a, c;
  MPSK_TRY(tensor_as_form(ctx, A, F_KL, false, &a.t, &a.own));
  MPSK_TRY(tensor_as_form(ctx, C, F_SR, false, &c.t, &c.own));
  Tensor* o = nullptr;
  MPSK_TRY(tensor_create(ctx, C->Vl, A->mch, A->Vr, A->dtype, F_KL, true, &o));
  std::unique_ptr<Tensor> og(o);
  std::vector<GemmProblemHost> pr;
  for (const Blk& bk : o->blks) {
    const Blk* cb = c.t->find(0, bk.jl);
    if (!cb) continue;
    const Blk* ab = a.t->find(bk.m, cb->jr);
    if (!ab || ab->jr != bk.jr) continue;
    pr.push_back(GemmProblemHost{a.t->ptr(ab->off), c.t->ptr(cb->off), o->ptr(bk.off), a.t->groups[ab->jl].ld, c.t->groups[bk.jl].ld,
                                 o->groups[bk.jl].ld, (int32_t)o->Vr.n[bk.jr], (int32_t)o->Vl.n[bk.jl], (int32_t)A->Vl.n[ab->jl], 0});
  }
  GemmPlan g;
  MPSK_TRY(gemm_plan_build(ctx, cx, pr, &g));
  MPSK_TRY(gemm_plan_run(ctx, &g));
  *out = og.release();
  return MPSK_OK;
}

}  // namespace mpsk
