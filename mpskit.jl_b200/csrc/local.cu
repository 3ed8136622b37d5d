// Environment updates (transfer.jl:137-259 through FinEnv.jl:114-145) and the local tensor algebra of the
// sweep bodies (dmrg.jl:92,108; orthoview.jl:95-106), all expressed as grouped NT GEMMs + one mix pass.
#include "internal.h"

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

// GL'[a' k; b'] = sum GL[a j; b] A[b s; b'] O[j,k][s',s] conj(Abar[a s'; a'])        (transfer.jl:105-107,166-211)
int transfer_left(Ctx* ctx, const MpoSlice* h, const Tensor* A, const Tensor* Abar, const Tensor* GL, Tensor** out) {
  if (A->dtype != GL->dtype || Abar->dtype != GL->dtype) return ctx->fail(MPSK_EINVAL, "mixed scalar types");
  if (!(GL->Vr == A->Vl)) return ctx->fail(MPSK_ESPACE, "transfer_left: ket space of GL != left space of A");
  if (!(GL->Vl == Abar->Vl)) return ctx->fail(MPSK_ESPACE, "transfer_left: bra space of GL != left space of Abar");
  if (h->odim != GL->nm() || h->d != A->nm() || h->d != Abar->nm()) return ctx->fail(MPSK_ESPACE, "transfer_left: MPO slice does not match");
  const bool cx = GL->dtype == MPSK_C128;
  const size_t es = cx ? 16 : 8;
  const int od = h->odim, d = h->d;
  for (int j = 0; j < od; ++j) if (A->Vl.norm(-(int64_t)GL->mch[j]) != A->Vl.norm(h->cl[j])) return ctx->fail(MPSK_ESPACE, "transfer_left: level charges of GL and slice differ");
  Owned Ak, Gk, Bs;
  MPSK_TRY(tensor_as_form(ctx, A, F_KL, false, &Ak.t, &Ak.own));
  MPSK_TRY(tensor_as_form(ctx, GL, F_KR, false, &Gk.t, &Gk.own));
  MPSK_TRY(tensor_as_form(ctx, Abar, F_SL, true, &Bs.t, &Bs.own));
  std::vector<int32_t> mo(od);
  for (int k = 0; k < od; ++k) mo[k] = -h->cr[k];
  Tensor* o = nullptr;
  MPSK_TRY(tensor_create(ctx, Abar->Vr, mo, A->Vr, GL->dtype, F_KR, true, &o));
  std::unique_ptr<Tensor> og(o);
  const Space& Vb = A->Vl;
  const int nb = Vb.nsec();
  // step 1: T^T_jb[(s,b') x (j,a)] = A_KL(jb) * GL_KR(jb)^T
  std::vector<int64_t> Toff(nb), Tld(nb); int64_t tot = 0;
  for (int jb = 0; jb < nb; ++jb) { Tld[jb] = pad(Gk.t->groups[jb].rows); Toff[jb] = tot; tot += padg(Ak.t->groups[jb].rows * Tld[jb]); }
  Work T; MPSK_TRY(T.init(ctx, tot, es));
  std::vector<GemmProblemHost> p1;
  for (int jb = 0; jb < nb; ++jb) {
    const Group& ga = Ak.t->groups[jb]; const Group& gg = Gk.t->groups[jb];
    if (!ga.rows || !gg.rows) continue;
    p1.push_back(GemmProblemHost{Ak.t->ptr(ga.off), Gk.t->ptr(gg.off), T.at(Toff[jb]), ga.ld, gg.ld, Tld[jb], (int32_t)ga.rows, (int32_t)gg.rows, (int32_t)ga.cols, 0});
  }
  // step 2/3
  std::vector<int64_t> Qoff(o->blks.size()); tot = 0;
  for (size_t b = 0; b < o->blks.size(); ++b) { Qoff[b] = tot; tot += padg(o->Vr.n[o->blks[b].jr] * Bs.t->groups[o->blks[b].jl].ld); }
  Work Q; MPSK_TRY(Q.init(ctx, tot, es));
  std::vector<MixTask> tasks; std::vector<MixTerm> terms; std::vector<GemmProblemHost> p3;
  const int ngB = Bs.t->Vr.nsec();   // SL groups = r sectors of Abar (a')
  for (size_t b = 0; b < o->blks.size(); ++b) {
    const Blk& bk = o->blks[b];
    const int k = bk.m, jap = bk.jl, jbp = bk.jr;
    const Group& gB = Bs.t->groups[jap];
    for (int sp = 0; sp < d; ++sp) {
      const int64_t seg = Bs.t->seg[(size_t)sp * ngB + jap];
      if (seg < 0) continue;
      const int ja = Bs.t->Vl.find((int64_t)Bs.t->Vr.q[jap] - Bs.t->mch[sp]);
      MixTask tk; tk.dst = Q.at(Qoff[b] + seg); tk.ldd = gB.ld; tk.rows = (int32_t)o->Vr.n[jbp]; tk.cols = (int32_t)Bs.t->Vl.n[ja];
      tk.term_begin = (int32_t)terms.size();
      for (int j = 0; j < od; ++j) {
        if (!h->kind[j * od + k]) continue;
        for (int s = 0; s < d; ++s) {
          const cplx c = h->elem(j, k, sp, s);
          if (c == cplx(0.0)) continue;
          const int jb = Vb.find((int64_t)A->Vr.q[jbp] - A->mch[s]);
          if (jb < 0) continue;
          const Blk* ab = Ak.t->find(s, jb);
          if (!ab || ab->jr != jbp) continue;
          const int jbra = Gk.t->Vl.find((int64_t)Vb.q[jb] - Gk.t->mch[j]);
          if (jbra < 0) continue;
          if (jbra != ja) return ctx->fail(MPSK_ESPACE, "transfer_left: MPO entry (%d,%d) violates charge conservation", j, k);
          const int64_t rowoff = Ak.t->seg[(size_t)s * nb + jb], coloff = Gk.t->seg[(size_t)j * nb + jb];
          if (rowoff < 0 || coloff < 0) return ctx->fail(MPSK_EINTERNAL, "transfer_left: missing segment");
          terms.push_back(MixTerm{T.at(Toff[jb] + rowoff * Tld[jb] + coloff), Tld[jb], c.real(), c.imag()});
        }
      }
      tk.term_end = (int32_t)terms.size();
      if (tk.term_end > tk.term_begin) tasks.push_back(tk);
    }
    if (!gB.cols) continue;
    p3.push_back(GemmProblemHost{Bs.t->ptr(gB.off), Q.at(Qoff[b]), o->ptr(bk.off), gB.ld, gB.ld, o->groups[jbp].ld, (int32_t)gB.rows, (int32_t)o->Vr.n[jbp], (int32_t)gB.cols, 0});
  }
  GemmPlan g1, g3; MixPlan mx;
  MPSK_TRY(gemm_plan_build(ctx, cx, p1, &g1));
  MPSK_TRY(mix_plan_build(ctx, cx, tasks, terms, &mx));
  MPSK_TRY(gemm_plan_build(ctx, cx, p3, &g3));
  MPSK_TRY(gemm_plan_run(ctx, &g1));
  MPSK_TRY(mix_plan_run(ctx, &mx));
  MPSK_TRY(gemm_plan_run(ctx, &g3));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // plans and workspaces die with this scope
  *out = og.release();
  return MPSK_OK;
}

// GR'[a j; b] = sum A[a s; c] O[j,k][s',s] conj(Abar[b s'; c']) GR[c k; c']            (transfer.jl:108-110,212-259)
int transfer_right(Ctx* ctx, const MpoSlice* h, const Tensor* A, const Tensor* Abar, const Tensor* GR, Tensor** out) {
  if (A->dtype != GR->dtype || Abar->dtype != GR->dtype) return ctx->fail(MPSK_EINVAL, "mixed scalar types");
  if (!(GR->Vl == A->Vr)) return ctx->fail(MPSK_ESPACE, "transfer_right: ket space of GR != right space of A");
  if (!(GR->Vr == Abar->Vr)) return ctx->fail(MPSK_ESPACE, "transfer_right: bra space of GR != right space of Abar");
  if (h->odim != GR->nm() || h->d != A->nm() || h->d != Abar->nm()) return ctx->fail(MPSK_ESPACE, "transfer_right: MPO slice does not match");
  const bool cx = GR->dtype == MPSK_C128;
  const size_t es = cx ? 16 : 8;
  const int od = h->odim, d = h->d;
  for (int k = 0; k < od; ++k) if (A->Vl.norm((int64_t)GR->mch[k]) != A->Vl.norm(h->cr[k])) return ctx->fail(MPSK_ESPACE, "transfer_right: level charges of GR and slice differ");
  Owned Ak, Gk, Bs;
  MPSK_TRY(tensor_as_form(ctx, A, F_KR, false, &Ak.t, &Ak.own));
  MPSK_TRY(tensor_as_form(ctx, GR, F_KL, false, &Gk.t, &Gk.own));
  MPSK_TRY(tensor_as_form(ctx, Abar, F_SR, true, &Bs.t, &Bs.own));
  std::vector<int32_t> mo(od);
  for (int j = 0; j < od; ++j) mo[j] = h->cl[j];
  Tensor* o = nullptr;
  MPSK_TRY(tensor_create(ctx, A->Vl, mo, Abar->Vl, GR->dtype, F_KL, true, &o));
  std::unique_ptr<Tensor> og(o);
  const Space& Vc = A->Vr;
  const int nc = Vc.nsec();
  // step 1: T^T_jc[(s,a) x (k,c')] = A_KR(jc) * GR_KL(jc)^T
  std::vector<int64_t> Toff(nc), Tld(nc); int64_t tot = 0;
  for (int jc = 0; jc < nc; ++jc) { Tld[jc] = pad(Gk.t->groups[jc].rows); Toff[jc] = tot; tot += padg(Ak.t->groups[jc].rows * Tld[jc]); }
  Work T; MPSK_TRY(T.init(ctx, tot, es));
  std::vector<GemmProblemHost> p1;
  for (int jc = 0; jc < nc; ++jc) {
    const Group& ga = Ak.t->groups[jc]; const Group& gg = Gk.t->groups[jc];
    if (!ga.rows || !gg.rows) continue;
    p1.push_back(GemmProblemHost{Ak.t->ptr(ga.off), Gk.t->ptr(gg.off), T.at(Toff[jc]), ga.ld, gg.ld, Tld[jc], (int32_t)ga.rows, (int32_t)gg.rows, (int32_t)ga.cols, 0});
  }
  std::vector<int64_t> Qoff(o->blks.size()); tot = 0;
  for (size_t b = 0; b < o->blks.size(); ++b) { Qoff[b] = tot; tot += padg(o->Vl.n[o->blks[b].jl] * Bs.t->groups[o->blks[b].jr].ld); }
  Work Q; MPSK_TRY(Q.init(ctx, tot, es));
  std::vector<MixTask> tasks; std::vector<MixTerm> terms; std::vector<GemmProblemHost> p3;
  const int ngB = Bs.t->Vl.nsec();   // SR groups = l sectors of Abar (b)
  for (size_t b = 0; b < o->blks.size(); ++b) {
    const Blk& bk = o->blks[b];
    const int j = bk.m, ja = bk.jl, jb = bk.jr;
    const Group& gB = Bs.t->groups[jb];
    for (int sp = 0; sp < d; ++sp) {
      const int64_t seg = Bs.t->seg[(size_t)sp * ngB + jb];
      if (seg < 0) continue;
      const int jcp = Bs.t->Vr.find((int64_t)Bs.t->Vl.q[jb] + Bs.t->mch[sp]);
      MixTask tk; tk.dst = Q.at(Qoff[b] + seg); tk.ldd = gB.ld; tk.rows = (int32_t)o->Vl.n[ja]; tk.cols = (int32_t)Bs.t->Vr.n[jcp];
      tk.term_begin = (int32_t)terms.size();
      for (int k = 0; k < od; ++k) {
        if (!h->kind[j * od + k]) continue;
        for (int s = 0; s < d; ++s) {
          const cplx c = h->elem(j, k, sp, s);
          if (c == cplx(0.0)) continue;
          const Blk* ab = Ak.t->find(s, ja);
          if (!ab) continue;
          const int jc = ab->jr;
          const int jbra = Gk.t->Vr.find((int64_t)Vc.q[jc] + Gk.t->mch[k]);
          if (jbra < 0) continue;
          if (jbra != jcp) return ctx->fail(MPSK_ESPACE, "transfer_right: MPO entry (%d,%d) violates charge conservation", j, k);
          const int64_t rowoff = Ak.t->seg[(size_t)s * nc + jc], coloff = Gk.t->seg[(size_t)k * nc + jc];
          if (rowoff < 0 || coloff < 0) return ctx->fail(MPSK_EINTERNAL, "transfer_right: missing segment");
          terms.push_back(MixTerm{T.at(Toff[jc] + rowoff * Tld[jc] + coloff), Tld[jc], c.real(), c.imag()});
        }
      }
      tk.term_end = (int32_t)terms.size();
      if (tk.term_end > tk.term_begin) tasks.push_back(tk);
    }
    if (!gB.cols) continue;
    p3.push_back(GemmProblemHost{Bs.t->ptr(gB.off), Q.at(Qoff[b]), o->ptr(bk.off), gB.ld, gB.ld, o->groups[ja].ld, (int32_t)gB.rows, (int32_t)o->Vl.n[ja], (int32_t)gB.cols, 0});
  }
  GemmPlan g1, g3; MixPlan mx;
  MPSK_TRY(gemm_plan_build(ctx, cx, p1, &g1));
  MPSK_TRY(mix_plan_build(ctx, cx, tasks, terms, &mx));
  MPSK_TRY(gemm_plan_build(ctx, cx, p3, &g3));
  MPSK_TRY(gemm_plan_run(ctx, &g1));
  MPSK_TRY(mix_plan_run(ctx, &mx));
  MPSK_TRY(gemm_plan_run(ctx, &g3));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *out = og.release();
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
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *out = xg.release();
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
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
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
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *out = og.release();
  return MPSK_OK;
}

}  // namespace mpsk
