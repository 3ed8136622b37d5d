// extern "C" surface of libmpskit_b200 (see include/mpskit_b200.h for the reference method each entry replaces).
#include "internal.h"
#include <cmath>
#include <dlfcn.h>

using namespace mpsk;

#define CTX(c) (reinterpret_cast<Ctx*>(c))
// every ABI entry that touches a context serialises on it and selects its device (contexts are thread-safe; calls from
// concurrent host threads / Julia tasks on the SAME context run one after the other)
#define ENTER(c) std::lock_guard<std::recursive_mutex> _lock(CTX(c)->mu); cudaSetDevice(CTX(c)->device)
#define OPH(o) (const_cast<LinOp*>(reinterpret_cast<const LinOp*>(o)))
#define TEN(t) (reinterpret_cast<Tensor*>(t))
#define CTEN(t) (reinterpret_cast<const Tensor*>(t))
#define SLICE(s) (reinterpret_cast<const MpoSlice*>(s))

// stage host blocks (arbitrary strides) through one pinned buffer, then scatter with the strided-copy kernel
template <typename T>
static void pack_block(T* dst, const T* src, int64_t nl, int64_t nr, int64_t sl, int64_t sr) {
  for (int64_t i = 0; i < nl; ++i) for (int64_t j = 0; j < nr; ++j) dst[i * nr + j] = src[i * sl + j * sr];
}
template <typename T>
static void unpack_block(T* dst, const T* src, int64_t nl, int64_t nr, int64_t sl, int64_t sr) {
  for (int64_t i = 0; i < nl; ++i) for (int64_t j = 0; j < nr; ++j) dst[i * sl + j * sr] = src[i * nr + j];
}


static void* nccl_lib() {
  static void* lib = nullptr;
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  return lib;
}

extern "C" {

int32_t mpsk_version(void) { return 200; }

int32_t mpsk_create(int32_t device, uint64_t flags, mpsk_ctx** out) {
  if (!out) return MPSK_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return MPSK_ENODEVICE; }
  if (device < 0 || device >= ndev) return MPSK_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return MPSK_ECUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MPSK_ECUDA;
  if (prop.major != 10) return MPSK_ENODEVICE;   // sm_100a only: no other code path exists
  Ctx* c = new Ctx;
  c->device = device; c->flags = flags; c->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return MPSK_ECUDA; }
  c->stream = c->own_stream;

  *out = reinterpret_cast<mpsk_ctx*>(c);
  return MPSK_OK;
}

void mpsk_destroy(mpsk_ctx* ctx) {
  if (!ctx) return;
  Ctx* c = CTX(ctx);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  pool_destroy(c);
  if (c->nccl) {
    typedef int (*pfn_destroy)(void*);
    pfn_destroy f = (pfn_destroy)dlsym(nccl_lib(), "ncclCommDestroy");
    if (f) f(c->nccl);
    c->nccl = nullptr;
  }
  if (c->d_scratch) c->free(c->d_scratch);
  if (c->h_arena) cudaFreeHost(c->h_arena);
  {   // everything still live belongs to handles the caller leaked: release the chunks regardless
    std::vector<void*> live;
    for (auto& kv : c->dev.live) live.push_back(kv.first);
    for (void* q : live) c->dev.free(q);
    c->release_cache();
  }
  if (c->h_scratch) cudaFreeHost(c->h_scratch);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

const char* mpsk_last_error(const mpsk_ctx* ctx) { return ctx ? reinterpret_cast<const Ctx*>(ctx)->err.c_str() : "null context"; }

int32_t mpsk_set_stream(mpsk_ctx* ctx, void* s) { ENTER(ctx);
  Ctx* c = CTX(ctx);
  cudaStreamSynchronize(c->stream);
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return MPSK_OK;
}

int32_t mpsk_synchronize(mpsk_ctx* ctx) { ENTER(ctx); Ctx* c = CTX(ctx); CUDA_TRY(c, cudaStreamSynchronize(c->stream)); return MPSK_OK; }
int64_t mpsk_launch_count(const mpsk_ctx* ctx) { return reinterpret_cast<const Ctx*>(ctx)->launches; }
int64_t mpsk_bytes_in_use(const mpsk_ctx* ctx) { return reinterpret_cast<const Ctx*>(ctx)->bytes_in_use; }

// ------------------------------------------------------------------------------------------ tensors
static int spaces_from_abi(Ctx* c, const mpsk_space* Vl, const mpsk_space* Vr, Space* l, Space* r) {
  MPSK_TRY(space_from_abi(c, Vl, l));
  MPSK_TRY(space_from_abi(c, Vr, r));
  if (l->mod != r->mod) return c->fail(MPSK_ESPACE, "bond spaces belong to different groups");
  return MPSK_OK;
}

int32_t mpsk_block_table(const mpsk_space* Vl, int32_t nm, const int32_t* mcharge, const mpsk_space* Vr, int32_t* nblocks,
                         int32_t* m, int32_t* lsec, int32_t* rsec) {
  Ctx tmp; Space l, r;
  if (spaces_from_abi(&tmp, Vl, Vr, &l, &r) != MPSK_OK) return MPSK_ESPACE;
  int n = 0;
  for (int mm = 0; mm < nm; ++mm)
    for (int jl = 0; jl < l.nsec(); ++jl) {
      int jr = r.find((int64_t)l.q[jl] + mcharge[mm]);
      if (jr < 0) continue;
      if (m) { m[n] = mm; lsec[n] = jl; rsec[n] = jr; }
      ++n;
    }
  if (nblocks) *nblocks = n;
  return MPSK_OK;
}

int32_t mpsk_tensor_zeros(mpsk_ctx* ctx, const mpsk_space* Vl, int32_t nm, const int32_t* mcharge, const mpsk_space* Vr, int32_t dtype,
                          mpsk_tensor** out) { ENTER(ctx);
  Ctx* c = CTX(ctx); Space l, r;
  if (dtype != MPSK_F64 && dtype != MPSK_C128) return c->fail(MPSK_EINVAL, "bad dtype");
  MPSK_TRY(spaces_from_abi(c, Vl, Vr, &l, &r));
  Tensor* t = nullptr;
  MPSK_TRY(tensor_create(c, l, std::vector<int32_t>(mcharge, mcharge + nm), r, dtype, F_KL, true, &t));
  *out = reinterpret_cast<mpsk_tensor*>(t);
  return MPSK_OK;
}

static bool contiguous_rm(int64_t nl, int64_t nr, int64_t sl, int64_t sr) { return sr == 1 && (sl == nr || nl == 1); }
static bool contiguous_cm(int64_t nl, int64_t nr, int64_t sl, int64_t sr) { return sl == 1 && (sr == nl || nr == 1); }

static void* pinned_stage(Ctx* c, size_t bytes) {
  if (c->h_stage_bytes < bytes) {
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr; c->h_stage_bytes = 0;
    size_t want = bytes + bytes / 4;
    if (cudaMallocHost(&c->h_stage, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    c->h_stage_bytes = want;
  }
  return c->h_stage;
}

int32_t mpsk_tensor_upload(mpsk_ctx* ctx, const mpsk_space* Vl, int32_t nm, const int32_t* mcharge, const mpsk_space* Vr, int32_t dtype,
                           const mpsk_blockref* blocks, int32_t nblocks, mpsk_tensor** out) { ENTER(ctx);
  Ctx* c = CTX(ctx);
  mpsk_tensor* th = nullptr;
  MPSK_TRY(mpsk_tensor_zeros(ctx, Vl, nm, mcharge, Vr, dtype, &th));
  std::unique_ptr<Tensor> t(TEN(th));
  if ((int)t->blks.size() != nblocks) return c->fail(MPSK_ESPACE, "upload: %d blocks given, structure has %zu", nblocks, t->blks.size());
  const size_t es = t->esize();
  int64_t total = 0, need_pack = 0;
  for (size_t i = 0; i < t->blks.size(); ++i) {
    const Blk& b = t->blks[i];
    const int64_t nl = t->Vl.n[b.jl], nr = t->Vr.n[b.jr];
    if (!blocks[i].ptr) return c->fail(MPSK_EINVAL, "upload: block %zu has a null pointer", i);
    total += nl * nr;
    if (!contiguous_rm(nl, nr, blocks[i].stride_l, blocks[i].stride_r) && !contiguous_cm(nl, nr, blocks[i].stride_l, blocks[i].stride_r)) need_pack += nl * nr;
  }
  if (total == 0) { *out = reinterpret_cast<mpsk_tensor*>(t.release()); return MPSK_OK; }
  void* hstage = need_pack ? pinned_stage(c, need_pack * es) : nullptr;
  if (need_pack && !hstage) return c->fail(MPSK_ENOMEM, "pinned staging allocation failed");
  void* dstage = c->alloc(total * es);
  if (!dstage) return MPSK_ENOMEM;
  std::vector<CopyTask> tasks;
  int64_t off = 0, hoff = 0;
  for (size_t i = 0; i < t->blks.size(); ++i) {
    const Blk& b = t->blks[i];
    const int64_t nl = t->Vl.n[b.jl], nr = t->Vr.n[b.jr];
    const int64_t sl = blocks[i].stride_l, sr = blocks[i].stride_r;
    char* dptr = (char*)dstage + off * es;
    if (contiguous_rm(nl, nr, sl, sr)) {
      CUDA_TRY(c, cudaMemcpyAsync(dptr, blocks[i].ptr, nl * nr * es, cudaMemcpyHostToDevice, c->stream));
      tasks.push_back(CopyTask{dptr, t->ptr(b.off), nr, 1, b.sl, b.sr, (int32_t)nl, (int32_t)nr, 0, 0, nullptr});
    } else if (contiguous_cm(nl, nr, sl, sr)) {
      CUDA_TRY(c, cudaMemcpyAsync(dptr, blocks[i].ptr, nl * nr * es, cudaMemcpyHostToDevice, c->stream));
      tasks.push_back(CopyTask{dptr, t->ptr(b.off), 1, nl, b.sl, b.sr, (int32_t)nl, (int32_t)nr, 0, 0, nullptr});
    } else {
      char* hp = (char*)hstage + hoff * es;
      if (es == 16) pack_block((cplx*)hp, (const cplx*)blocks[i].ptr, nl, nr, sl, sr);
      else pack_block((double*)hp, (const double*)blocks[i].ptr, nl, nr, sl, sr);
      CUDA_TRY(c, cudaMemcpyAsync(dptr, hp, nl * nr * es, cudaMemcpyHostToDevice, c->stream));
      tasks.push_back(CopyTask{dptr, t->ptr(b.off), nr, 1, b.sl, b.sr, (int32_t)nl, (int32_t)nr, 0, 0, nullptr});
      hoff += nl * nr;
    }
    off += nl * nr;
  }
  int st = run_copy_tasks(c, tasks, es == 16);
  cudaStreamSynchronize(c->stream);
  c->free(dstage);
  if (st != MPSK_OK) return st;
  *out = reinterpret_cast<mpsk_tensor*>(t.release());
  return MPSK_OK;
}

int32_t mpsk_tensor_download(mpsk_ctx* ctx, const mpsk_tensor* th, const mpsk_blockref_mut* blocks, int32_t nblocks) { ENTER(ctx);
  Ctx* c = CTX(ctx); const Tensor* t = CTEN(th);
  if ((int)t->blks.size() != nblocks) return c->fail(MPSK_ESPACE, "download: %d blocks given, structure has %zu", nblocks, t->blks.size());
  const size_t es = t->esize();
  int64_t total = 0, need_pack = 0;
  for (size_t i = 0; i < t->blks.size(); ++i) {
    const Blk& b = t->blks[i];
    const int64_t nl = t->Vl.n[b.jl], nr = t->Vr.n[b.jr];
    total += nl * nr;
    if (!contiguous_rm(nl, nr, blocks[i].stride_l, blocks[i].stride_r) && !contiguous_cm(nl, nr, blocks[i].stride_l, blocks[i].stride_r)) need_pack += nl * nr;
  }
  if (total == 0) return MPSK_OK;
  void* hstage = need_pack ? pinned_stage(c, need_pack * es) : nullptr;
  if (need_pack && !hstage) return c->fail(MPSK_ENOMEM, "pinned staging allocation failed");
  void* dstage = c->alloc(total * es);
  if (!dstage) return MPSK_ENOMEM;
  std::vector<CopyTask> tasks;
  int64_t off = 0;
  for (size_t i = 0; i < t->blks.size(); ++i) {
    const Blk& b = t->blks[i];
    const int64_t nl = t->Vl.n[b.jl], nr = t->Vr.n[b.jr];
    const bool cm = !contiguous_rm(nl, nr, blocks[i].stride_l, blocks[i].stride_r) && contiguous_cm(nl, nr, blocks[i].stride_l, blocks[i].stride_r);
    if (cm) tasks.push_back(CopyTask{t->ptr(b.off), (char*)dstage + off * es, b.sl, b.sr, 1, nl, (int32_t)nl, (int32_t)nr, 0, 0, nullptr});
    else tasks.push_back(CopyTask{t->ptr(b.off), (char*)dstage + off * es, b.sl, b.sr, nr, 1, (int32_t)nl, (int32_t)nr, 0, 0, nullptr});
    off += nl * nr;
  }
  int st = run_copy_tasks(c, tasks, es == 16);
  if (st != MPSK_OK) { c->free(dstage); return st; }
  off = 0; int64_t hoff = 0;
  std::vector<std::pair<size_t, int64_t>> packed;   // (block, offset in hstage)
  for (size_t i = 0; i < t->blks.size(); ++i) {
    const Blk& b = t->blks[i];
    const int64_t nl = t->Vl.n[b.jl], nr = t->Vr.n[b.jr];
    const int64_t sl = blocks[i].stride_l, sr = blocks[i].stride_r;
    const char* dptr = (const char*)dstage + off * es;
    cudaError_t e;
    if (contiguous_rm(nl, nr, sl, sr) || contiguous_cm(nl, nr, sl, sr)) e = cudaMemcpyAsync(blocks[i].ptr, dptr, nl * nr * es, cudaMemcpyDeviceToHost, c->stream);
    else { e = cudaMemcpyAsync((char*)hstage + hoff * es, dptr, nl * nr * es, cudaMemcpyDeviceToHost, c->stream); packed.push_back({i, hoff}); hoff += nl * nr; }
    if (e != cudaSuccess) { c->free(dstage); return c->fail(MPSK_ECUDA, "D2H copy failed: %s", cudaGetErrorString(e)); }
    off += nl * nr;
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  c->free(dstage);
  if (e != cudaSuccess) return c->fail(MPSK_ECUDA, "D2H copy failed: %s", cudaGetErrorString(e));
  for (auto& pk : packed) {
    const Blk& b = t->blks[pk.first];
    const int64_t nl = t->Vl.n[b.jl], nr = t->Vr.n[b.jr];
    if (es == 16) unpack_block((cplx*)blocks[pk.first].ptr, (const cplx*)hstage + pk.second, nl, nr, blocks[pk.first].stride_l, blocks[pk.first].stride_r);
    else unpack_block((double*)blocks[pk.first].ptr, (const double*)hstage + pk.second, nl, nr, blocks[pk.first].stride_l, blocks[pk.first].stride_r);
  }
  return MPSK_OK;
}

int32_t mpsk_tensor_stack_labels(mpsk_ctx* ctx, const mpsk_tensor* const* ts, int32_t n, mpsk_tensor** out) { ENTER(ctx);
  std::vector<const Tensor*> v;
  for (int i = 0; i < n; ++i) v.push_back(CTEN(ts[i]));
  Tensor* o = nullptr;
  MPSK_TRY(tensor_stack_labels(CTX(ctx), v, &o));
  *out = reinterpret_cast<mpsk_tensor*>(o);
  return MPSK_OK;
}
int32_t mpsk_tensor_select_labels(mpsk_ctx* ctx, const mpsk_tensor* t, const int32_t* labels, int32_t n, mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(tensor_select_labels(CTX(ctx), CTEN(t), labels, n, &o));
  *out = reinterpret_cast<mpsk_tensor*>(o);
  return MPSK_OK;
}
int32_t mpsk_tensor_adjoint(mpsk_ctx* ctx, const mpsk_tensor* t, mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(tensor_adjoint(CTX(ctx), CTEN(t), &o));
  *out = reinterpret_cast<mpsk_tensor*>(o);
  return MPSK_OK;
}

int32_t mpsk_tensor_fill_random(mpsk_ctx* ctx, mpsk_tensor* t, uint64_t seed) { ENTER(ctx); return tensor_fill_random(CTX(ctx), TEN(t), seed); }

int32_t mpsk_tensor_copy(mpsk_ctx* ctx, const mpsk_tensor* th, mpsk_tensor** out) { ENTER(ctx);
  Ctx* c = CTX(ctx); const Tensor* t = CTEN(th);
  Tensor* n = nullptr;
  MPSK_TRY(tensor_repack(c, t, F_KL, false, &n));
  *out = reinterpret_cast<mpsk_tensor*>(n);
  return MPSK_OK;
}

void mpsk_tensor_free(mpsk_ctx* ctx, mpsk_tensor* t) { if (!t) return; if (ctx) { ENTER(ctx); delete TEN(t); } else delete TEN(t); }

int32_t mpsk_tensor_structure(const mpsk_tensor* th, int32_t* nsec_l, int32_t* nsec_r, int32_t* nm, int32_t* dtype) {
  const Tensor* t = CTEN(th);
  if (nsec_l) *nsec_l = t->Vl.nsec();
  if (nsec_r) *nsec_r = t->Vr.nsec();
  if (nm) *nm = t->nm();
  if (dtype) *dtype = t->dtype;
  return MPSK_OK;
}

int32_t mpsk_tensor_spaces(const mpsk_tensor* th, int32_t* ql, int64_t* nl, int32_t* qr, int64_t* nr, int32_t* mch) {
  const Tensor* t = CTEN(th);
  for (int i = 0; i < t->Vl.nsec(); ++i) { if (ql) ql[i] = t->Vl.q[i]; if (nl) nl[i] = t->Vl.n[i]; }
  for (int i = 0; i < t->Vr.nsec(); ++i) { if (qr) qr[i] = t->Vr.q[i]; if (nr) nr[i] = t->Vr.n[i]; }
  if (mch) for (int i = 0; i < t->nm(); ++i) mch[i] = t->mch[i];
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ vector interface
static int check_pair(Ctx* c, const Tensor* x, const Tensor* y) {
  if (!x->same_structure(*y)) return c->fail(MPSK_ESPACE, "tensors have different structure");
  if (x->form != y->form) return c->fail(MPSK_EINTERNAL, "tensors are stored in different forms");
  return MPSK_OK;
}

int32_t mpsk_dot(mpsk_ctx* ctx, const mpsk_tensor* xh, const mpsk_tensor* yh, double* re, double* im) { ENTER(ctx);
  Ctx* c = CTX(ctx); const Tensor* x = CTEN(xh); const Tensor* y = CTEN(yh);
  MPSK_TRY(check_pair(c, x, y));
  double* d = (double*)c->alloc(4 * sizeof(double));
  const void* V[1] = {x->d};
  int st = vec_multidot(c, x->dtype == MPSK_C128, 1, V, y->d, x->nelem, d);
  double h[2] = {0, 0};
  if (st == MPSK_OK) { cudaMemcpyAsync(h, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream); cudaStreamSynchronize(c->stream); }
  c->free(d);
  if (re) *re = h[0];
  if (im) *im = h[1];
  return st;
}

int32_t mpsk_norm(mpsk_ctx* ctx, const mpsk_tensor* x, double* out) { ENTER(ctx);
  double re = 0, im = 0;
  MPSK_TRY(mpsk_dot(ctx, x, x, &re, &im));
  *out = std::sqrt(re);
  return MPSK_OK;
}

int32_t mpsk_axpy(mpsk_ctx* ctx, double ar, double ai, const mpsk_tensor* xh, mpsk_tensor* yh) { ENTER(ctx);
  Ctx* c = CTX(ctx); const Tensor* x = CTEN(xh); Tensor* y = TEN(yh);
  MPSK_TRY(check_pair(c, x, y));
  if (x->dtype == MPSK_F64 && ai != 0.0) return c->fail(MPSK_EINVAL, "complex coefficient on a real tensor");
  return vec_axpy(c, x->dtype == MPSK_C128, cplx(ar, ai), x->d, y->d, x->nelem);
}

int32_t mpsk_scale(mpsk_ctx* ctx, double ar, double ai, mpsk_tensor* xh) { ENTER(ctx);
  Ctx* c = CTX(ctx); Tensor* x = TEN(xh);
  if (x->dtype == MPSK_F64 && ai != 0.0) return c->fail(MPSK_EINVAL, "complex coefficient on a real tensor");
  return vec_scale(c, x->dtype == MPSK_C128, x->d, x->nelem, cplx(ar, ai));
}

// ------------------------------------------------------------------------------------------ MPO slices
int32_t mpsk_mposlice_create_zn(mpsk_ctx* ctx, int32_t odim, const int32_t* cl, const int32_t* cr, int32_t d, const int32_t* pch,
                                const int32_t* kind, const double* scal, const double* ops, int32_t modulus, mpsk_mposlice** out) { ENTER(ctx);
  Ctx* c = CTX(ctx);
  if (odim <= 0 || d <= 0 || modulus < 0) return c->fail(MPSK_EINVAL, "bad slice dimensions");
  auto nrm = [&](int64_t q) -> int64_t { if (!modulus) return q; int64_t r = q % modulus; return r < 0 ? r + modulus : r; };
  std::unique_ptr<MpoSlice> s(new MpoSlice);
  s->odim = odim; s->d = d;
  s->cl.assign(cl, cl + odim); s->cr.assign(cr, cr + odim); s->pch.assign(pch, pch + d);
  s->kind.assign(kind, kind + odim * odim);
  s->scal.assign((size_t)odim * odim, cplx(0.0)); s->ops.resize((size_t)odim * odim);
  size_t nop = 0;
  for (int i = 0; i < odim; ++i) for (int j = 0; j < odim; ++j) {
    const int k = kind[i * odim + j];
    if (k < 0 || k > 2) return c->fail(MPSK_EINVAL, "bad entry kind");
    if (k == 1) {
      s->scal[i * odim + j] = cplx(scal[2 * (i * odim + j)], scal[2 * (i * odim + j) + 1]);
      if (nrm(cl[i]) != nrm(cr[j])) return c->fail(MPSK_ESPACE, "scalar entry (%d,%d) connects levels of different charge", i, j);
    } else if (k == 2) {
      std::vector<cplx>& o = s->ops[i * odim + j];
      o.resize((size_t)d * d);
      for (int e = 0; e < d * d; ++e) o[e] = cplx(ops[2 * (nop * d * d + e)], ops[2 * (nop * d * d + e) + 1]);
      for (int a = 0; a < d; ++a) for (int b = 0; b < d; ++b)
        if (o[a * d + b] != cplx(0.0) && nrm(cr[j]) != nrm((int64_t)cl[i] + pch[a] - pch[b]))
          return c->fail(MPSK_ESPACE, "operator entry (%d,%d) element (%d,%d) violates charge conservation", i, j, a, b);
      ++nop;
    }
  }
  *out = reinterpret_cast<mpsk_mposlice*>(s.release());
  return MPSK_OK;
}

int32_t mpsk_mposlice_create(mpsk_ctx* ctx, int32_t odim, const int32_t* cl, const int32_t* cr, int32_t d, const int32_t* pch,
                             const int32_t* kind, const double* scal, const double* ops, mpsk_mposlice** out) {
  return mpsk_mposlice_create_zn(ctx, odim, cl, cr, d, pch, kind, scal, ops, 0, out);
}

void mpsk_mposlice_free(mpsk_ctx*, mpsk_mposlice* s) { delete reinterpret_cast<MpoSlice*>(const_cast<mpsk_mposlice*>(s)); }

int32_t mpsk_mposlice_keys(const mpsk_mposlice* sh, int32_t* count, int32_t* ij) {
  const MpoSlice* s = SLICE(sh);
  int n = 0;
  for (int j = 0; j < s->odim; ++j) for (int i = 0; i < s->odim; ++i) {   // product(1:odim,1:odim): first index fastest
    if (!s->kind[i * s->odim + j]) continue;
    if (ij) { ij[2 * n] = i; ij[2 * n + 1] = j; }
    ++n;
  }
  if (count) *count = n;
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ operators
int32_t mpsk_op_create(mpsk_ctx* ctx, int32_t kind, const mpsk_mposlice* h1, const mpsk_mposlice* h2, const mpsk_tensor* GL,
                       const mpsk_tensor* GR, const mpsk_tensor* x_like, mpsk_op** out) { ENTER(ctx);
  HeffOp* op = nullptr;
  MPSK_TRY(heff_create(CTX(ctx), kind, SLICE(h1), SLICE(h2), CTEN(GL), CTEN(GR), CTEN(x_like), &op));
  *out = reinterpret_cast<mpsk_op*>(static_cast<LinOp*>(op));
  return MPSK_OK;
}
void mpsk_op_free(mpsk_ctx* ctx, mpsk_op* op) { if (!op) return; if (ctx) { ENTER(ctx); delete reinterpret_cast<LinOp*>(op); } else delete reinterpret_cast<LinOp*>(op); }
double mpsk_op_flops(const mpsk_op* op) { return reinterpret_cast<const LinOp*>(op)->flops(); }

// measurement hook for bench.py: average device time (ms) of the two grouped-GEMM launches of one matvec,
// CUDA events on the launching stream, inputs as left by the last apply.
int32_t mpsk_op_time_gemms(mpsk_ctx* ctx, const mpsk_op* oph, int32_t reps, double* ms_out) { ENTER(ctx);
  Ctx* c = CTX(ctx); const HeffOp* op = dynamic_cast<const HeffOp*>(reinterpret_cast<const LinOp*>(oph));
  if (!op) return c->fail(MPSK_EINVAL, "not an effective-Hamiltonian operator");
  cudaEvent_t e0, e1;
  CUDA_TRY(c, cudaEventCreate(&e0)); CUDA_TRY(c, cudaEventCreate(&e1));
  double tot = 0;
  for (int r = 0; r < reps; ++r) {
    float a = 0, b = 0;
    CUDA_TRY(c, cudaEventRecord(e0, c->stream));
    MPSK_TRY(gemm_plan_run(c, op->gemmA.get()));
    CUDA_TRY(c, cudaEventRecord(e1, c->stream));
    CUDA_TRY(c, cudaEventSynchronize(e1));
    CUDA_TRY(c, cudaEventElapsedTime(&a, e0, e1));
    MPSK_TRY(mix_plan_run(c, op->mix.get()));
    CUDA_TRY(c, cudaEventRecord(e0, c->stream));
    MPSK_TRY(gemm_plan_run(c, op->gemmC.get()));
    CUDA_TRY(c, cudaEventRecord(e1, c->stream));
    CUDA_TRY(c, cudaEventSynchronize(e1));
    CUDA_TRY(c, cudaEventElapsedTime(&b, e0, e1));
    tot += a + b;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_out = tot / reps;
  return MPSK_OK;
}

// measurement hook: average device time (ms) and algorithmic bytes of the mix launch of one matvec
int32_t mpsk_op_time_mix(mpsk_ctx* ctx, const mpsk_op* oph, int32_t reps, double* ms_out, double* bytes_out) { ENTER(ctx);
  Ctx* c = CTX(ctx); const HeffOp* op = dynamic_cast<const HeffOp*>(reinterpret_cast<const LinOp*>(oph));
  if (!op) return c->fail(MPSK_EINVAL, "not an effective-Hamiltonian operator");
  cudaEvent_t e0, e1;
  CUDA_TRY(c, cudaEventCreate(&e0)); CUDA_TRY(c, cudaEventCreate(&e1));
  double tot = 0;
  for (int r = 0; r < reps; ++r) {
    float a = 0;
    CUDA_TRY(c, cudaEventRecord(e0, c->stream));
    MPSK_TRY(mix_plan_run(c, op->mix.get()));
    CUDA_TRY(c, cudaEventRecord(e1, c->stream));
    CUDA_TRY(c, cudaEventSynchronize(e1));
    CUDA_TRY(c, cudaEventElapsedTime(&a, e0, e1));
    tot += a;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (ms_out) *ms_out = tot / std::max(reps, 1);
  if (bytes_out) *bytes_out = op->mix->bytes;
  return MPSK_OK;
}

static int as_kl(Ctx* c, const Tensor* t) { return t->form == F_KL ? MPSK_OK : c->fail(MPSK_EINTERNAL, "tensor is not in vector (KL) form"); }

int32_t mpsk_op_apply(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* x, mpsk_tensor* y) { ENTER(ctx);
  MPSK_TRY(as_kl(CTX(ctx), CTEN(x))); MPSK_TRY(as_kl(CTX(ctx), CTEN(y)));
  return OPH(op)->apply(CTEN(x), TEN(y));
}

int32_t mpsk_eigsolve(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* x0, int32_t which, int32_t hermitian, int32_t krylovdim,
                      int32_t maxiter, double tol, double* val_re, double* val_im, mpsk_tensor* vec, mpsk_info* info) { ENTER(ctx);
  Ctx* c = CTX(ctx);
  MPSK_TRY(as_kl(c, CTEN(x0))); MPSK_TRY(as_kl(c, CTEN(vec)));
  cplx v = 0;
  LinOp* h = OPH(op);
  int st = hermitian ? eigsolve_lanczos(c, h, CTEN(x0), which, krylovdim, maxiter, tol, &v, TEN(vec), info)
                     : eigsolve_arnoldi(c, h, CTEN(x0), which, krylovdim, maxiter, tol, &v, TEN(vec), info);
  if (val_re) *val_re = v.real();
  if (val_im) *val_im = v.imag();
  return st;
}

int32_t mpsk_eigsolve_values(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* x0, int32_t which, int32_t howmany, int32_t krylovdim,
                             int32_t maxiter, double tol, double* vals, int32_t* nvals, int32_t* nconverged, mpsk_info* info) { ENTER(ctx);
  Ctx* c = CTX(ctx);
  if (!vals || howmany < 1) return c->fail(MPSK_EINVAL, "eigsolve_values: bad arguments");
  MPSK_TRY(as_kl(c, CTEN(x0)));
  std::vector<cplx> v((size_t)howmany, cplx(0.0));
  int nc = 0, got = 0;
  mpsk_info li; li.converged = 0; li.numiter = 0; li.numops = 0; li.normres = 0;
  MPSK_TRY(eigsolve_arnoldi(c, OPH(op), CTEN(x0), which, krylovdim, maxiter, tol, nullptr, nullptr, &li, howmany, v.data(), &got, &nc));
  for (int j = 0; j < howmany; ++j) { vals[2 * j] = j < got ? v[j].real() : 0.0; vals[2 * j + 1] = j < got ? v[j].imag() : 0.0; }
  if (nvals) *nvals = got;
  if (nconverged) *nconverged = nc;
  if (info) *info = li;
  return MPSK_OK;
}

int32_t mpsk_expv(mpsk_ctx* ctx, const mpsk_op* op, double tr, double ti, const mpsk_tensor* y0, int32_t krylovdim, int32_t maxiter,
                  double tol, mpsk_tensor* y, mpsk_info* info) { ENTER(ctx);
  Ctx* c = CTX(ctx);
  MPSK_TRY(as_kl(c, CTEN(y0))); MPSK_TRY(as_kl(c, CTEN(y)));
  return expv_lanczos(c, OPH(op), cplx(tr, ti), CTEN(y0), krylovdim, maxiter, tol, TEN(y), info);
}

int32_t mpsk_linsolve(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* b, const mpsk_tensor* x0, double a0_re, double a0_im,
                      double a1_re, double a1_im, int32_t krylovdim, int32_t maxiter, double tol, mpsk_tensor* x, mpsk_info* info) { ENTER(ctx);
  Ctx* c = CTX(ctx);
  MPSK_TRY(as_kl(c, CTEN(b))); MPSK_TRY(as_kl(c, CTEN(x0))); MPSK_TRY(as_kl(c, CTEN(x)));
  return linsolve_gmres(c, OPH(op), CTEN(b), CTEN(x0), cplx(a0_re, a0_im), cplx(a1_re, a1_im), krylovdim, maxiter, tol, TEN(x), info);
}

// transfer matrix over a unit cell as an operator handle (transfermatrix.jl:4-90): usable with mpsk_op_apply, mpsk_eigsolve
// (hermitian = 0: Arnoldi) and mpsk_linsolve
int32_t mpsk_tmop_create(mpsk_ctx* ctx, int32_t left, int32_t nsites, const mpsk_mposlice* const* slices, const mpsk_tensor* const* A,
                         const mpsk_tensor* const* Abar, const mpsk_tensor* v_like, const mpsk_tensor* lvec_adj, const mpsk_tensor* rvec,
                         mpsk_op** out) { ENTER(ctx);
  Ctx* c = CTX(ctx);
  if (nsites <= 0 || !A || !Abar || !v_like) return c->fail(MPSK_EINVAL, "tmop: bad arguments");
  if ((lvec_adj == nullptr) != (rvec == nullptr)) return c->fail(MPSK_EINVAL, "tmop: regularisation needs both vectors");
  const Tensor* vl = CTEN(v_like);
  std::unique_ptr<TmOp> op(new TmOp);
  op->ctx = c; op->dtype = vl->dtype; op->left = left != 0;
  { Tensor* t = nullptr; MPSK_TRY(tensor_create(c, vl->Vl, vl->mch, vl->Vr, vl->dtype, F_KL, true, &t)); op->vlike.reset(t); }
  std::unique_ptr<Tensor> cur_like;     // structure of the vector entering the next site
  const Tensor* cur = op->vlike.get();
  for (int s = 0; s < nsites; ++s) {
    const int site = left ? s : nsites - 1 - s;                      // v * tm sweeps left -> right, tm * v right -> left
    const Tensor* a = CTEN(A[site]); const Tensor* ab = CTEN(Abar[site]);
    const MpoSlice* h = slices ? SLICE(slices[site]) : nullptr;
    MpoSlice unit;
    if (!h) {     // MPO-free transfer (transfer.jl:18-45): one level, identity on the physical labels
      if (cur->nm() != 1) return c->fail(MPSK_ESPACE, "tmop: MPO-free transfer needs single-level vectors");
      unit.odim = 1; unit.d = a->nm(); unit.cl = {-cur->mch[0] * (left ? 1 : -1)}; unit.cr = unit.cl;
      unit.pch.assign(a->mch.begin(), a->mch.end()); unit.kind = {1}; unit.scal = {cplx(1.0)}; unit.ops.resize(1);
      h = &unit;
    }
    TransferPlan* p = nullptr;
    MPSK_TRY(transfer_plan_create(c, left != 0, h, a, ab, cur, /*sharded=*/false, &p));
    op->sites.emplace_back(p);
    op->flops_ += p->flops;
    cur = p->result();
  }
  if (!cur->same_structure(*op->vlike)) return c->fail(MPSK_ESPACE, "tmop: the unit cell does not map the vector space onto itself");
  if (lvec_adj) {
    const Tensor* la = CTEN(lvec_adj); const Tensor* rv = CTEN(rvec);
    if (!la->same_structure(*op->vlike) || !rv->same_structure(*op->vlike) || la->form != F_KL || rv->form != F_KL)
      return c->fail(MPSK_ESPACE, "tmop: regularisation vectors do not match the vector space");
    Tensor* t = nullptr;
    MPSK_TRY(tensor_repack(c, la, F_KL, false, &t)); op->ladj.reset(t);
    MPSK_TRY(tensor_repack(c, rv, F_KL, false, &t)); op->rvec.reset(t);
    op->d_coef = (double*)c->alloc(4 * sizeof(double));
    if (!op->d_coef) return MPSK_ENOMEM;
  }
  *out = reinterpret_cast<mpsk_op*>(static_cast<LinOp*>(op.release()));
  return MPSK_OK;
}

// direct sum of operators with a cyclic shift (multi-row derivatives of the statmech VUMPS, derivatives.jl:107-110,163-166,195-197)
int32_t mpsk_multiop_create(mpsk_ctx* ctx, int32_t n, const mpsk_op* const* ops, int32_t shift, mpsk_op** out) { ENTER(ctx);
  std::vector<LinOp*> v;
  for (int i = 0; i < n; ++i) v.push_back(OPH(ops[i]));
  MultiOp* m = nullptr;
  MPSK_TRY(multiop_create(CTX(ctx), v, shift, &m));
  *out = reinterpret_cast<mpsk_op*>(static_cast<LinOp*>(m));
  return MPSK_OK;
}

// host-only self-test of the device allocator's free-list logic (devalloc.h) with malloc as the backend: random
// allocate / release traffic, invariants checked after every operation; returns 0 on success
static void* host_chunk_alloc(size_t b) { return malloc(b); }
static void host_chunk_free(void* p) { free(p); }
int32_t mpsk_host_alloc_selftest(uint64_t seed, int32_t nops) {
  BlockAllocator a;
  a.chunk_alloc = host_chunk_alloc; a.chunk_free = host_chunk_free; a.chunk_bytes = (size_t)1 << 20;
  std::vector<std::pair<char*, size_t>> livev;
  uint64_t x = seed * 2862933555777941757ull + 3037000493ull;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  for (int i = 0; i < nops; ++i) {
    const bool do_alloc = livev.empty() || (rnd() % 100) < 55;
    if (do_alloc) {
      size_t sz = (size_t)(rnd() % (((rnd() % 8) == 0) ? (3u << 20) : (64u << 10))) + 1;
      char* p = (char*)a.alloc(sz);
      if (!p) return -10;
      for (auto& lv : livev) if (p < lv.first + lv.second && lv.first < p + sz) return -11;      // overlap with a live block
      memset(p, (int)(i & 0xff), sz);
      livev.push_back({p, sz});
    } else {
      const size_t k = rnd() % livev.size();
      a.free(livev[k].first);
      livev[k] = livev.back(); livev.pop_back();
    }
    if (!a.check()) return -12;
  }
  for (auto& lv : livev) a.free(lv.first);
  if (!a.check() || a.bytes_live != 0) return -13;
  if (a.free_by_addr.size() != a.chunks.size()) return -14;                                      // everything coalesced back
  a.trim();
  return a.chunks.empty() && a.bytes_reserved == 0 ? 0 : -15;
}

// host-only numerical helper exposed for the CPU test-suite: complex Schur decomposition (+ optional reordering) used by
// the Krylov-Schur eigensolver.  H: n x n column-major complex (interleaved), overwritten by T; Z: n x n Schur vectors.
int32_t mpsk_host_schur(int32_t n, double* H, double* Z, int32_t move_from, int32_t move_to) {
  if (n <= 0) return MPSK_EINVAL;
  std::vector<cplx> h((size_t)n * n), z;
  for (size_t i = 0; i < h.size(); ++i) h[i] = cplx(H[2 * i], H[2 * i + 1]);
  int st = schur_complex(n, h, z);
  if (st != MPSK_OK) return st;
  if (move_from >= 0 && move_to >= 0 && move_to <= move_from && move_from < n) schur_move(n, h, z, move_from, move_to);
  for (size_t i = 0; i < h.size(); ++i) { H[2 * i] = h[i].real(); H[2 * i + 1] = h[i].imag(); Z[2 * i] = z[i].real(); Z[2 * i + 1] = z[i].imag(); }
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ transfers / local algebra
static int to_kl(Ctx* c, Tensor* t, mpsk_tensor** out) {
  // every tensor leaving the library is in KL (vector) form
  if (t->form == F_KL) { *out = reinterpret_cast<mpsk_tensor*>(t); return MPSK_OK; }
  Tensor* n = nullptr;
  int st = tensor_repack(c, t, F_KL, false, &n);
  delete t;
  if (st != MPSK_OK) return st;
  *out = reinterpret_cast<mpsk_tensor*>(n);
  return MPSK_OK;
}

int32_t mpsk_transfer_left(mpsk_ctx* ctx, const mpsk_mposlice* h, const mpsk_tensor* A, const mpsk_tensor* Abar, const mpsk_tensor* GL,
                           mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(transfer_left(CTX(ctx), SLICE(h), CTEN(A), CTEN(Abar), CTEN(GL), &o));
  return to_kl(CTX(ctx), o, out);
}
int32_t mpsk_transfer_right(mpsk_ctx* ctx, const mpsk_mposlice* h, const mpsk_tensor* A, const mpsk_tensor* Abar, const mpsk_tensor* GR,
                            mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(transfer_right(CTX(ctx), SLICE(h), CTEN(A), CTEN(Abar), CTEN(GR), &o));
  return to_kl(CTX(ctx), o, out);
}
int32_t mpsk_contract_ac2(mpsk_ctx* ctx, const mpsk_tensor* A1, const mpsk_tensor* A2, mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(contract_ac2(CTX(ctx), CTEN(A1), CTEN(A2), &o));
  return to_kl(CTX(ctx), o, out);
}
int32_t mpsk_mul_bond_right(mpsk_ctx* ctx, const mpsk_tensor* A, const mpsk_tensor* C, mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(mul_bond_right(CTX(ctx), CTEN(A), CTEN(C), &o));
  return to_kl(CTX(ctx), o, out);
}
int32_t mpsk_mul_bond_left(mpsk_ctx* ctx, const mpsk_tensor* C, const mpsk_tensor* A, mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(mul_bond_left(CTX(ctx), CTEN(C), CTEN(A), &o));
  return to_kl(CTX(ctx), o, out);
}

int32_t mpsk_tsvd(mpsk_ctx* ctx, const mpsk_tensor* ac2, int32_t d1, const int32_t* p1_charge, mpsk_trunc trunc, mpsk_tensor** U,
                  mpsk_tensor** S, mpsk_tensor** Vh, double* eps) { ENTER(ctx);
  Tensor *u = nullptr, *s = nullptr, *v = nullptr;
  MPSK_TRY(tsvd(CTX(ctx), CTEN(ac2), d1, p1_charge, trunc, &u, &s, &v, eps));
  *U = reinterpret_cast<mpsk_tensor*>(u); *S = reinterpret_cast<mpsk_tensor*>(s); *Vh = reinterpret_cast<mpsk_tensor*>(v);
  return MPSK_OK;
}

int32_t mpsk_tsvd_spectrum(mpsk_ctx* ctx, int32_t* nsec, int32_t* charge, int32_t* count, int32_t* kept, double* values) { ENTER(ctx);
  const TsvdSpectrum& sp = CTX(ctx)->last_spectrum;
  if (nsec) *nsec = (int32_t)sp.charge.size();
  for (size_t i = 0; i < sp.charge.size(); ++i) { if (charge) charge[i] = sp.charge[i]; if (count) count[i] = sp.count[i]; if (kept) kept[i] = sp.kept[i]; }
  if (values) std::copy(sp.values.begin(), sp.values.end(), values);
  return MPSK_OK;
}

int32_t mpsk_leftnull(mpsk_ctx* ctx, const mpsk_tensor* A, mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(nullspace(CTX(ctx), CTEN(A), true, &o));
  *out = reinterpret_cast<mpsk_tensor*>(o);
  return MPSK_OK;
}
int32_t mpsk_rightnull(mpsk_ctx* ctx, const mpsk_tensor* A, mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(nullspace(CTX(ctx), CTEN(A), false, &o));
  *out = reinterpret_cast<mpsk_tensor*>(o);
  return MPSK_OK;
}
int32_t mpsk_contract_adjoint_left(mpsk_ctx* ctx, const mpsk_tensor* N, const mpsk_tensor* x2, mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(contract_adjoint_left(CTX(ctx), CTEN(N), CTEN(x2), &o));
  *out = reinterpret_cast<mpsk_tensor*>(o);
  return MPSK_OK;
}
int32_t mpsk_tensor_cat(mpsk_ctx* ctx, const mpsk_tensor* A, const mpsk_tensor* B, int32_t axis, mpsk_tensor** out) { ENTER(ctx);
  if (axis != 0 && axis != 1) return CTX(ctx)->fail(MPSK_EINVAL, "axis must be 0 or 1");
  Tensor* o = nullptr;
  MPSK_TRY(tensor_cat(CTX(ctx), CTEN(A), CTEN(B), axis, &o));
  *out = reinterpret_cast<mpsk_tensor*>(o);
  return MPSK_OK;
}

int32_t mpsk_bond_inverse(mpsk_ctx* ctx, const mpsk_tensor* C, mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(bond_inverse(CTX(ctx), CTEN(C), &o));
  *out = reinterpret_cast<mpsk_tensor*>(o);
  return MPSK_OK;
}

int32_t mpsk_qrpos(mpsk_ctx* ctx, const mpsk_tensor* A, mpsk_tensor** Q, mpsk_tensor** R) { ENTER(ctx);
  Tensor *q = nullptr, *r = nullptr;
  MPSK_TRY(qrpos(CTX(ctx), CTEN(A), true, &q, &r));
  *Q = reinterpret_cast<mpsk_tensor*>(q); *R = reinterpret_cast<mpsk_tensor*>(r);
  return MPSK_OK;
}
int32_t mpsk_lqpos(mpsk_ctx* ctx, const mpsk_tensor* A, mpsk_tensor** L, mpsk_tensor** Q) { ENTER(ctx);
  Tensor *q = nullptr, *l = nullptr;
  MPSK_TRY(qrpos(CTX(ctx), CTEN(A), false, &q, &l));
  *Q = reinterpret_cast<mpsk_tensor*>(q); *L = reinterpret_cast<mpsk_tensor*>(l);
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ NCCL (loaded lazily; one process per GPU)
typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*pfn_ncclGetUniqueId)(nccl_uid_t*);
typedef int (*pfn_ncclCommInitRank)(void**, int, nccl_uid_t, int);
typedef int (*pfn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);

int32_t mpsk_comm_unique_id(uint8_t out128[128]) {
  void* lib = nccl_lib();
  if (!lib) return MPSK_ECUDA;
  pfn_ncclGetUniqueId f = (pfn_ncclGetUniqueId)dlsym(lib, "ncclGetUniqueId");
  nccl_uid_t id;
  if (!f || f(&id) != 0) return MPSK_ECUDA;
  memcpy(out128, id.internal, 128);
  return MPSK_OK;
}

int32_t mpsk_comm_init(mpsk_ctx* ctx, const uint8_t id128[128], int32_t rank, int32_t nranks) { ENTER(ctx);
  Ctx* c = CTX(ctx);
  void* lib = nccl_lib();
  if (!lib) return c->fail(MPSK_ECUDA, "libnccl.so.2 not found: %s", dlerror());
  pfn_ncclCommInitRank f = (pfn_ncclCommInitRank)dlsym(lib, "ncclCommInitRank");
  if (!f) return c->fail(MPSK_ECUDA, "ncclCommInitRank not found");
  nccl_uid_t id; memcpy(id.internal, id128, 128);
  void* comm = nullptr;
  int r = f(&comm, nranks, id, rank);
  if (r != 0) return c->fail(MPSK_ECUDA, "ncclCommInitRank failed (%d)", r);
  c->nccl = comm; c->rank = rank; c->nranks = nranks;
  return MPSK_OK;
}

int32_t mpsk_tensor_row_panel(mpsk_ctx* ctx, const mpsk_tensor* t, int32_t rank, int32_t nranks, mpsk_tensor** out) { ENTER(ctx);
  Tensor* o = nullptr;
  MPSK_TRY(tensor_row_panel(CTX(ctx), CTEN(t), rank, nranks, &o));
  *out = reinterpret_cast<mpsk_tensor*>(o);
  return MPSK_OK;
}

int32_t mpsk_tensor_scatter_panel(mpsk_ctx* ctx, const mpsk_tensor* panel, int32_t rank, int32_t nranks, mpsk_tensor* full) { ENTER(ctx);
  Ctx* c = CTX(ctx); const Tensor* p = CTEN(panel); Tensor* f = TEN(full);
  if (p->form != F_KL || p->nelem != panel_nelem(f, rank, nranks) || p->dtype != f->dtype) return c->fail(MPSK_ESPACE, "tensor is not the row panel of that rank");
  return tensor_scatter_panel(c, p->d, rank, nranks, f, true);
}

static HeffOp* as_heff(const mpsk_op* op) { return dynamic_cast<HeffOp*>(OPH(op)); }

int32_t mpsk_op_apply_local(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* x, mpsk_tensor* y_panel) { ENTER(ctx);
  HeffOp* h = as_heff(op);
  if (!h) return CTX(ctx)->fail(MPSK_EINVAL, "not an effective-Hamiltonian operator");
  return heff_apply_local(CTX(ctx), h, CTEN(x), TEN(y_panel));
}

int32_t mpsk_op_apply_panel(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* x_panel, mpsk_tensor* y_panel) { ENTER(ctx);
  HeffOp* h = as_heff(op);
  if (!h) return CTX(ctx)->fail(MPSK_EINVAL, "not an effective-Hamiltonian operator");
  return heff_apply_panel(CTX(ctx), h, CTEN(x_panel), TEN(y_panel));
}

// ---- streamed matvecs on host buffers: H2D copy, matvec and D2H copy of successive independent vectors overlap -------------
// (copy engines on two side streams, events between them and the context stream; `depth` slots of device buffers)
namespace {
struct MatvecStream {
  Ctx* ctx = nullptr; HeffOp* op = nullptr; int depth = 2; int layout = 0; int64_t submitted = 0;
  struct Slot {
    Tensor* x = nullptr; Tensor* y = nullptr; void* sx = nullptr; void* sy = nullptr;
    CopyPlan in, out; cudaEvent_t up = nullptr, done = nullptr, down = nullptr; bool busy = false;
  };
  std::vector<Slot> slots;
  cudaStream_t h2d = nullptr, d2h = nullptr;
  void drain() { for (auto& k : slots) if (k.busy) { cudaEventSynchronize(k.down); k.busy = false; } }
  ~MatvecStream() {
    drain();
    if (ctx) cudaStreamSynchronize(ctx->stream);
    for (auto& k : slots) {
      delete k.x; delete k.y;
      if (ctx) { ctx->free(k.sx); ctx->free(k.sy); }
      if (k.up) cudaEventDestroy(k.up);
      if (k.done) cudaEventDestroy(k.done);
      if (k.down) cudaEventDestroy(k.down);
    }
    if (h2d) cudaStreamDestroy(h2d);
    if (d2h) cudaStreamDestroy(d2h);
  }
};
// contiguous staging image of a tensor's blocks (ABI order), row-major (layout 0) or column-major (layout 1) per block
static void staging_tasks(const Tensor* t, void* stage, int layout, bool to_tensor, std::vector<CopyTask>* tasks) {
  const size_t es = t->esize(); int64_t off = 0;
  for (const Blk& b : t->blks) {
    const int64_t nl = t->Vl.n[b.jl], nr = t->Vr.n[b.jr];
    char* sp = (char*)stage + off * es;
    const int64_t ssl = layout ? 1 : nr, ssr = layout ? nl : 1;
    if (to_tensor) tasks->push_back(CopyTask{sp, t->ptr(b.off), ssl, ssr, b.sl, b.sr, (int32_t)nl, (int32_t)nr, 0, 0, nullptr});
    else tasks->push_back(CopyTask{t->ptr(b.off), sp, b.sl, b.sr, ssl, ssr, (int32_t)nl, (int32_t)nr, 0, 0, nullptr});
    off += nl * nr;
  }
}
static int64_t dense_elems(const Tensor* t) { int64_t n = 0; for (const Blk& b : t->blks) n += t->Vl.n[b.jl] * t->Vr.n[b.jr]; return n; }
}  // namespace

int32_t mpsk_matvec_stream_create(mpsk_ctx* ctx, const mpsk_op* oph, const mpsk_tensor* x_like, const mpsk_tensor* y_like, int32_t depth,
                                  int32_t host_layout, mpsk_matvec_stream** out) { ENTER(ctx);
  Ctx* c = CTX(ctx);
  HeffOp* h = as_heff(oph);
  if (!h) return c->fail(MPSK_EINVAL, "not an effective-Hamiltonian operator");
  if (depth < 1 || depth > 8) return c->fail(MPSK_EINVAL, "depth must be 1..8");
  if (host_layout != 0 && host_layout != 1) return c->fail(MPSK_EINVAL, "host_layout: 0 row-major blocks, 1 column-major blocks");
  const Tensor* xl = CTEN(x_like); const Tensor* yl = CTEN(y_like);
  std::unique_ptr<MatvecStream> s(new MatvecStream);
  s->ctx = c; s->op = h; s->depth = depth; s->layout = host_layout;
  CUDA_TRY(c, cudaStreamCreateWithFlags(&s->h2d, cudaStreamNonBlocking));
  CUDA_TRY(c, cudaStreamCreateWithFlags(&s->d2h, cudaStreamNonBlocking));
  s->slots.resize(depth);
  for (auto& k : s->slots) {
    MPSK_TRY(tensor_create(c, xl->Vl, xl->mch, xl->Vr, xl->dtype, F_KL, true, &k.x));
    MPSK_TRY(tensor_create(c, yl->Vl, yl->mch, yl->Vr, yl->dtype, F_KL, true, &k.y));
    k.sx = c->alloc(std::max<int64_t>(dense_elems(k.x), 1) * k.x->esize());
    k.sy = c->alloc(std::max<int64_t>(dense_elems(k.y), 1) * k.y->esize());
    if (!k.sx || !k.sy) return MPSK_ENOMEM;
    std::vector<CopyTask> tin, tout;
    staging_tasks(k.x, k.sx, host_layout, true, &tin);
    staging_tasks(k.y, k.sy, host_layout, false, &tout);
    MPSK_TRY(copy_plan_build(c, tin, k.x->dtype == MPSK_C128, &k.in));
    MPSK_TRY(copy_plan_build(c, tout, k.y->dtype == MPSK_C128, &k.out));
    CUDA_TRY(c, cudaEventCreateWithFlags(&k.up, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventCreateWithFlags(&k.down, cudaEventDisableTiming));
  }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  *out = reinterpret_cast<mpsk_matvec_stream*>(s.release());
  return MPSK_OK;
}

int32_t mpsk_matvec_stream_submit(mpsk_ctx* ctx, mpsk_matvec_stream* sh, const mpsk_blockref* xblocks, int32_t nx,
                                  const mpsk_blockref_mut* yblocks, int32_t ny) { ENTER(ctx);
  Ctx* c = CTX(ctx); MatvecStream* s = reinterpret_cast<MatvecStream*>(sh);
  MatvecStream::Slot& k = s->slots[s->submitted % s->depth];
  if ((int)k.x->blks.size() != nx || (int)k.y->blks.size() != ny)
    return c->fail(MPSK_ESPACE, "stream submit: %d / %d blocks given, structures have %zu / %zu", nx, ny, k.x->blks.size(), k.y->blks.size());
  auto layout_ok = [&](int64_t nl, int64_t nr, int64_t sl, int64_t sr) { return s->layout ? contiguous_cm(nl, nr, sl, sr) : contiguous_rm(nl, nr, sl, sr); };
  for (int i = 0; i < nx; ++i) {
    const Blk& b = k.x->blks[i];
    if (!xblocks[i].ptr || !layout_ok(k.x->Vl.n[b.jl], k.x->Vr.n[b.jr], xblocks[i].stride_l, xblocks[i].stride_r))
      return c->fail(MPSK_EINVAL, "stream submit: x block %d is not a contiguous %s block", i, s->layout ? "column-major" : "row-major");
  }
  for (int i = 0; i < ny; ++i) {
    const Blk& b = k.y->blks[i];
    if (!yblocks[i].ptr || !layout_ok(k.y->Vl.n[b.jl], k.y->Vr.n[b.jr], yblocks[i].stride_l, yblocks[i].stride_r))
      return c->fail(MPSK_EINVAL, "stream submit: y block %d is not a contiguous %s block", i, s->layout ? "column-major" : "row-major");
  }
  if (k.busy) { CUDA_TRY(c, cudaEventSynchronize(k.down)); k.busy = false; }     // the slot's previous result has landed
  const size_t es = k.x->esize();
  int64_t off = 0;
  for (int i = 0; i < nx; ++i) {
    const Blk& b = k.x->blks[i];
    const int64_t n = k.x->Vl.n[b.jl] * k.x->Vr.n[b.jr];
    CUDA_TRY(c, cudaMemcpyAsync((char*)k.sx + off * es, xblocks[i].ptr, n * es, cudaMemcpyHostToDevice, s->h2d));
    off += n;
  }
  CUDA_TRY(c, cudaEventRecord(k.up, s->h2d));
  CUDA_TRY(c, cudaStreamWaitEvent(c->stream, k.up, 0));
  MPSK_TRY(copy_plan_run(c, &k.in));
  MPSK_TRY(heff_apply_panel(c, s->op, k.x, k.y));
  MPSK_TRY(copy_plan_run(c, &k.out));
  CUDA_TRY(c, cudaEventRecord(k.done, c->stream));
  CUDA_TRY(c, cudaStreamWaitEvent(s->d2h, k.done, 0));
  off = 0;
  for (int i = 0; i < ny; ++i) {
    const Blk& b = k.y->blks[i];
    const int64_t n = k.y->Vl.n[b.jl] * k.y->Vr.n[b.jr];
    CUDA_TRY(c, cudaMemcpyAsync(yblocks[i].ptr, (const char*)k.sy + off * es, n * es, cudaMemcpyDeviceToHost, s->d2h));
    off += n;
  }
  CUDA_TRY(c, cudaEventRecord(k.down, s->d2h));
  k.busy = true;
  s->submitted++;
  return MPSK_OK;
}

int32_t mpsk_matvec_stream_wait(mpsk_ctx* ctx, mpsk_matvec_stream* sh) { ENTER(ctx);
  Ctx* c = CTX(ctx); MatvecStream* s = reinterpret_cast<MatvecStream*>(sh);
  for (auto& k : s->slots) if (k.busy) { CUDA_TRY(c, cudaEventSynchronize(k.down)); k.busy = false; }
  return MPSK_OK;
}

void mpsk_matvec_stream_free(mpsk_ctx* ctx, mpsk_matvec_stream* sh) {
  if (!sh) return;
  ENTER(ctx);
  delete reinterpret_cast<MatvecStream*>(sh);
}

int32_t mpsk_tensor_panel(mpsk_ctx* ctx, const mpsk_tensor* t, int32_t axis, int32_t rank, int32_t nranks, mpsk_tensor** out) { ENTER(ctx);
  if (axis != 0 && axis != 1) return CTX(ctx)->fail(MPSK_EINVAL, "axis must be 0 (left index) or 1 (right index)");
  Tensor* o = nullptr;
  MPSK_TRY(tensor_panel(CTX(ctx), CTEN(t), axis, rank, nranks, &o));
  *out = reinterpret_cast<mpsk_tensor*>(o);
  return MPSK_OK;
}

int32_t mpsk_tensor_allgather(mpsk_ctx* ctx, const mpsk_tensor* panel, mpsk_tensor* full) { ENTER(ctx);
  return tensor_allgather_panels(CTX(ctx), CTEN(panel), TEN(full));
}

int32_t mpsk_allreduce_sum(mpsk_ctx* ctx, mpsk_tensor* xh) { ENTER(ctx);
  Ctx* c = CTX(ctx); Tensor* x = TEN(xh);
  if (c->nranks <= 1) return MPSK_OK;
  pfn_ncclAllReduce f = (pfn_ncclAllReduce)dlsym(nccl_lib(), "ncclAllReduce");
  if (!f || !c->nccl) return c->fail(MPSK_ECUDA, "communicator not initialised");
  const size_t count = (size_t)x->nelem * (x->dtype == MPSK_C128 ? 2 : 1);
  int r = f(x->d, x->d, count, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->nccl, c->stream);
  if (r != 0) return c->fail(MPSK_ECUDA, "ncclAllReduce failed (%d)", r);
  c->launches++;
  return MPSK_OK;
}

}  // extern "C"

namespace mpsk {
int comm_allreduce_sum(Ctx* c, void* buf, size_t ndoubles) {
  if (c->nranks <= 1 || ndoubles == 0) return MPSK_OK;
  typedef int (*pfn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  pfn_ncclAllReduce f = (pfn_ncclAllReduce)dlsym(nccl_lib(), "ncclAllReduce");
  if (!f || !c->nccl) return c->fail(MPSK_ECUDA, "communicator not initialised");
  int r = f(buf, buf, ndoubles, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->nccl, c->stream);
  if (r != 0) return c->fail(MPSK_ECUDA, "ncclAllReduce failed (%d)", r);
  c->launches++;
  return MPSK_OK;
}
int comm_allgather(Ctx* c, const void* send, void* recv, size_t bytes_per_rank) {
  typedef int (*pfn_ncclAllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
  pfn_ncclAllGather f = (pfn_ncclAllGather)dlsym(nccl_lib(), "ncclAllGather");
  if (!f || !c->nccl) return c->fail(MPSK_ECUDA, "communicator not initialised");
  int r = f(send, recv, bytes_per_rank / 8, /*ncclDouble*/ 8, c->nccl, c->stream);
  if (r != 0) return c->fail(MPSK_ECUDA, "ncclAllGather failed (%d)", r);
  c->launches++;
  return MPSK_OK;
}
}  // namespace mpsk
