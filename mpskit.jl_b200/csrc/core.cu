// Context, allocator, tensor layouts, strided block copies and the "mix" kernel.
#include "internal.h"

namespace mpsk {

// ------------------------------------------------------------------------------------------ context
int Ctx::fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  err = buf;
  return code;
}

static void* cuda_chunk_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
static void cuda_chunk_free(void* p) { cudaFree(p); }

// Device memory of a context: see devalloc.h.  Chunks of 8 GB (or the request, if larger) come from cudaMalloc; in steady
// state no driver call is made.
void* Ctx::alloc(size_t bytes) {
  if (!dev.chunk_alloc) { dev.chunk_alloc = cuda_chunk_alloc; dev.chunk_free = cuda_chunk_free; }
  void* p = dev.alloc(bytes);
  if (!p) {
    cudaStreamSynchronize(stream);
    dev.trim();
    p = dev.alloc(bytes);
    if (!p) { fail(MPSK_ENOMEM, "device allocation of %zu bytes failed (%zu live, %zu reserved)", bytes, dev.bytes_live, dev.bytes_reserved); return nullptr; }
  }
  bytes_in_use = (int64_t)dev.bytes_live;
  return p;
}

void Ctx::free(void* p) {
  if (!p) return;
  dev.free(p);
  bytes_in_use = (int64_t)dev.bytes_live;
}

void Ctx::release_cache() {
  cudaStreamSynchronize(stream);
  dev.trim();
}

void* Ctx::lease(int slot, size_t bytes, bool* leased) {
  Scratch& sc = scratch[slot];
  if (sc.busy) { *leased = false; return alloc(bytes); }
  if (sc.cap < bytes) {
    if (sc.p) free(sc.p);
    sc.p = nullptr; sc.cap = 0;
    const size_t want = bytes + bytes / 8 + ((size_t)1 << 20);
    sc.p = alloc(want);
    if (!sc.p) { sc.p = alloc(bytes); if (!sc.p) { *leased = false; return nullptr; } sc.cap = bytes; }
    else sc.cap = want;
  }
  sc.busy = true; *leased = true;
  return sc.p;
}

void Ctx::unlease(int slot, void* p, bool leased) {
  if (!p) return;
  if (leased) scratch[slot].busy = false;
  else free(p);
}


// pinned bump arena (see internal.h): memory handed out stays valid until the arena wraps, and wrapping synchronises
// the stream first, so asynchronous H2D copies out of it never read recycled bytes
void* Ctx::arena_take(size_t bytes) {
  bytes = (bytes + 63) / 64 * 64;
  if (bytes > h_arena_bytes) {
    cudaStreamSynchronize(stream);
    if (h_arena) cudaFreeHost(h_arena);
    h_arena = nullptr; h_arena_bytes = 0; h_arena_used = 0;
    const size_t want = std::max<size_t>(bytes * 2, (size_t)8 << 20);
    if (cudaMallocHost(&h_arena, want) != cudaSuccess) { cudaGetLastError(); fail(MPSK_ENOMEM, "pinned arena allocation failed"); return nullptr; }
    h_arena_bytes = want;
  }
  if (h_arena_used + bytes > h_arena_bytes) { cudaStreamSynchronize(stream); h_arena_used = 0; }
  void* p = h_arena + h_arena_used;
  h_arena_used += bytes;
  return p;
}

// stream-ordered upload of a small host table (pinned arena -> device, no host synchronisation)
int upload_table(Ctx* ctx, const void* src, size_t bytes, void** d_out) {
  void* h = ctx->arena_take(bytes);
  void* d = ctx->alloc(bytes);
  if (!h || !d) { if (d) ctx->free(d); return MPSK_ENOMEM; }
  memcpy(h, src, bytes);
  CUDA_TRY(ctx, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *d_out = d;
  return MPSK_OK;
}

int space_from_abi(Ctx* ctx, const mpsk_space* s, Space* out) {
  if (!s || s->nsec < 0) return ctx->fail(MPSK_EINVAL, "null space");
  out->q.assign(s->charge, s->charge + s->nsec);
  out->n.assign(s->dim, s->dim + s->nsec);
  out->mod = s->modulus;
  for (int i = 0; i < s->nsec; ++i) {
    if (out->n[i] <= 0) return ctx->fail(MPSK_ESPACE, "sector %d has non-positive dimension", i);
    if (i && out->q[i] <= out->q[i - 1]) return ctx->fail(MPSK_ESPACE, "sectors must be listed in strictly ascending charge");
    if (out->mod && (out->q[i] < 0 || out->q[i] >= out->mod)) return ctx->fail(MPSK_ESPACE, "Z_N charge out of range");
  }
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ layouts
static void build_layout(Tensor* t) {
  const int nl = t->Vl.nsec(), nr = t->Vr.nsec(), nm = t->nm();
  t->blks.clear();
  t->blk_index.assign((size_t)nm * nl, -1);
  for (int m = 0; m < nm; ++m)
    for (int jl = 0; jl < nl; ++jl) {
      int jr = t->Vr.find((int64_t)t->Vl.q[jl] + t->mch[m]);
      if (jr < 0) continue;
      t->blk_index[(size_t)m * nl + jl] = (int)t->blks.size();
      t->blks.push_back(Blk{m, jl, jr, 0, 0, 0});
    }
  const bool by_l = (t->form == F_KL || t->form == F_SR);
  const bool kform = (t->form == F_KL || t->form == F_KR);
  const int ng = by_l ? nl : nr;
  t->groups.assign(ng, Group{0, 0, 0, 0});
  t->seg.assign((size_t)nm * ng, -1);
  int64_t off = 0;
  for (int g = 0; g < ng; ++g) {
    Group& G = t->groups[g];
    const int64_t ngdim = by_l ? t->Vl.n[g] : t->Vr.n[g];
    int64_t run = 0;
    for (int m = 0; m < nm; ++m) {
      int jl, jr;
      if (by_l) { jl = g; jr = t->Vr.find((int64_t)t->Vl.q[g] + t->mch[m]); }
      else { jr = g; jl = t->Vl.find((int64_t)t->Vr.q[g] - t->mch[m]); }
      if (jl < 0 || jr < 0) continue;
      const int64_t other = by_l ? t->Vr.n[jr] : t->Vl.n[jl];
      t->seg[(size_t)m * ng + g] = run;
      run += kform ? other : pad(other);
    }
    if (kform) { G.rows = run; G.cols = ngdim; G.ld = pad(ngdim); }
    else { G.rows = ngdim; G.cols = run; G.ld = pad(run); }
    G.off = off;
    off += padg(G.rows * G.ld);
  }
  t->nelem = std::max<int64_t>(off, GROUP_ALIGN);
  for (Blk& b : t->blks) {
    const int g = by_l ? b.jl : b.jr;
    const Group& G = t->groups[g];
    const int64_t s = t->seg[(size_t)b.m * ng + g];
    switch (t->form) {
      case F_KL: b.off = G.off + s * G.ld; b.sl = 1; b.sr = G.ld; break;
      case F_KR: b.off = G.off + s * G.ld; b.sl = G.ld; b.sr = 1; break;
      case F_SL: b.off = G.off + s; b.sl = 1; b.sr = G.ld; break;
      case F_SR: b.off = G.off + s; b.sl = G.ld; b.sr = 1; break;
    }
  }
}

void tensor_layout_only(Tensor* t) { build_layout(t); }

int tensor_create(Ctx* ctx, const Space& Vl, const std::vector<int32_t>& mch, const Space& Vr, int dtype, Form form,
                  bool zero, Tensor** out) {
  if (Vl.mod != Vr.mod) return ctx->fail(MPSK_ESPACE, "bond spaces belong to different groups");
  std::unique_ptr<Tensor> t(new Tensor);
  t->ctx = ctx; t->Vl = Vl; t->Vr = Vr; t->mch = mch; t->dtype = dtype; t->form = form;
  build_layout(t.get());
  t->d = ctx->alloc(t->bytes());
  if (!t->d) return MPSK_ENOMEM;
  if (zero) CUDA_TRY(ctx, cudaMemsetAsync(t->d, 0, t->bytes(), ctx->stream));
  *out = t.release();
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ strided copy
struct CopyTile { int32_t task, l0, r0, pad; };

template <typename T> __device__ __forceinline__ T conj_if(T v, int) { return v; }
template <> __device__ __forceinline__ double2 conj_if<double2>(double2 v, int c) { if (c) v.y = -v.y; return v; }
__device__ __forceinline__ double scale_val(double v, double s) { return v * s; }
__device__ __forceinline__ double2 scale_val(double2 v, double s) { return make_double2(v.x * s, v.y * s); }

template <typename T>
__global__ void __launch_bounds__(256) copy_kernel(const CopyTask* __restrict__ tasks, const CopyTile* __restrict__ tiles, int ntiles) {
  __shared__ T tile[32][33];   // [r][l]
  for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
    const CopyTile tl = tiles[ti];
    const CopyTask tk = tasks[tl.task];
    const T* src = (const T*)tk.src;
    T* dst = (T*)tk.dst;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // read: fast thread index along the contiguous source axis
    if (tk.ssl <= tk.ssr) {
      const int l = tl.l0 + tx;
#pragma unroll
      for (int rr = ty; rr < 32; rr += 8) {
        const int r = tl.r0 + rr;
        if (l < tk.nl && r < tk.nr) tile[rr][tx] = src[l * tk.ssl + r * tk.ssr];
      }
    } else {
      const int r = tl.r0 + tx;
#pragma unroll
      for (int ll = ty; ll < 32; ll += 8) {
        const int l = tl.l0 + ll;
        if (l < tk.nl && r < tk.nr) tile[tx][ll] = src[l * tk.ssl + r * tk.ssr];
      }
    }
    __syncthreads();
    if (tk.dsl <= tk.dsr) {
      const int l = tl.l0 + tx;
#pragma unroll
      for (int rr = ty; rr < 32; rr += 8) {
        const int r = tl.r0 + rr;
        if (l < tk.nl && r < tk.nr) {
          T v = conj_if<T>(tile[rr][tx], tk.conj);
          if (tk.scale_axis == 1) v = scale_val(v, tk.scale[l]);
          else if (tk.scale_axis == 2) v = scale_val(v, tk.scale[r]);
          dst[l * tk.dsl + r * tk.dsr] = v;
        }
      }
    } else {
      const int r = tl.r0 + tx;
#pragma unroll
      for (int ll = ty; ll < 32; ll += 8) {
        const int l = tl.l0 + ll;
        if (l < tk.nl && r < tk.nr) {
          T v = conj_if<T>(tile[tx][ll], tk.conj);
          if (tk.scale_axis == 1) v = scale_val(v, tk.scale[l]);
          else if (tk.scale_axis == 2) v = scale_val(v, tk.scale[r]);
          dst[l * tk.dsl + r * tk.dsr] = v;
        }
      }
    }
    __syncthreads();
  }
}

int run_copy_tasks(Ctx* ctx, const std::vector<CopyTask>& tasks, bool cplx) {
  if (tasks.empty()) return MPSK_OK;
  std::vector<CopyTile> tiles;
  for (size_t i = 0; i < tasks.size(); ++i)
    for (int r0 = 0; r0 < tasks[i].nr; r0 += 32)
      for (int l0 = 0; l0 < tasks[i].nl; l0 += 32) tiles.push_back(CopyTile{(int)i, l0, r0, 0});
  if (tiles.empty()) return MPSK_OK;
  // the tables travel through the pinned arena: no host synchronisation; the device copies are released in stream order
  void *d_tasks = nullptr, *d_tiles = nullptr;
  MPSK_TRY(upload_table(ctx, tasks.data(), tasks.size() * sizeof(CopyTask), &d_tasks));
  int st = upload_table(ctx, tiles.data(), tiles.size() * sizeof(CopyTile), &d_tiles);
  if (st != MPSK_OK) { ctx->free(d_tasks); return st; }
  const int grid = (int)std::min<size_t>(tiles.size(), (size_t)ctx->num_sms * 8);
  if (cplx) copy_kernel<double2><<<grid, 256, 0, ctx->stream>>>((const CopyTask*)d_tasks, (const CopyTile*)d_tiles, (int)tiles.size());
  else copy_kernel<double><<<grid, 256, 0, ctx->stream>>>((const CopyTask*)d_tasks, (const CopyTile*)d_tiles, (int)tiles.size());
  ctx->launches++;
  ctx->free(d_tasks); ctx->free(d_tiles);
  CUDA_TRY(ctx, cudaGetLastError());
  return MPSK_OK;
}

// persistent copy plan: task / tile tables stay on the device, one launch per run, no host synchronisation
CopyPlan::~CopyPlan() { if (ctx) { ctx->free(d_tasks); ctx->free(d_tiles); } }

int copy_plan_build(Ctx* ctx, const std::vector<CopyTask>& tasks, bool cplx, CopyPlan* plan) {
  plan->ctx = ctx; plan->cplx = cplx; plan->ntiles = 0;
  std::vector<CopyTile> tiles;
  for (size_t i = 0; i < tasks.size(); ++i)
    for (int r0 = 0; r0 < tasks[i].nr; r0 += 32)
      for (int l0 = 0; l0 < tasks[i].nl; l0 += 32) tiles.push_back(CopyTile{(int)i, l0, r0, 0});
  if (tiles.empty()) return MPSK_OK;
  MPSK_TRY(upload_table(ctx, tasks.data(), tasks.size() * sizeof(CopyTask), &plan->d_tasks));
  MPSK_TRY(upload_table(ctx, tiles.data(), tiles.size() * sizeof(CopyTile), &plan->d_tiles));
  plan->ntiles = (int)tiles.size();
  return MPSK_OK;
}

int copy_plan_run(Ctx* ctx, const CopyPlan* plan) {
  if (!plan->ntiles) return MPSK_OK;
  const int grid = std::min(plan->ntiles, ctx->num_sms * 8);
  if (plan->cplx) copy_kernel<double2><<<grid, 256, 0, ctx->stream>>>((const CopyTask*)plan->d_tasks, (const CopyTile*)plan->d_tiles, plan->ntiles);
  else copy_kernel<double><<<grid, 256, 0, ctx->stream>>>((const CopyTask*)plan->d_tasks, (const CopyTile*)plan->d_tiles, plan->ntiles);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return MPSK_OK;
}

int tensor_repack(Ctx* ctx, const Tensor* src, Form form, bool conj, Tensor** out) {
  Tensor* t = nullptr;
  MPSK_TRY(tensor_create(ctx, src->Vl, src->mch, src->Vr, src->dtype, form, true, &t));
  std::vector<CopyTask> tasks;
  for (size_t b = 0; b < src->blks.size(); ++b) {
    const Blk& s = src->blks[b];
    const Blk& d = t->blks[b];
    tasks.push_back(CopyTask{src->ptr(s.off), t->ptr(d.off), s.sl, s.sr, d.sl, d.sr, (int32_t)src->Vl.n[s.jl],
                             (int32_t)src->Vr.n[s.jr], conj && src->dtype == MPSK_C128 ? 1 : 0, 0, nullptr});
  }
  int st = run_copy_tasks(ctx, tasks, src->dtype == MPSK_C128);
  if (st != MPSK_OK) { delete t; return st; }
  *out = t;
  return MPSK_OK;
}

int tensor_as_form(Ctx* ctx, const Tensor* src, Form form, bool conj, const Tensor** out, bool* owned) {
  if (src->form == form && !(conj && src->dtype == MPSK_C128)) { *out = src; *owned = false; return MPSK_OK; }
  Tensor* t = nullptr;
  MPSK_TRY(tensor_repack(ctx, src, form, conj, &t));
  *out = t; *owned = true;
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ label algebra / adjoint
// out[l, (labels of all inputs), r]: environments are handled level by level in the infinite-MPS solvers
// (mpohaminfenv.jl:76-215) and as one stacked tensor by the matvec
int tensor_stack_labels(Ctx* ctx, const std::vector<const Tensor*>& ts, Tensor** out) {
  if (ts.empty()) return ctx->fail(MPSK_EINVAL, "stack: no tensors");
  std::vector<int32_t> mch;
  for (const Tensor* t : ts) {
    if (!(t->Vl == ts[0]->Vl) || !(t->Vr == ts[0]->Vr) || t->dtype != ts[0]->dtype) return ctx->fail(MPSK_ESPACE, "stack: tensors differ in their bond spaces");
    mch.insert(mch.end(), t->mch.begin(), t->mch.end());
  }
  Tensor* o = nullptr;
  MPSK_TRY(tensor_create(ctx, ts[0]->Vl, mch, ts[0]->Vr, ts[0]->dtype, F_KL, true, &o));
  std::unique_ptr<Tensor> og(o);
  std::vector<CopyTask> tasks;
  int base = 0;
  for (const Tensor* t : ts) {
    for (const Blk& b : t->blks) {
      const Blk* d = o->find(base + b.m, b.jl);
      if (!d) return ctx->fail(MPSK_EINTERNAL, "stack: missing block");
      tasks.push_back(CopyTask{t->ptr(b.off), o->ptr(d->off), b.sl, b.sr, d->sl, d->sr, (int32_t)t->Vl.n[b.jl], (int32_t)t->Vr.n[b.jr], 0, 0, nullptr});
    }
    base += t->nm();
  }
  MPSK_TRY(run_copy_tasks(ctx, tasks, o->dtype == MPSK_C128));
  *out = og.release();
  return MPSK_OK;
}

int tensor_select_labels(Ctx* ctx, const Tensor* t, const int32_t* labels, int n, Tensor** out) {
  std::vector<int32_t> mch(n);
  for (int i = 0; i < n; ++i) { if (labels[i] < 0 || labels[i] >= t->nm()) return ctx->fail(MPSK_EINVAL, "select: label out of range"); mch[i] = t->mch[labels[i]]; }
  Tensor* o = nullptr;
  MPSK_TRY(tensor_create(ctx, t->Vl, mch, t->Vr, t->dtype, F_KL, true, &o));
  std::unique_ptr<Tensor> og(o);
  std::vector<CopyTask> tasks;
  for (const Blk& d : o->blks) {
    const Blk* b = t->find(labels[d.m], d.jl);
    if (!b) return ctx->fail(MPSK_EINTERNAL, "select: missing block");
    tasks.push_back(CopyTask{t->ptr(b->off), o->ptr(d.off), b->sl, b->sr, d.sl, d.sr, (int32_t)t->Vl.n[b->jl], (int32_t)t->Vr.n[b->jr], 0, 0, nullptr});
  }
  MPSK_TRY(run_copy_tasks(ctx, tasks, o->dtype == MPSK_C128));
  *out = og.release();
  return MPSK_OK;
}

// out[r, m, l] = conj(t[l, m, r]) with label charges negated (adjoint of bond matrices and of single env levels)
int tensor_adjoint(Ctx* ctx, const Tensor* t, Tensor** out) {
  std::vector<int32_t> mch(t->nm());
  for (int i = 0; i < t->nm(); ++i) mch[i] = -t->mch[i];
  Tensor* o = nullptr;
  MPSK_TRY(tensor_create(ctx, t->Vr, mch, t->Vl, t->dtype, F_KL, true, &o));
  std::unique_ptr<Tensor> og(o);
  std::vector<CopyTask> tasks;
  for (const Blk& b : t->blks) {
    const Blk* d = o->find(b.m, b.jr);
    if (!d || d->jr != b.jl) return ctx->fail(MPSK_EINTERNAL, "adjoint: missing block");
    // d is indexed (l' = r idx, r' = l idx): element (il, ir) of b goes to (l' = ir, r' = il)
    tasks.push_back(CopyTask{t->ptr(b.off), o->ptr(d->off), b.sl, b.sr, d->sr, d->sl, (int32_t)t->Vl.n[b.jl], (int32_t)t->Vr.n[b.jr],
                             t->dtype == MPSK_C128 ? 1 : 0, 0, nullptr});
  }
  MPSK_TRY(run_copy_tasks(ctx, tasks, o->dtype == MPSK_C128));
  *out = og.release();
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ row panels (multi-GPU sharding)
// Rank r of N owns rows [floor(n r / N), floor(n (r+1) / N)) of every l-sector (dimension n): a perfectly balanced split
// of the bra/left index.  Sectors whose share is empty disappear from the panel's space.
// Row-panel rule (pure integer arithmetic; mirrored bit-exactly by mpskit_b200/sharding.py:panel_table): rank r of N
// owns rows [floor(n r / N), floor(n (r+1) / N)) of every sector (dimension n) of the left bond -- every rank gets the
// same share of every sector, hence exactly 1/N of the flops whatever the operator.  (A tile-aligned, weight-balanced
// rule was measured and rejected in round 1: 19.0 ms vs 16.4 ms per matvec at 8 GPUs -- load imbalance costs more than
// the 13 % tile-quantisation loss of the even split; tools/panel_cost.py is the offline model.)
void panel_table(const Space& V, int nranks, std::vector<int64_t>* lo, std::vector<int64_t>* hi) {
  const int ns = V.nsec();
  lo->assign((size_t)nranks * ns, 0); hi->assign((size_t)nranks * ns, 0);
  for (int r = 0; r < nranks; ++r)
    for (int j = 0; j < ns; ++j) {
      (*lo)[(size_t)r * ns + j] = V.n[j] * r / nranks;
      (*hi)[(size_t)r * ns + j] = V.n[j] * (r + 1) / nranks;
    }
}

void panel_bounds(const Space& V, int j, int rank, int nranks, int64_t* lo, int64_t* hi) {
  std::vector<int64_t> l, h;
  panel_table(V, nranks, &l, &h);
  *lo = l[(size_t)rank * V.nsec() + j]; *hi = h[(size_t)rank * V.nsec() + j];
}

static Space panel_space(const Space& V, int rank, int nranks) {
  Space P; P.mod = V.mod;
  std::vector<int64_t> l, h;
  panel_table(V, nranks, &l, &h);
  for (int j = 0; j < V.nsec(); ++j) {
    const int64_t lo = l[(size_t)rank * V.nsec() + j], hi = h[(size_t)rank * V.nsec() + j];
    if (hi > lo) { P.q.push_back(V.q[j]); P.n.push_back(hi - lo); }
  }
  return P;
}

// out = the part of t whose l index (axis 0) or r index (axis 1) belongs to `rank`; any form in, KL form out
int tensor_panel(Ctx* ctx, const Tensor* t, int axis, int rank, int nranks, Tensor** out) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return ctx->fail(MPSK_EINVAL, "bad rank / nranks");
  const Space& Vfull = axis == 0 ? t->Vl : t->Vr;
  Space P = panel_space(Vfull, rank, nranks);
  Tensor* o = nullptr;
  if (axis == 0) MPSK_TRY(tensor_create(ctx, P, t->mch, t->Vr, t->dtype, F_KL, true, &o));
  else MPSK_TRY(tensor_create(ctx, t->Vl, t->mch, P, t->dtype, F_KL, true, &o));
  std::unique_ptr<Tensor> og(o);
  std::vector<CopyTask> tasks;
  for (const Blk& b : o->blks) {
    const int jl_full = axis == 0 ? t->Vl.find(P.q[b.jl]) : b.jl;
    const Blk* s = t->find(b.m, jl_full);
    if (!s) return ctx->fail(MPSK_EINTERNAL, "panel: missing block");
    int64_t lo, hi;
    if (axis == 0) {
      panel_bounds(t->Vl, jl_full, rank, nranks, &lo, &hi);
      tasks.push_back(CopyTask{t->ptr(s->off + lo * s->sl), o->ptr(b.off), s->sl, s->sr, b.sl, b.sr, (int32_t)(hi - lo), (int32_t)t->Vr.n[s->jr], 0, 0, nullptr});
    } else {
      if (t->Vr.q[s->jr] != P.q[b.jr]) return ctx->fail(MPSK_EINTERNAL, "panel: sector mismatch");
      panel_bounds(t->Vr, s->jr, rank, nranks, &lo, &hi);
      tasks.push_back(CopyTask{t->ptr(s->off + lo * s->sr), o->ptr(b.off), s->sl, s->sr, b.sl, b.sr, (int32_t)t->Vl.n[s->jl], (int32_t)(hi - lo), 0, 0, nullptr});
    }
  }
  MPSK_TRY(run_copy_tasks(ctx, tasks, t->dtype == MPSK_C128));
  *out = og.release();
  return MPSK_OK;
}

int tensor_row_panel(Ctx* ctx, const Tensor* t, int rank, int nranks, Tensor** out) { return tensor_panel(ctx, t, 0, rank, nranks, out); }

static void panel_layout(const Tensor* full, int axis, int rank, int nranks, Tensor* lay, Space* P) {
  *P = panel_space(axis == 0 ? full->Vl : full->Vr, rank, nranks);
  lay->ctx = nullptr; lay->mch = full->mch; lay->dtype = full->dtype; lay->form = F_KL;
  if (axis == 0) { lay->Vl = *P; lay->Vr = full->Vr; } else { lay->Vl = full->Vl; lay->Vr = *P; }
  tensor_layout_only(lay);
}

// full[part of `rank`] = panel data; `src` points at a buffer laid out like the KL-form panel tensor of that rank
int scatter_panel_tasks_axis(Ctx* ctx, const void* src, int axis, int rank, int nranks, Tensor* full, std::vector<CopyTask>* tasks) {
  Tensor lay; Space P;
  panel_layout(full, axis, rank, nranks, &lay, &P);
  const size_t es = full->esize();
  for (const Blk& b : lay.blks) {
    const int jl_full = axis == 0 ? full->Vl.find(P.q[b.jl]) : b.jl;
    const Blk* d = full->find(b.m, jl_full);
    if (!d) return ctx->fail(MPSK_EINTERNAL, "scatter_panel: missing block");
    int64_t lo, hi;
    if (axis == 0) {
      panel_bounds(full->Vl, jl_full, rank, nranks, &lo, &hi);
      tasks->push_back(CopyTask{(const char*)src + b.off * es, full->ptr(d->off + lo * d->sl), b.sl, b.sr, d->sl, d->sr, (int32_t)(hi - lo),
                                (int32_t)full->Vr.n[d->jr], 0, 0, nullptr});
    } else {
      panel_bounds(full->Vr, d->jr, rank, nranks, &lo, &hi);
      tasks->push_back(CopyTask{(const char*)src + b.off * es, full->ptr(d->off + lo * d->sr), b.sl, b.sr, d->sl, d->sr, (int32_t)full->Vl.n[d->jl],
                                (int32_t)(hi - lo), 0, 0, nullptr});
    }
  }
  return MPSK_OK;
}

int scatter_panel_tasks(Ctx* ctx, const void* src, int rank, int nranks, Tensor* full, std::vector<CopyTask>* tasks) {
  return scatter_panel_tasks_axis(ctx, src, 0, rank, nranks, full, tasks);
}

int tensor_scatter_panel(Ctx* ctx, const void* src, int rank, int nranks, Tensor* full, bool sync) {
  std::vector<CopyTask> tasks;
  MPSK_TRY(scatter_panel_tasks_axis(ctx, src, 0, rank, nranks, full, &tasks));
  (void)sync;
  return run_copy_tasks(ctx, tasks, full->dtype == MPSK_C128);
}

int64_t panel_nelem_axis(const Tensor* full, int axis, int rank, int nranks) {
  Tensor lay; Space P;
  panel_layout(full, axis, rank, nranks, &lay, &P);
  return lay.nelem;
}
int64_t panel_nelem(const Tensor* full, int rank, int nranks) { return panel_nelem_axis(full, 0, rank, nranks); }

// every rank contributes its axis-0 panel (KL form) of a tensor with the structure of `full`; afterwards `full` holds
// the whole tensor on every rank (used to re-activate environments that are stored sharded)
int tensor_allgather_panels(Ctx* ctx, const Tensor* panel, Tensor* full) {
  if (ctx->nranks <= 1) return ctx->fail(MPSK_EINVAL, "allgather without a communicator");
  if (panel->form != F_KL || full->form != F_KL) return ctx->fail(MPSK_EINTERNAL, "allgather: tensors must be in KL form");
  const size_t es = full->esize();
  int64_t mx = 0;
  for (int r = 0; r < ctx->nranks; ++r) mx = std::max(mx, panel_nelem(full, r, ctx->nranks));
  if (panel->nelem != panel_nelem(full, ctx->rank, ctx->nranks) || panel->dtype != full->dtype) return ctx->fail(MPSK_ESPACE, "tensor is not this rank's row panel");
  char* buf = (char*)ctx->alloc((size_t)mx * (ctx->nranks + 1) * es);
  if (!buf) return MPSK_ENOMEM;
  char* send = buf + (size_t)mx * ctx->nranks * es;
  CUDA_TRY(ctx, cudaMemsetAsync(send, 0, (size_t)mx * es, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(send, panel->d, panel->bytes(), cudaMemcpyDeviceToDevice, ctx->stream));
  int st = comm_allgather(ctx, send, buf, (size_t)mx * es);
  if (st == MPSK_OK) {
    std::vector<CopyTask> tasks;
    for (int r = 0; r < ctx->nranks && st == MPSK_OK; ++r) st = scatter_panel_tasks(ctx, buf + (size_t)r * mx * es, r, ctx->nranks, full, &tasks);
    if (st == MPSK_OK) st = run_copy_tasks(ctx, tasks, full->dtype == MPSK_C128);
  }
  ctx->free(buf);
  return st;
}

// ------------------------------------------------------------------------------------------ random fill
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double u01(uint64_t h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

struct FillTask { void* dst; int64_t sl, sr; int32_t nl, nr; uint64_t base; };

// uniform(-1,1) entries inside the logical blocks only (padding stays zero); value depends on (seed, block, i, j) only
template <bool CPLX>
__global__ void fill_kernel(const FillTask* __restrict__ tasks, int ntasks, uint64_t seed) {
  for (int t = blockIdx.y; t < ntasks; t += gridDim.y) {
    const FillTask tk = tasks[t];
    const int64_t total = (int64_t)tk.nl * tk.nr;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      const int64_t il = e / tk.nr, ir = e - il * tk.nr;
      const uint64_t h = splitmix64(seed ^ splitmix64(tk.base + (uint64_t)e));
      if (CPLX) {
        double2 v = make_double2(2.0 * u01(h) - 1.0, 2.0 * u01(splitmix64(h)) - 1.0);
        ((double2*)tk.dst)[il * tk.sl + ir * tk.sr] = v;
      } else {
        ((double*)tk.dst)[il * tk.sl + ir * tk.sr] = 2.0 * u01(h) - 1.0;
      }
    }
  }
}

int tensor_fill_random(Ctx* ctx, Tensor* t, uint64_t seed) {
  std::vector<FillTask> tasks;
  uint64_t base = 0;
  for (const Blk& b : t->blks) {
    tasks.push_back(FillTask{t->ptr(b.off), b.sl, b.sr, (int32_t)t->Vl.n[b.jl], (int32_t)t->Vr.n[b.jr], base});
    base += (uint64_t)t->Vl.n[b.jl] * t->Vr.n[b.jr];
  }
  if (tasks.empty()) return MPSK_OK;
  void* dv = nullptr;
  MPSK_TRY(upload_table(ctx, tasks.data(), tasks.size() * sizeof(FillTask), &dv));
  FillTask* d = (FillTask*)dv;
  dim3 grid(64, (unsigned)std::min<size_t>(tasks.size(), 1024));
  if (t->dtype == MPSK_C128) fill_kernel<true><<<grid, 256, 0, ctx->stream>>>(d, (int)tasks.size(), seed);
  else fill_kernel<false><<<grid, 256, 0, ctx->stream>>>(d, (int)tasks.size(), seed);
  ctx->launches++;
  ctx->free(d);
  CUDA_TRY(ctx, cudaGetLastError());
  return MPSK_OK;
}

// ------------------------------------------------------------------------------------------ mix
// out[row, col] = sum_t coeff_t * src_t[row, col] over row-major rectangles: HBM streaming.
// A CTA owns a tile of the destination rectangle, the term table of the task sits in shared memory, every thread keeps
// MIX_R independent row accumulators for one column (MIX_R loads in flight per term and thread, fully coalesced along
// the contiguous column index, no integer division anywhere).  Two tile shapes: 256 columns x 64 rows for wide
// rectangles, 64 columns x 256 rows for narrow ones (small sectors); each thread walks MIX_G groups of MIX_R rows.
struct MixTile { int32_t task, row0, col0, shape; };
constexpr int MIX_R = 8;          // rows per thread
constexpr int MIX_MAXT = 32;      // terms staged per round
constexpr int MIX_G = 8;          // row groups per tile

template <typename T> __device__ __forceinline__ T mix_zero();
template <> __device__ __forceinline__ double mix_zero<double>() { return 0.0; }
template <> __device__ __forceinline__ double2 mix_zero<double2>() { return make_double2(0.0, 0.0); }
__device__ __forceinline__ void mix_fma(double& acc, double v, double cre, double) { acc += cre * v; }
__device__ __forceinline__ void mix_fma(double2& acc, double2 v, double cre, double cim) {
  acc.x += cre * v.x - cim * v.y; acc.y += cre * v.y + cim * v.x;
}

template <typename T>
__global__ void __launch_bounds__(256, 2) mix_kernel(const MixTask* __restrict__ tasks, const MixTerm* __restrict__ terms,
                                                     const MixTile* __restrict__ tiles, int ntiles) {
  __shared__ MixTerm sterm[MIX_MAXT];
  for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
    const MixTile tl = tiles[ti];
    const MixTask tk = tasks[tl.task];
    const int nt = tk.term_end - tk.term_begin;
    // a tile is MIX_G row groups deep, so the descriptor fetches (tile -> task -> terms: a chain of dependent global
    // loads) are paid once per MIX_G * MIX_R rows
    int col, rbase, rstep;
    if (tl.shape == 0) { col = tl.col0 + (int)threadIdx.x; rbase = tl.row0; rstep = MIX_R; }
    else { col = tl.col0 + ((int)threadIdx.x & 63); rbase = tl.row0 + ((int)threadIdx.x >> 6) * MIX_R; rstep = 4 * MIX_R; }
    const bool colok = col < tk.cols;
    if (nt <= MIX_MAXT) {
      __syncthreads();
      if ((int)threadIdx.x < nt) sterm[threadIdx.x] = terms[tk.term_begin + threadIdx.x];
      __syncthreads();
      if (colok) {
        for (int g = 0; g < MIX_G; ++g) {
          const int row0 = rbase + g * rstep;
          const int nrow = min(MIX_R, tk.rows - row0);
          if (nrow <= 0) break;
          T acc[MIX_R];
#pragma unroll
          for (int r = 0; r < MIX_R; ++r) acc[r] = mix_zero<T>();
          for (int t = 0; t < nt; ++t) {
            const double cre = sterm[t].cre, cim = sterm[t].cim;
            const int64_t lds = sterm[t].lds;
            const T* src = (const T*)sterm[t].src + (int64_t)row0 * lds + col;
            T v[MIX_R];
#pragma unroll
            for (int r = 0; r < MIX_R; ++r) v[r] = r < nrow ? src[(int64_t)r * lds] : mix_zero<T>();
#pragma unroll
            for (int r = 0; r < MIX_R; ++r) mix_fma(acc[r], v[r], cre, cim);
          }
          T* dst = (T*)tk.dst + (int64_t)row0 * tk.ldd + col;
#pragma unroll
          for (int r = 0; r < MIX_R; ++r) if (r < nrow) dst[(int64_t)r * tk.ldd] = acc[r];
        }
      }
    } else {
      // more terms than the shared table holds (split-K reductions with many slices never get here; kept for generality)
      for (int g = 0; g < MIX_G; ++g) {
        const int row0 = rbase + g * rstep;
        const int nrow = min(MIX_R, tk.rows - row0);
        T acc[MIX_R];
#pragma unroll
        for (int r = 0; r < MIX_R; ++r) acc[r] = mix_zero<T>();
        for (int t0 = 0; t0 < nt; t0 += MIX_MAXT) {
          const int cnt = min(MIX_MAXT, nt - t0);
          __syncthreads();
          if ((int)threadIdx.x < cnt) sterm[threadIdx.x] = terms[tk.term_begin + t0 + threadIdx.x];
          __syncthreads();
          if (colok && nrow > 0)
            for (int t = 0; t < cnt; ++t) {
              const T* src = (const T*)sterm[t].src + (int64_t)row0 * sterm[t].lds + col;
#pragma unroll
              for (int r = 0; r < MIX_R; ++r) if (r < nrow) mix_fma(acc[r], src[(int64_t)r * sterm[t].lds], sterm[t].cre, sterm[t].cim);
            }
        }
        if (colok && nrow > 0) {
          T* dst = (T*)tk.dst + (int64_t)row0 * tk.ldd + col;
#pragma unroll
          for (int r = 0; r < MIX_R; ++r) if (r < nrow) dst[(int64_t)r * tk.ldd] = acc[r];
        }
      }
    }
  }
}

MixPlan::~MixPlan() {
  if (ctx) { ctx->free(d_tasks); ctx->free(d_terms); ctx->free(d_tiles); }
}

int mix_plan_build(Ctx* ctx, bool cplx, const std::vector<MixTask>& tasks, const std::vector<MixTerm>& terms, MixPlan* plan) {
  plan->ctx = ctx; plan->cplx = cplx; plan->ntiles = 0; plan->bytes = 0;
  std::vector<MixTile> tiles;
  const double es = cplx ? 16.0 : 8.0;
  for (size_t i = 0; i < tasks.size(); ++i) {
    const int rows = tasks[i].rows, cols = tasks[i].cols;
    if (rows <= 0 || cols <= 0) continue;
    if (cols > 96) {
      for (int r0 = 0; r0 < rows; r0 += MIX_R * MIX_G)
        for (int c0 = 0; c0 < cols; c0 += 256) tiles.push_back(MixTile{(int)i, r0, c0, 0});
    } else {
      for (int r0 = 0; r0 < rows; r0 += 4 * MIX_R * MIX_G)
        for (int c0 = 0; c0 < cols; c0 += 64) tiles.push_back(MixTile{(int)i, r0, c0, 1});
    }
    plan->bytes += es * (double)rows * cols * (1 + tasks[i].term_end - tasks[i].term_begin);
  }
  plan->ntiles = (int)tiles.size();
  if (tiles.empty()) return MPSK_OK;
  MPSK_TRY(upload_table(ctx, tasks.data(), tasks.size() * sizeof(MixTask), &plan->d_tasks));
  MixTerm dummy{nullptr, 0, 0.0, 0.0};
  MPSK_TRY(upload_table(ctx, terms.empty() ? (const void*)&dummy : (const void*)terms.data(), std::max<size_t>(terms.size(), 1) * sizeof(MixTerm), &plan->d_terms));
  MPSK_TRY(upload_table(ctx, tiles.data(), tiles.size() * sizeof(MixTile), &plan->d_tiles));
  return MPSK_OK;
}

int mix_plan_run(Ctx* ctx, const MixPlan* plan) {
  if (plan->ntiles == 0) return MPSK_OK;
  const int grid = std::min(plan->ntiles, ctx->num_sms * 8);
  if (plan->cplx)
    mix_kernel<double2><<<grid, 256, 0, ctx->stream>>>((const MixTask*)plan->d_tasks, (const MixTerm*)plan->d_terms,
                                                        (const MixTile*)plan->d_tiles, plan->ntiles);
  else
    mix_kernel<double><<<grid, 256, 0, ctx->stream>>>((const MixTask*)plan->d_tasks, (const MixTerm*)plan->d_terms,
                                                       (const MixTile*)plan->d_tiles, plan->ntiles);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return MPSK_OK;
}

}  // namespace mpsk
