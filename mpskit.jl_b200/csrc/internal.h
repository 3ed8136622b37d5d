// Internal structures of libmpskit_b200 (not part of the ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <complex>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>
#include "../../include/mpskit_b200.h"
#include "gemm.cuh"
#include "devalloc.h"

namespace mpsk {

typedef std::complex<double> cplx;

struct Ctx;
#define CUDA_TRY(ctx, expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) return (ctx)->fail(MPSK_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define MPSK_TRY(expr) do { int _s = (expr); if (_s != MPSK_OK) return _s; } while (0)

constexpr int64_t ALIGN = 4;          // leading dimensions / segment widths are padded to this many elements
constexpr int64_t GROUP_ALIGN = 32;   // group matrices start on multiples of this many elements
static inline int64_t pad(int64_t n) { return (n + ALIGN - 1) / ALIGN * ALIGN; }
static inline int64_t padg(int64_t n) { return (n + GROUP_ALIGN - 1) / GROUP_ALIGN * GROUP_ALIGN; }

struct TsvdSpectrum { std::vector<int32_t> charge, count, kept; std::vector<double> values; };

struct Ctx {
  int device = 0;
  uint64_t flags = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int num_sms = 148;
  std::string err;
  int64_t launches = 0;
  // device memory: chunked best-fit allocator with coalescing (devalloc.h); stream-ordered use on the context stream
  BlockAllocator dev;
  int64_t bytes_in_use = 0;
  void* cusolver = nullptr;     // cusolverDnHandle_t, created lazily
  void* pool_stream[4] = {nullptr, nullptr, nullptr, nullptr};   // worker streams / cuSOLVER handles of the factorisation pool
  void* pool_solver[4] = {nullptr, nullptr, nullptr, nullptr};
  int* pool_info[4] = {nullptr, nullptr, nullptr, nullptr};
  void* pool = nullptr;         // WorkerPool* (svd.cu): persistent host threads driving the per-sector factorizations
  void* nccl = nullptr;         // ncclComm_t
  int rank = 0, nranks = 1;
  double* h_scratch = nullptr;  // pinned host scratch (scalars coming back from Krylov kernels)
  void* h_stage = nullptr;      // cached pinned staging buffer for strided host blocks
  size_t h_stage_bytes = 0;
  double* d_scratch = nullptr;  // device scratch for reductions
  size_t d_scratch_bytes = 0;
  TsvdSpectrum last_spectrum;
  // pinned bump arena for the small task / tile tables of one-shot launches: the H2D copies are asynchronous and the
  // arena is only recycled after a stream synchronisation, so building a launch never blocks the host
  char* h_arena = nullptr; size_t h_arena_bytes = 0, h_arena_used = 0;
  void* arena_take(size_t bytes);      // pinned host memory valid until the next wrap (which synchronises the stream)
  bool gemm_attr_set[8] = {false, false, false, false, false, false, false, false};   // cudaFuncSetAttribute is per device -> per context
  int profile = -1;                    // MPSK_PROFILE=1: host timers around the Krylov phases
  double t_apply = 0, t_orth = 0, t_sync = 0;
  std::recursive_mutex mu;             // every ABI entry locks the context: concurrent callers serialise
  // Leased scratch slots: the multi-GB transient workspaces of an update (GEMM intermediates of the effective Hamiltonian
  // and of the environment update, the Krylov basis, the SVD arena) live in persistent per-context buffers that only ever
  // grow.  A slot that is busy (e.g. two operators alive at once) falls back to the allocator.  Without this the
  // workspaces -- ~25 GB allocated and released per centre update at D = 4096, with sizes drifting from site to site --
  // go through the memory pool, which degrades badly once > 100 GB of environments are resident (full sweeps on 1 GPU).
  enum { S_HEFF_T = 0, S_HEFF_U, S_TR_T, S_TR_Q, S_TSVD, S_KRYLOV, S_KRYLOV_SPARE, S_NSLOT };
  struct Scratch { void* p = nullptr; size_t cap = 0; bool busy = false; };
  Scratch scratch[S_NSLOT];
  void* lease(int slot, size_t bytes, bool* leased);
  void unlease(int slot, void* p, bool leased);

  void* alloc(size_t bytes);
  void free(void* p);
  void release_cache();
  int fail(int code, const char* fmt, ...);
};

struct Space {
  std::vector<int32_t> q;
  std::vector<int64_t> n;
  int32_t mod = 0;
  int nsec() const { return (int)q.size(); }
  int64_t dim() const { int64_t s = 0; for (auto v : n) s += v; return s; }
  int64_t norm(int64_t c) const { if (!mod) return c; int64_t r = c % mod; return r < 0 ? r + mod : r; }
  int find(int64_t c) const {
    c = norm(c);
    auto it = std::lower_bound(q.begin(), q.end(), (int32_t)c);
    return (it != q.end() && *it == c) ? (int)(it - q.begin()) : -1;
  }
  bool operator==(const Space& o) const { return q == o.q && n == o.n && mod == o.mod; }
};
int space_from_abi(Ctx* ctx, const mpsk_space* s, Space* out);
int upload_table(Ctx* ctx, const void* src, size_t bytes, void** d_out);   // async H2D of a small host table via the pinned arena

// Device layouts ("forms") of a T[l, m, r] tensor.  Every form is a set of dense row-major group
// matrices, one per sector of the grouping leg, chosen so that the contraction GEMMs need no
// standalone permutes (kernel family K2 "eliminated by layout choice"):
//   KL: group = l-sector, rows (m, r), cols l           (l contiguous)   -- x / y vectors, ket tensors
//   KR: group = r-sector, rows (m, l), cols r           (r contiguous)   -- left environments
//   SL: group = r-sector, rows r, cols segments (m, l)  (l contiguous)   -- K-stacked operands
//   SR: group = l-sector, rows l, cols segments (m, r)  (r contiguous)
enum Form { F_KL = 0, F_KR = 1, F_SL = 2, F_SR = 3 };

struct Blk { int32_t m, jl, jr; int64_t off, sl, sr; };       // element offset and strides of block (m, jl)
struct Group { int64_t off, rows, cols, ld; };

struct Tensor {
  Ctx* ctx = nullptr;
  Space Vl, Vr;
  std::vector<int32_t> mch;
  int dtype = MPSK_C128;
  Form form = F_KL;
  std::vector<Blk> blks;               // canonical order: m major, l-sector minor (== oracle block_keys)
  std::vector<int32_t> blk_index;      // [m * nsec_l + jl] -> index into blks or -1
  std::vector<Group> groups;
  std::vector<int64_t> seg;            // [m * ngroups + g] row (K forms) / column (S forms) offset of label m in group g, -1 absent
  void* d = nullptr;
  bool owns = true;                    // false: `d` aliases another buffer (views used by the Krylov solvers)
  int64_t nelem = 0;                   // elements including padding
  int nm() const { return (int)mch.size(); }
  int ngroups() const { return (int)groups.size(); }
  size_t esize() const { return dtype == MPSK_C128 ? 16 : 8; }
  size_t bytes() const { return (size_t)nelem * esize(); }
  const Blk* find(int m, int jl) const { int i = blk_index[(size_t)m * Vl.nsec() + jl]; return i < 0 ? nullptr : &blks[i]; }
  char* ptr(int64_t off) const { return (char*)d + off * (int64_t)esize(); }
  bool same_structure(const Tensor& o) const { return Vl == o.Vl && Vr == o.Vr && mch == o.mch && dtype == o.dtype; }
  ~Tensor() { if (d && ctx && owns) ctx->free(d); }
};

int tensor_create(Ctx* ctx, const Space& Vl, const std::vector<int32_t>& mch, const Space& Vr, int dtype, Form form,
                  bool zero, Tensor** out);
// copy with optional complex conjugation into another form (new tensor)
int tensor_repack(Ctx* ctx, const Tensor* src, Form form, bool conj, Tensor** out);
// returns src itself when it already has the form (and !conj); *owned tells whether the caller must delete
int tensor_as_form(Ctx* ctx, const Tensor* src, Form form, bool conj, const Tensor** out, bool* owned);

int tensor_fill_random(Ctx* ctx, Tensor* t, uint64_t seed);
int tensor_stack_labels(Ctx* ctx, const std::vector<const Tensor*>& ts, Tensor** out);
int tensor_select_labels(Ctx* ctx, const Tensor* t, const int32_t* labels, int n, Tensor** out);
int tensor_adjoint(Ctx* ctx, const Tensor* t, Tensor** out);
void tensor_layout_only(Tensor* t);   // fills blks / groups / nelem for (Vl, mch, Vr, dtype, form) without allocating
struct CopyTask;
void panel_table(const Space& V, int nranks, std::vector<int64_t>* lo, std::vector<int64_t>* hi);
void panel_bounds(const Space& V, int j, int rank, int nranks, int64_t* lo, int64_t* hi);
int tensor_row_panel(Ctx* ctx, const Tensor* t, int rank, int nranks, Tensor** out);
int tensor_panel(Ctx* ctx, const Tensor* t, int axis, int rank, int nranks, Tensor** out);   // axis 0: l index, 1: r index
int scatter_panel_tasks_axis(Ctx* ctx, const void* src, int axis, int rank, int nranks, Tensor* full, std::vector<CopyTask>* tasks);
int64_t panel_nelem_axis(const Tensor* full, int axis, int rank, int nranks);
int tensor_allgather_panels(Ctx* ctx, const Tensor* panel, Tensor* full);
int tensor_scatter_panel(Ctx* ctx, const void* src, int rank, int nranks, Tensor* full, bool sync);
struct CopyTask;
int scatter_panel_tasks(Ctx* ctx, const void* src, int rank, int nranks, Tensor* full, std::vector<CopyTask>* tasks);
int64_t panel_nelem(const Tensor* full, int rank, int nranks);
int comm_allgather(Ctx* ctx, const void* send, void* recv, size_t bytes_per_rank);

// ---- strided block copies (upload / download / repack / SVD assembly)
struct CopyTask {
  const void* src; void* dst;
  int64_t ssl, ssr, dsl, dsr;   // element strides of the l and r index in src / dst
  int32_t nl, nr;
  int32_t conj;                 // conjugate while copying
  int32_t scale_axis;           // 0 none, 1 scale[l], 2 scale[r]
  const double* scale;          // real scale vector on the device
};
int run_copy_tasks(Ctx* ctx, const std::vector<CopyTask>& tasks, bool cplx);
struct CopyPlan { Ctx* ctx = nullptr; bool cplx = true; void* d_tasks = nullptr; void* d_tiles = nullptr; int ntiles = 0; ~CopyPlan(); };
int copy_plan_build(Ctx* ctx, const std::vector<CopyTask>& tasks, bool cplx, CopyPlan* plan);
int copy_plan_run(Ctx* ctx, const CopyPlan* plan);

// ---- "mix": out_rect = sum_terms coeff * src_rect  (row-major rectangles, unit column stride)
struct MixTerm { const void* src; int64_t lds; double cre, cim; };
struct MixTask { void* dst; int64_t ldd; int32_t rows, cols; int32_t term_begin, term_end; };
struct MixPlan {
  Ctx* ctx = nullptr; bool cplx = true;
  void* d_tasks = nullptr; void* d_terms = nullptr; void* d_tiles = nullptr; int ntiles = 0;
  double bytes = 0;   // algorithmic bytes moved per run
  ~MixPlan();
};
int mix_plan_build(Ctx* ctx, bool cplx, const std::vector<MixTask>& tasks, const std::vector<MixTerm>& terms, MixPlan* plan);
int mix_plan_run(Ctx* ctx, const MixPlan* plan);

// ---- grouped GEMM
struct GemmPlan {
  Ctx* ctx = nullptr; bool cplx = true;
  GemmProblemDev* d_probs = nullptr; GemmTile* d_tiles = nullptr;
  int nprob = 0, ntiles = 0; double flops = 0;
  bool narrow = false;      // 128 x 32 tile configuration (complex, narrow-N plans)
  bool narrow16 = false;    // 128 x 16 tile configuration (row panels far from a multiple of 32)
  bool m3 = false;          // 3-multiplication complex product (gemm.cuh)
  int splitk = 1; void* partials = nullptr; MixPlan* reduce = nullptr;    // split-K partial products + their reduction pass
  ~GemmPlan();
};
int gemm_plan_build(Ctx* ctx, bool cplx, const std::vector<GemmProblemHost>& probs, GemmPlan* plan);
int gemm_plan_run(Ctx* ctx, const GemmPlan* plan);

// ---- Krylov BLAS-1 on flat vectors.  Coefficients that come from reductions stay on the device, so a
// Lanczos/Arnoldi step needs a single host synchronisation (kernel family K3).
int vec_multidot(Ctx* ctx, bool cplx, int nv, const void* const* V, const void* w, int64_t nelem, double* d_out /*2*nv*/);
int vec_multiaxpy(Ctx* ctx, bool cplx, int nv, const void* const* V, const double* d_coef /*2*nv*/, double sign, void* w, int64_t nelem);
int vec_norm2(Ctx* ctx, bool cplx, const void* x, int64_t nelem, double* d_out /*1*/);
int vec_scale_rsqrt(Ctx* ctx, bool cplx, void* x, int64_t nelem, const double* d_norm2);   // x /= sqrt(*d_norm2)
int vec_scale(Ctx* ctx, bool is_cplx, void* x, int64_t nelem, cplx a);
int vec_axpy(Ctx* ctx, bool is_cplx, cplx a, const void* x, void* y, int64_t nelem);
int vec_add_dev(Ctx* ctx, int n, const double* a, double* b);   // b[i] += a[i] (tiny, device scalars)

// ---- MPO slice
struct MpoSlice {
  int odim = 0, d = 0;
  std::vector<int32_t> cl, cr, pch, kind;         // kind[i*odim+j]
  std::vector<cplx> scal;                         // [i*odim+j]
  std::vector<std::vector<cplx>> ops;             // [i*odim+j] -> d*d row-major O[p_out,p_in] (empty unless kind==2)
  cplx elem(int i, int j, int po, int pi) const {
    int k = kind[i * odim + j];
    if (k == 0) return 0.0;
    if (k == 1) return po == pi ? scal[i * odim + j] : cplx(0.0);
    return ops[i * odim + j][po * d + pi];
  }
};

// ---- linear operators the Krylov solvers act on (effective Hamiltonians and transfer matrices)
// Krylov vectors are tensors with the structure of `local_like()`: the full vector for ordinary operators, this rank's
// row panel for a sharded effective Hamiltonian (then the inner products are partial sums + one fused allreduce).
struct LinOp {
  Ctx* ctx = nullptr; int dtype = MPSK_C128;
  virtual ~LinOp() {}
  virtual const Tensor* like() const = 0;                          // structure of the full vectors (ABI side)
  virtual const Tensor* local_like() const { return like(); }      // structure of the Krylov vectors of this rank
  virtual bool sharded() const { return false; }
  virtual int apply(const Tensor* x, Tensor* y) = 0;               // y = f(x) on full vectors
  virtual int apply_k(const Tensor* v, Tensor* w) { return apply(v, w); }   // on Krylov (local) vectors
  virtual int to_local(const Tensor* x, Tensor* loc);              // default: copy
  virtual int from_local(const Tensor* loc, Tensor* x);            // default: copy
  virtual double flops() const { return 0.0; }
};

// ---- effective Hamiltonian = stage-A GEMMs -> mix -> stage-C GEMMs
struct WEntry { int32_t i, k, mout, min; cplx c; };
struct HeffOp : LinOp {
  int kind = 0;
  std::unique_ptr<Tensor> GLk;   // left env in KR form (owned copy)
  std::unique_ptr<Tensor> GRs;   // right env in SL form (owned copy)
  std::unique_ptr<Tensor> xs;    // input staging buffer / structure template of x (KL form)
  std::unique_ptr<Tensor> ybuf;  // output staging buffer (KL form, bra structure); plans are bound to xs / ybuf
  bool rectangular = false;      // bra space of GL != ket space of x: this op produces a row panel of y (multi-GPU sharding)
  void* gather = nullptr; int64_t gather_elems = 0;   // allgather buffer (nranks x max panel elems) + one send slot
  std::unique_ptr<Tensor> yfull; std::unique_ptr<CopyPlan> scatter;     // gathered panels -> yfull (full-vector apply)
  std::unique_ptr<CopyPlan> scatter_x;                                   // gathered panels -> xs   (Krylov apply)
  std::vector<WEntry> W;
  // workspaces (leased context slots when free)
  void* T = nullptr; void* U = nullptr; int64_t T_elems = 0, U_elems = 0; bool T_leased = false, U_leased = false;
  std::unique_ptr<GemmPlan> gemmA, gemmC; std::unique_ptr<MixPlan> mix;
  double flops_ = 0;
  ~HeffOp();
  const Tensor* like() const override { return xs.get(); }
  const Tensor* local_like() const override { return rectangular ? ybuf.get() : xs.get(); }
  bool sharded() const override { return rectangular; }
  int apply(const Tensor* x, Tensor* y) override;
  int apply_k(const Tensor* v, Tensor* w) override;
  int to_local(const Tensor* x, Tensor* loc) override;
  int from_local(const Tensor* loc, Tensor* x) override;
  double flops() const override { return flops_; }
  int ensure_gather();
};
int heff_create(Ctx* ctx, int kind, const MpoSlice* h1, const MpoSlice* h2, const Tensor* GL, const Tensor* GR,
                const Tensor* x_like, HeffOp** out);
int heff_apply(Ctx* ctx, HeffOp* op, const Tensor* x, Tensor* y);
int heff_apply_local(Ctx* ctx, HeffOp* op, const Tensor* x, Tensor* y);
int heff_apply_panel(Ctx* ctx, HeffOp* op, const Tensor* x_panel, Tensor* y_panel);

// ---- one-site transfer with persistent plans: out = v * T(A, O, Abar) (left) or T(A, O, Abar) * v (right)
struct TransferPlan {
  Ctx* ctx = nullptr; bool left = true; bool cx = true;
  std::unique_ptr<Tensor> Ak, Bs;            // site tensors in their GEMM forms (owned copies)
  std::unique_ptr<Tensor> vin;               // input staging (KR form for left, KL for right)
  std::unique_ptr<Tensor> vout;              // output in its native form (KR for left, KL for right)
  std::unique_ptr<Tensor> in_like, out_like; // KL-form structure templates of input / output (no storage)
  std::unique_ptr<CopyPlan> in_repack, out_repack;   // KL input -> vin (left only); vout -> KL out_kl (left only)
  std::unique_ptr<Tensor> out_kl;            // left only: KL-form output staging
  void* T = nullptr; void* Q = nullptr; bool T_leased = false, Q_leased = false;
  GemmPlan g1, g3; MixPlan mx;
  double flops = 0;
  // row-panel sharding of the output bra index (multi-GPU): this rank computes its panel, allgather assembles the rest
  bool sharded = false; void* gather = nullptr; int64_t gather_elems = 0; std::unique_ptr<CopyPlan> scatter;
  std::unique_ptr<Tensor> out_full;
  ~TransferPlan();
  const Tensor* result() const { return sharded ? out_full.get() : (left ? out_kl.get() : vout.get()); }
};
int transfer_plan_create(Ctx* ctx, bool left, const MpoSlice* h, const Tensor* A, const Tensor* Abar, const Tensor* v_like,
                         bool sharded, TransferPlan** out);
int transfer_plan_run(Ctx* ctx, TransferPlan* p, const Tensor* v);     // result in p->result() (KL form)

// ---- transfer matrix over a unit cell as a linear operator (transfermatrix.jl:43-90), optionally regularised:
// v -> T(v) - <l, T(v)> r
struct TmOp : LinOp {
  bool left = true;                                   // true: v * tm (sweep left -> right); false: tm * v
  std::vector<std::unique_ptr<TransferPlan>> sites;   // in application order
  std::unique_ptr<Tensor> vlike;                      // structure of the vectors (KL form, no storage... owns a buffer for apply staging)
  std::unique_ptr<Tensor> ladj, rvec;                 // regularisation: ladj = adjoint(lvec) (dot partner), rvec
  double* d_coef = nullptr;
  double flops_ = 0;
  ~TmOp();
  const Tensor* like() const override { return vlike.get(); }
  int apply(const Tensor* x, Tensor* y) override;
  double flops() const override { return flops_; }
};

// ---- direct sum of operators with a cyclic shift: y[(r + shift) mod n] = ops[r](x[r]) -- the multi-row (MPSMultiline)
// derivatives of the statmech VUMPS (derivatives.jl:107-110,163-166,195-197: RecursiveVec + circshift).  The composite
// vector is ONE tensor whose labels are the rows' labels stacked (mpsk_tensor_stack_labels), so the Krylov BLAS-1 and
// every solver work on it unchanged; all rows must share their bond spaces.
struct MultiOp : LinOp {
  std::vector<LinOp*> ops;                            // not owned
  int shift = 1;
  std::unique_ptr<Tensor> vlike;                      // stacked structure (storage unused)
  std::vector<std::unique_ptr<Tensor>> xin, yout;     // per-row staging vectors
  const Tensor* like() const override { return vlike.get(); }
  int apply(const Tensor* x, Tensor* y) override;
  double flops() const override { double f = 0; for (auto* o : ops) f += o->flops(); return f; }
};
int multiop_create(Ctx* ctx, const std::vector<LinOp*>& ops, int shift, MultiOp** out);

// ---- Krylov solvers (eigsolve.cu)
int eigsolve_lanczos(Ctx* ctx, LinOp* op, const Tensor* x0, int which, int krylovdim, int maxiter, double tol,
                     cplx* val, Tensor* vec, mpsk_info* info);
int eigsolve_arnoldi(Ctx* ctx, LinOp* op, const Tensor* x0, int which, int krylovdim, int maxiter, double tol,
                     cplx* val, Tensor* vec, mpsk_info* info, int howmany = 1, cplx* vals = nullptr, int* nvals = nullptr, int* nconv = nullptr);
int expv_lanczos(Ctx* ctx, LinOp* op, cplx tau, const Tensor* y0, int krylovdim, int maxiter, double tol, Tensor* y,
                 mpsk_info* info);
// (a0 + a1 f) x = b, restarted GMRES (KrylovKit.linsolve, mpohaminfenv.jl:95,113,146,164)
int linsolve_gmres(Ctx* ctx, LinOp* op, const Tensor* b, const Tensor* x0, cplx a0, cplx a1, int krylovdim, int maxiter,
                   double tol, Tensor* x, mpsk_info* info);
// small dense host linear algebra
void jacobi_eigh(int n, std::vector<double>& A /*n*n row-major, destroyed*/, std::vector<double>& evals, std::vector<double>& evecs);
// complex Schur decomposition H = Z T Z^H of a general n x n matrix (column-major, overwritten by T), Z accumulated
int schur_complex(int n, std::vector<cplx>& H, std::vector<cplx>& Z);
// reorder the Schur form so that the eigenvalue at position `from` moves to position `to` (to <= from)
void schur_move(int n, std::vector<cplx>& T, std::vector<cplx>& Z, int from, int to);

// ---- transfers, local algebra, factorizations
int transfer_left(Ctx* ctx, const MpoSlice* h, const Tensor* A, const Tensor* Abar, const Tensor* GL, Tensor** out);
int transfer_right(Ctx* ctx, const MpoSlice* h, const Tensor* A, const Tensor* Abar, const Tensor* GR, Tensor** out);
int contract_ac2(Ctx* ctx, const Tensor* A1, const Tensor* A2, Tensor** out);
int mul_bond_right(Ctx* ctx, const Tensor* A, const Tensor* C, Tensor** out);
int mul_bond_left(Ctx* ctx, const Tensor* C, const Tensor* A, Tensor** out);
int tsvd(Ctx* ctx, const Tensor* ac2, int d1, const int32_t* p1_charge, mpsk_trunc trunc, Tensor** U, Tensor** S, Tensor** Vh, double* eps);
int qrpos(Ctx* ctx, const Tensor* A, bool left, Tensor** Q, Tensor** R);
void pool_destroy(Ctx* ctx);
int bond_inverse(Ctx* ctx, const Tensor* C, Tensor** out);
int nullspace(Ctx* ctx, const Tensor* A, bool left, Tensor** out);
int contract_adjoint_left(Ctx* ctx, const Tensor* N, const Tensor* x2, Tensor** out);
int tensor_cat(Ctx* ctx, const Tensor* A, const Tensor* B, int axis, Tensor** out);
int comm_allreduce_sum(Ctx* ctx, void* buf, size_t ndoubles);

}  // namespace mpsk

// ABI handles are opaque pointers to the structs above (reinterpret_cast in api.cu).
