/* mpskit_b200 -- C ABI of the B200-native MPSKit.jl hot path.
 *
 * The reference (MPSKit.jl v0.10.2, pure Julia) has no FFI; the seam is Julia multiple dispatch.
 * Every entry point below names the reference method (file:line under the reference tree) that a
 * thin `ccall` shim overloads with it (INTEGRATION.md shows the shim).  Plain pointers and sizes
 * only; no exceptions cross the boundary; every function returns an mpsk_status (0 = ok, <0 error)
 * and `mpsk_last_error` gives the message.  Thread safety: every entry point locks its context, so
 * calls from concurrent host threads / Julia tasks (vumps.jl:39-45, mpohaminfenv.jl:65-68) on the
 * same context serialise; distinct contexts are independent and may live on different devices (each
 * call selects the context's device).  Handles are freed through the context that created them.
 *
 * Data model (TensorKit abelian block structure).  All tensors on the path are of
 * the generic three-leg form  T[l, m, r]  with graded bond spaces l, r and a flat label set m whose
 * labels carry a charge; charge conservation is  q_r = q_l + mcharge[m]  (mod `modulus` for Z_N):
 *    MPS tensor  A[l p; r]           m = physical state,        mcharge = s_p
 *    two-site    x[l p1; r p2]       m = p1 * d2 + p2,          mcharge = s_p1 + s_p2
 *    bond        C[l; r]             m = one label,             mcharge = 0
 *    left env    GL[l_bra w; l_ket]  m = MPO level (width 1),   mcharge = -c_w
 *    right env   GR[r_ket w; r_bra]  m = MPO level,             mcharge = +c_w
 * Charges are int32 (2q for half-integer U(1) irreps).  Sectors are listed in ascending charge.
 * One dense block per (m, l-sector) whose r-sector exists; the host hands each block as a pointer
 * plus element strides, so TensorKit's column-major fusion-tree sub-blocks and numpy arrays are
 * both passed without copies on the caller's side.  Host blocks are only read during upload and
 * only written during download; device storage belongs to the library until mpsk_tensor_free.
 */
#ifndef MPSKIT_B200_H
#define MPSKIT_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpsk_ctx mpsk_ctx;
typedef struct mpsk_tensor mpsk_tensor;
typedef struct mpsk_mposlice mpsk_mposlice;
typedef struct mpsk_op mpsk_op;
typedef struct mpsk_matvec_stream mpsk_matvec_stream;

typedef enum {
  MPSK_OK = 0,
  MPSK_EINVAL = -1,    /* bad argument */
  MPSK_ESPACE = -2,    /* sector / space mismatch (TensorKit SpaceMismatch) */
  MPSK_ECUDA = -3,     /* CUDA runtime / cuSOLVER / NCCL failure */
  MPSK_ENOMEM = -4,
  MPSK_EINTERNAL = -5,
  MPSK_ENODEVICE = -6  /* no usable GPU: there is no CPU fallback */
} mpsk_status;

enum { MPSK_F64 = 0, MPSK_C128 = 1 };                 /* scalar type (Defaults.eltype = ComplexF64, defaults.jl:18) */
enum { MPSK_SR = 0, MPSK_LM = 1, MPSK_LR = 2 };       /* `which` of fixedpoint (fixedpoint.jl:9,19) */
enum { MPSK_NOTRUNC = 0, MPSK_TRUNCDIM = 1, MPSK_TRUNCERR = 2, MPSK_TRUNCBELOW = 3 }; /* TensorKit truncation schemes used at dmrg.jl:75,96 */
#define MPSK_TRUNC_NO_NORMALIZE 0x100   /* OR into mpsk_trunc.kind: skip normalize!(c) (TDVP2, tdvp.jl:124-126) */
enum { MPSK_OP_DC = 0, MPSK_OP_DAC = 1, MPSK_OP_DAC2 = 2 };
#define MPSK_OP_SHARDED 0x100   /* OR into `kind`: GL is this rank's row panel; apply/eigsolve/expv allgather the panels of y */

#define MPSK_FLAG_EXACT_SVD 2ull      /* tsvd: cuSOLVER gesvd for every sector (default: two-level Gram + heevd + QR route for sectors >= 32:
                                         every singular value to ~1e-11 sigma_max absolute) */
#define MPSK_FLAG_SVD_RELATIVE 16ull  /* tsvd: continue the Gram levels down to the rounding floor (full relative accuracy, up to 6 levels) */
#define MPSK_FLAG_FAST_SVD_ALL 4ull   /* tsvd: Gram route for every sector (tests) */
#define MPSK_FLAG_DEBUG_GEMM 1ull  /* route contractions through the one-thread-per-element cross-check kernel (tests only) */
#define MPSK_FLAG_GEMM_4M 32ull       /* ComplexF64 contractions with the conventional 4-multiplication complex product (default: 3M,
                                         three real DMMA products per complex one; same normwise error bound, 25 % less tensor work) */
#define MPSK_FLAG_REPLICATE_LOCAL 8ull /* with a communicator: keep environment updates and tsvd replicated per rank (default: distributed) */

typedef struct { int32_t nsec; const int32_t* charge; const int64_t* dim; int32_t modulus; } mpsk_space;
typedef struct { const void* ptr; int64_t stride_l; int64_t stride_r; } mpsk_blockref;   /* strides in scalar elements */
typedef struct { void* ptr; int64_t stride_l; int64_t stride_r; } mpsk_blockref_mut;
typedef struct { int32_t kind; double value; } mpsk_trunc;                               /* value: D or epsilon */
typedef struct { int32_t converged; int32_t numiter; int32_t numops; double normres; } mpsk_info; /* KrylovKit ConvergenceInfo */

/* ---- context ------------------------------------------------------------------------------ */
int32_t mpsk_create(int32_t device, uint64_t flags, mpsk_ctx** out);
void mpsk_destroy(mpsk_ctx* ctx);
const char* mpsk_last_error(const mpsk_ctx* ctx);
int32_t mpsk_version(void);
int32_t mpsk_set_stream(mpsk_ctx* ctx, void* cuda_stream);       /* adopt the caller's cudaStream_t (NULL = library stream) */
int32_t mpsk_synchronize(mpsk_ctx* ctx);
int64_t mpsk_launch_count(const mpsk_ctx* ctx);                  /* kernels launched so far by this context */
int64_t mpsk_bytes_in_use(const mpsk_ctx* ctx);

/* ---- tensors (TensorMap blocks <-> device) --------------------------------------------------- */
/* number and order of blocks for a structure: block b <-> (m[b], lsec[b], rsec[b]); arrays may be NULL to query the count.
 * Integer bookkeeping, must match the oracle bit-exactly. */
int32_t mpsk_block_table(const mpsk_space* Vl, int32_t nm, const int32_t* mcharge, const mpsk_space* Vr,
                         int32_t* nblocks, int32_t* m, int32_t* lsec, int32_t* rsec);
int32_t mpsk_tensor_upload(mpsk_ctx* ctx, const mpsk_space* Vl, int32_t nm, const int32_t* mcharge, const mpsk_space* Vr,
                           int32_t dtype, const mpsk_blockref* blocks, int32_t nblocks, mpsk_tensor** out);
int32_t mpsk_tensor_zeros(mpsk_ctx* ctx, const mpsk_space* Vl, int32_t nm, const int32_t* mcharge, const mpsk_space* Vr,
                          int32_t dtype, mpsk_tensor** out);
int32_t mpsk_tensor_download(mpsk_ctx* ctx, const mpsk_tensor* t, const mpsk_blockref_mut* blocks, int32_t nblocks);
/* label algebra: environments are vectors of MPO levels (mpohaminfenv.jl:76-215 works level by level, the matvec on the whole vector) */
int32_t mpsk_tensor_stack_labels(mpsk_ctx* ctx, const mpsk_tensor* const* ts, int32_t n, mpsk_tensor** out);
int32_t mpsk_tensor_select_labels(mpsk_ctx* ctx, const mpsk_tensor* t, const int32_t* labels, int32_t n, mpsk_tensor** out);
/* out[r, m, l] = conj(t[l, m, r]), label charges negated (adjoint of a bond tensor: ortho.jl:127-131 Q_C') */
int32_t mpsk_tensor_adjoint(mpsk_ctx* ctx, const mpsk_tensor* t, mpsk_tensor** out);
/* synthetic data: uniform(-1,1) entries in every symmetry-allowed block, reproducible from `seed` (benchmarks, property tests) */
int32_t mpsk_tensor_fill_random(mpsk_ctx* ctx, mpsk_tensor* t, uint64_t seed);
int32_t mpsk_tensor_copy(mpsk_ctx* ctx, const mpsk_tensor* t, mpsk_tensor** out);
void mpsk_tensor_free(mpsk_ctx* ctx, mpsk_tensor* t);
/* structure query (two-phase: counts first, then arrays) -- sizes after truncation (tsvd) */
int32_t mpsk_tensor_structure(const mpsk_tensor* t, int32_t* nsec_l, int32_t* nsec_r, int32_t* nm, int32_t* dtype);
int32_t mpsk_tensor_spaces(const mpsk_tensor* t, int32_t* charge_l, int64_t* dim_l, int32_t* charge_r, int64_t* dim_r, int32_t* mcharge);

/* ---- VectorInterface on handles (inner, norm, add!!, scale!!; KrylovKit vectors) -- */
int32_t mpsk_dot(mpsk_ctx* ctx, const mpsk_tensor* x, const mpsk_tensor* y, double* re, double* im);   /* <x|y>, conj on x */
int32_t mpsk_norm(mpsk_ctx* ctx, const mpsk_tensor* x, double* out);
int32_t mpsk_axpy(mpsk_ctx* ctx, double a_re, double a_im, const mpsk_tensor* x, mpsk_tensor* y);    /* y += a x */
int32_t mpsk_scale(mpsk_ctx* ctx, double a_re, double a_im, mpsk_tensor* x);

/* ---- MPO slice: SparseMPOSlice (src/operators/sparsempo/sparseslice.jl:13-106) ------------------
 * odim width-1 levels with charges; `kind[i*odim+j]`: 0 absent, 1 scalar (Os[i,j] isa E), 2 operator;
 * `scal[2*(i*odim+j)]` re/im of scalar entries; operator entries are d x d row-major complex
 * matrices O[p_out, p_in] concatenated in (i-major, j-minor) order of the kind==2 entries. */
int32_t mpsk_mposlice_create(mpsk_ctx* ctx, int32_t odim, const int32_t* level_charge_left, const int32_t* level_charge_right,
                             int32_t d, const int32_t* phys_charge, const int32_t* kind, const double* scal,
                             const double* ops, mpsk_mposlice** out);
/* same for a Z_N-graded slice: charges are compared modulo `modulus` (0 = U(1) / trivial) */
int32_t mpsk_mposlice_create_zn(mpsk_ctx* ctx, int32_t odim, const int32_t* level_charge_left, const int32_t* level_charge_right,
                                int32_t d, const int32_t* phys_charge, const int32_t* kind, const double* scal,
                                const double* ops, int32_t modulus, mpsk_mposlice** out);
void mpsk_mposlice_free(mpsk_ctx* ctx, mpsk_mposlice* s);
/* keys(x) (sparseslice.jl:74-76), first index fastest; returns count, fills (i,j) pairs if non-NULL */
int32_t mpsk_mposlice_keys(const mpsk_mposlice* s, int32_t* count, int32_t* ij);

/* ---- effective Hamiltonians: MPO_ddC / MPO_ddAC / MPO_ddAC2 (src/algorithms/derivatives.jl:6-71) ---
 * GL, GR are environment tensors (level = label).  h1/h2: slices at pos / pos+1 (h2 NULL unless DAC2;
 * both NULL for DC).  `x_like` fixes the block structure of the vectors the operator acts on. */
int32_t mpsk_op_create(mpsk_ctx* ctx, int32_t kind, const mpsk_mposlice* h1, const mpsk_mposlice* h2,
                       const mpsk_tensor* GL, const mpsk_tensor* GR, const mpsk_tensor* x_like, mpsk_op** out);
void mpsk_op_free(mpsk_ctx* ctx, mpsk_op* op);
double mpsk_op_flops(const mpsk_op* op);                           /* algorithmic real flops of one apply */
/* measurement hook: mean device time (ms) of the grouped-GEMM launches of one apply (CUDA events on the context stream) */
int32_t mpsk_op_time_gemms(mpsk_ctx* ctx, const mpsk_op* op, int32_t reps, double* ms_out);
/* same for the memory-bound "mix" launch; bytes_out = its algorithmic bytes (16 B x rectangle elements x (terms + 1)) */
int32_t mpsk_op_time_mix(mpsk_ctx* ctx, const mpsk_op* op, int32_t reps, double* ms_out, double* bytes_out);
/* y = h(x): dC / dAC / dAC2 (derivatives.jl:77-197), or tm * v / v * tm for a transfer-matrix operator. */
int32_t mpsk_op_apply(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* x, mpsk_tensor* y);
/* fixedpoint(h, x0, which, alg) (src/algorithms/fixedpoint.jl:9-30): Krylov eigensolve on the device.
 * hermitian != 0: Lanczos with full re-orthogonalisation and thick restarts (KrylovKit.Lanczos);
 * hermitian == 0: Arnoldi with Krylov-Schur restarts (KrylovKit.Arnoldi = Defaults.eigsolver, defaults.jl:33).
 * Both test convergence after every expansion step (eager = true). */
int32_t mpsk_eigsolve(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* x0, int32_t which, int32_t hermitian,
                      int32_t krylovdim, int32_t maxiter, double tol, double* val_re, double* val_im,
                      mpsk_tensor* vec, mpsk_info* info);
/* eigsolve(f, x0, howmany, which; tol) with Arnoldi / Krylov-Schur for the `howmany` most wanted eigenvalues (values only):
 * transfer_spectrum (src/algorithms/toolbox.jl:42-58: eigsolve(flip(TransferMatrix(above.AL, below.AL)), init, num_vals, :LM)).
 * vals: 2 * howmany doubles (re, im interleaved), most wanted first; *nvals (<= howmany) how many were produced (fewer only
 * if the Krylov space closed early), *nconverged how many leading ones have residuals below tol.  howmany <= 3/5 krylovdim. */
int32_t mpsk_eigsolve_values(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* x0, int32_t which, int32_t howmany,
                             int32_t krylovdim, int32_t maxiter, double tol, double* vals, int32_t* nvals, int32_t* nconverged,
                             mpsk_info* info);
/* integrate -> exponentiate(h, tau, y0) (src/algorithms/timestep/integrators.jl:20-25): y = exp(tau h) y0 */
int32_t mpsk_expv(mpsk_ctx* ctx, const mpsk_op* op, double tau_re, double tau_im, const mpsk_tensor* y0,
                  int32_t krylovdim, int32_t maxiter, double tol, mpsk_tensor* y, mpsk_info* info);

/* linsolve(f, b, x0, GMRES(tol, krylovdim, maxiter), a0, a1): solve (a0 + a1 f) x = b with restarted GMRES
 * (src/environments/mpohaminfenv.jl:95,113,146,164; tol is absolute on the residual norm). */
int32_t mpsk_linsolve(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* b, const mpsk_tensor* x0, double a0_re, double a0_im,
                      double a1_re, double a1_im, int32_t krylovdim, int32_t maxiter, double tol, mpsk_tensor* x, mpsk_info* info);
/* TransferMatrix(above, middle, below) over a unit cell of `nsites` sites as an operator handle (transfermatrix.jl:4-66):
 * left != 0: v -> v * tm (sites 0 .. n-1), else v -> tm * v (sites n-1 .. 0).  slices == NULL or slices[i] == NULL: MPO-free
 * transfer (transfer.jl:18-45).  With lvec_adj / rvec the operator is regularised (transfermatrix.jl:68-76):
 * v -> T(v) - <lvec_adj, T(v)> rvec, lvec_adj being adjoint(lvec) in the vector space.  Plans, TMA descriptors and workspaces are
 * built once; the handle works with mpsk_op_apply, mpsk_eigsolve (hermitian = 0) and mpsk_linsolve, and is freed by mpsk_op_free. */
int32_t mpsk_tmop_create(mpsk_ctx* ctx, int32_t left, int32_t nsites, const mpsk_mposlice* const* slices, const mpsk_tensor* const* A,
                         const mpsk_tensor* const* Abar, const mpsk_tensor* v_like, const mpsk_tensor* lvec_adj, const mpsk_tensor* rvec,
                         mpsk_op** out);

/* Multi-row (MPSMultiline) derivatives: (dAC, dAC2, dC)(x::RecursiveVec, ...) = RecursiveVec(circshift(map(...), 1))
 * (derivatives.jl:107-110,163-166,195-197).  The composite operator acts on ONE tensor whose labels are the rows' labels
 * stacked (mpsk_tensor_stack_labels / mpsk_tensor_select_labels convert to and from the RecursiveVec): y[(r + shift) mod n] =
 * ops[r](x[r]).  All rows must share their bond spaces.  `ops` must outlive the handle; freed by mpsk_op_free. */
int32_t mpsk_multiop_create(mpsk_ctx* ctx, int32_t n, const mpsk_op* const* ops, int32_t shift, mpsk_op** out);

/* ---- transfer matrices / environment updates (src/transfermatrix/transfer.jl:137-259,
 *      src/environments/FinEnv.jl:114-145) --------------------------------------------------------- */
int32_t mpsk_transfer_left(mpsk_ctx* ctx, const mpsk_mposlice* h, const mpsk_tensor* A, const mpsk_tensor* Abar,
                           const mpsk_tensor* GL, mpsk_tensor** GL_out);
int32_t mpsk_transfer_right(mpsk_ctx* ctx, const mpsk_mposlice* h, const mpsk_tensor* A, const mpsk_tensor* Abar,
                            const mpsk_tensor* GR, mpsk_tensor** GR_out);

/* ---- local tensor algebra of the sweep bodies ------------------------------------------------------ */
/* ac2[l p1; r p2] = A1[l p1; k] A2[k p2; r]   (dmrg.jl:92,108) */
int32_t mpsk_contract_ac2(mpsk_ctx* ctx, const mpsk_tensor* A1, const mpsk_tensor* A2, mpsk_tensor** out);
/* A[l p; r] * C[r; r'] and C[l'; l] * A[l p; r]   (orthoview.jl:95-106) */
int32_t mpsk_mul_bond_right(mpsk_ctx* ctx, const mpsk_tensor* A, const mpsk_tensor* C, mpsk_tensor** out);
int32_t mpsk_mul_bond_left(mpsk_ctx* ctx, const mpsk_tensor* C, const mpsk_tensor* A, mpsk_tensor** out);
/* tsvd!(ac2; trunc, alg=SVD()) + normalize!(c) (dmrg.jl:96-97,112-113): U = al[l p1; k], S = diag c (normalised,
 * returned as a bond tensor), Vh = ar as an MPS tensor [k p2; r] (i.e. _transpose_front(ar)).  d1: size of p1,
 * p1_charge[d1]: charges of the p1 basis states (fixes the charges of the new bond k = l + p1).
 * eps_trunc: 2-norm of the discarded singular values (of the un-normalised input). */
int32_t mpsk_tsvd(mpsk_ctx* ctx, const mpsk_tensor* ac2, int32_t d1, const int32_t* p1_charge, mpsk_trunc trunc, mpsk_tensor** U,
                  mpsk_tensor** S, mpsk_tensor** Vh, double* eps_trunc);
/* singular values of the last mpsk_tsvd, per coupled sector, before truncation: two-phase query */
int32_t mpsk_tsvd_spectrum(mpsk_ctx* ctx, int32_t* nsec, int32_t* charge, int32_t* count, int32_t* kept, double* values);
/* leftnull(A): isometry N[l p; k'] onto the orthogonal complement of the range of A[l p; r]; rightnull(A): N[k' p; r] with
 * orthonormal rows spanning the complement of the row space of A[l; p r] (changebonds/optimalexpand.jl:28-29,95-96) */
int32_t mpsk_leftnull(mpsk_ctx* ctx, const mpsk_tensor* A, mpsk_tensor** out);
int32_t mpsk_rightnull(mpsk_ctx* ctx, const mpsk_tensor* A, mpsk_tensor** out);
/* B[k' p2; r] = sum_{l,p1} conj(N[l p1; k']) x2[l (p1 p2); r]  (adjoint(VL) * AC2, optimalexpand.jl:30,98) */
int32_t mpsk_contract_adjoint_left(mpsk_ctx* ctx, const mpsk_tensor* N, const mpsk_tensor* x2, mpsk_tensor** out);
/* direct sum along the left (axis 0: catcodomain) or right (axis 1: catdomain) bond (optimalexpand.jl:103-104) */
int32_t mpsk_tensor_cat(mpsk_ctx* ctx, const mpsk_tensor* A, const mpsk_tensor* B, int32_t axis, mpsk_tensor** out);
/* inv(C) of a bond tensor, sector by sector (the edge updates of IDMRG2, idmrg.jl:130-134,172-173) */
int32_t mpsk_bond_inverse(mpsk_ctx* ctx, const mpsk_tensor* C, mpsk_tensor** out);
/* leftorth!(A; alg=QRpos()) / rightorth!(A; alg=LQpos()) (orthoview.jl:49-60) */
int32_t mpsk_qrpos(mpsk_ctx* ctx, const mpsk_tensor* A, mpsk_tensor** Q, mpsk_tensor** R);
int32_t mpsk_lqpos(mpsk_ctx* ctx, const mpsk_tensor* A, mpsk_tensor** L, mpsk_tensor** Q);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) ------------------------------------------- */
int32_t mpsk_comm_unique_id(uint8_t out128[128]);
int32_t mpsk_comm_init(mpsk_ctx* ctx, const uint8_t id128[128], int32_t rank, int32_t nranks);
int32_t mpsk_allreduce_sum(mpsk_ctx* ctx, mpsk_tensor* x);        /* partial-sum reduction */
/* Row-panel sharding of the matvec: rank r of N owns rows [floor(n r/N), floor(n (r+1)/N)) of every sector of the left
 * (bra) index (csrc/core.cu panel_table, mirrored by mpskit_b200/sharding.py).  An operator created with the row panel of GL computes the same panel of y = h(x) (no redundant flops, no
 * partial sums); on a context with a communicator mpsk_op_apply / mpsk_eigsolve / mpsk_expv allgather the panels (NCCL). */
int32_t mpsk_tensor_row_panel(mpsk_ctx* ctx, const mpsk_tensor* t, int32_t rank, int32_t nranks, mpsk_tensor** out);
int32_t mpsk_tensor_scatter_panel(mpsk_ctx* ctx, const mpsk_tensor* panel, int32_t rank, int32_t nranks, mpsk_tensor* full);
int32_t mpsk_op_apply_local(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* x, mpsk_tensor* y_panel);
/* panel in -> panel out: allgathers the row panels of x into the operator's input (NCCL), then the local matvec.  This is the
 * step the sharded Krylov solvers iterate: basis vectors, inner products (partial sums + one fused allreduce of the step's
 * scalars) and axpys are 1/N per rank. */
int32_t mpsk_op_apply_panel(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* x_panel, mpsk_tensor* y_panel);
/* panel over the left (axis 0) or right (axis 1) bond index; mpsk_tensor_row_panel == axis 0 */
int32_t mpsk_tensor_panel(mpsk_ctx* ctx, const mpsk_tensor* t, int32_t axis, int32_t rank, int32_t nranks, mpsk_tensor** out);
/* every rank passes its axis-0 panel; `full` (pre-allocated, same structure on all ranks) receives the whole tensor.  Stored
 * environments are kept as panels (1/N of the memory per rank) and re-assembled when they become the active GL / GR.
 * With a communicator, mpsk_transfer_left/right shard the update over the ket index of the new environment and mpsk_tsvd
 * distributes the coupled sectors over the ranks; both return complete, bit-identical tensors on every rank. */
int32_t mpsk_tensor_allgather(mpsk_ctx* ctx, const mpsk_tensor* panel, mpsk_tensor* full);

/* Streamed matvecs on HOST buffers: y_i = op(x_i) for a sequence of independent vectors whose blocks live in (pinned) host
 * memory -- the host-buffer form of `(h::MPO_∂∂AC2)(x)` (derivatives.jl:26-31,119-154) for callers that keep their vectors on
 * the host (block Krylov methods, `map(h, xs)`).  The H2D copy of x_{i+1}, the matvec of x_i and the D2H copy of y_{i-1} overlap
 * (two copy streams + the context stream, `depth` device slots).  x_like / y_like give the block structures (row panels with a
 * communicator: the semantics of mpsk_op_apply_panel).  Blocks must be contiguous, all row-major (host_layout 0, numpy) or all
 * column-major (host_layout 1, TensorKit).  submit returns as soon as the work is enqueued and blocks only while `depth`
 * submissions are in flight; a y buffer may be read after the `depth`-th following submit or after wait; an x buffer may be
 * overwritten under the same rule. */
int32_t mpsk_matvec_stream_create(mpsk_ctx* ctx, const mpsk_op* op, const mpsk_tensor* x_like, const mpsk_tensor* y_like, int32_t depth,
                                  int32_t host_layout, mpsk_matvec_stream** out);
int32_t mpsk_matvec_stream_submit(mpsk_ctx* ctx, mpsk_matvec_stream* s, const mpsk_blockref* xblocks, int32_t nx,
                                  const mpsk_blockref_mut* yblocks, int32_t ny);
int32_t mpsk_matvec_stream_wait(mpsk_ctx* ctx, mpsk_matvec_stream* s);
void mpsk_matvec_stream_free(mpsk_ctx* ctx, mpsk_matvec_stream* s);

/* ---- host-only numerical helper (exposed for the CPU test-suite) ------------------------------------ */
/* complex Schur decomposition H = Z T Z^H used by the Krylov-Schur eigensolver; H, Z: n x n column-major interleaved complex;
 * if 0 <= move_to <= move_from < n the eigenvalue at position move_from is moved to move_to afterwards. */
int32_t mpsk_host_schur(int32_t n, double* H, double* Z, int32_t move_from, int32_t move_to);
/* randomised self-test of the device allocator's free-list logic (split / coalesce / trim) with malloc as the backend; 0 = ok */
int32_t mpsk_host_alloc_selftest(uint64_t seed, int32_t nops);

#ifdef __cplusplus
}
#endif
#endif
