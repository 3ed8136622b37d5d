# MPSKitB200.jl -- thin ccall shim that routes MPSKit.jl's hot path through libmpskit_b200 (include/mpskit_b200.h).
# Written against MPSKit v0.10.2 / TensorKit 0.12 (abelian sectors: Trivial, U1Irrep).  The build image of this
# repository has no Julia toolchain, so this file has NOT been executed there; every ccall below is exercised from
# Python through ctypes with the identical argument lists (mpskit.jl_b200/_lib.py, tests/).  See INTEGRATION.md.
# No algorithm file of MPSKit is edited: the shim only adds more specific methods (multiple dispatch is the seam).
#
# Residency: device handles are cached per host tensor OBJECT (a WeakKeyDict keyed by the TensorMap), so environments
# and site tensors that MPSKit re-uses between calls (FinEnv keeps the same Vector of environment tensors while a site
# is untouched, FinEnv.jl:114-145) are uploaded once, not once per `fixedpoint`.
module MPSKitB200

using MPSKit, TensorKit, KrylovKit
using MPSKit: MPO_∂∂C, MPO_∂∂AC, MPO_∂∂AC2, SparseMPOSlice, MPSTensor, MPOTensor, MPSBondTensor

const LIB = joinpath(@__DIR__, "..", "mpskit.jl_b200", "libmpskit_b200.so")

struct CSpace
    nsec::Int32; charge::Ptr{Int32}; dim::Ptr{Int64}; modulus::Int32
end
struct BlockRef
    ptr::Ptr{Cvoid}; stride_l::Int64; stride_r::Int64
end
struct CInfo
    converged::Int32; numiter::Int32; numops::Int32; normres::Float64
end
struct CTrunc
    kind::Int32; value::Float64
end

# ---- context: the library serialises calls on one context (every ABI entry locks it), so MPSKit's concurrent tasks
# (vumps.jl:39-45, tdvp.jl:27-38, mpohaminfenv.jl:65-68) may share it; one context per device.
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function __init__()
    h = Ref{Ptr{Cvoid}}()
    st = ccall((:mpsk_create, LIB), Int32, (Int32, UInt64, Ref{Ptr{Cvoid}}), 0, 0, h)
    st == 0 || error("mpskit_b200: no B200 available (status $st); there is no CPU fallback")
    CTX[] = h[]
end

function check(st)
    st == 0 && return
    msg = unsafe_string(ccall((:mpsk_last_error, LIB), Cstring, (Ptr{Cvoid},), CTX[]))
    st == -2 ? throw(SpaceMismatch(msg)) : error("mpskit_b200 error $st: $msg")
end

# device handle with a finalizer
mutable struct Handle
    p::Ptr{Cvoid}
    function Handle(p::Ptr{Cvoid})
        h = new(p)
        finalizer(x -> (x.p == C_NULL || ccall((:mpsk_tensor_free, LIB), Cvoid, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], x.p); x.p = C_NULL), h)
        return h
    end
end
const RESIDENT = WeakKeyDict{Any,Handle}()          # host tensor object -> device copy (uploaded once)

# ---- TensorKit adapter: the ONLY place that knows TensorKit's storage conventions -------------------------------
# U1Irrep charge -> Int32 (2q); the ABI lists sectors in ascending charge (TensorKit iterates |q| ascending, + before -).
charge2(c::U1Irrep) = Int32(round(2 * c.charge)); charge2(::Trivial) = Int32(0)
secs(V) = sort(collect(sectors(V)); by=charge2)                 # sectors as they appear in fusion trees of that leg
function cspace(V)
    s = secs(V)
    return Int32[charge2(c) for c in s], Int64[dim(V, c) for c in s]
end
# flat label list of a (physical or MPO-level) leg: one label per basis state, label charge = charge of its sector
labels(P) = [(c, k) for c in secs(P) for k in 1:dim(P, c)]

# Per-block pointers + strides straight into TensorKit's storage: `t[f1, f2]` is a strided view of the fusion-tree pair's
# sub-block inside the coupled-sector matrix, so nothing is copied or permuted on the Julia side.
#
# three-leg tensors T[l m; r] (MPS tensors A[l p; r], environment levels GL[l_bra w; l_ket], GR[r_ket w; r_bra]):
# conservation q_l + q_m = q_r in TensorKit's convention, i.e. the ABI label charge is charge2 of the sector of leg 2 as
# it appears in the fusion tree (environment MPO legs are dual spaces, FinEnv.jl:49-55, which gives -c_w for GL).
function blockrefs(A::AbstractTensorMap{S,2,1}) where {S}
    lab = labels(space(A, 2)); ls = secs(space(A, 1))
    trees = collect(fusiontrees(A))
    refs = BlockRef[]
    for (c, k) in lab, ql in ls
        for (f1, f2) in trees
            f1.uncoupled == (ql, c) || continue
            sv = A[f1, f2]                                    # size (n_l, dim(P, c), n_r)
            st = strides(sv)
            push!(refs, BlockRef(pointer(sv) + (k - 1) * st[2] * sizeof(eltype(A)), st[1], st[3]))
        end
    end
    return refs
end
mcharges(A::AbstractTensorMap{S,2,1}) where {S} = Int32[charge2(c) for (c, _) in labels(space(A, 2))]

# bond tensors C[l; r]: one label of charge 0
function blockrefs(C::AbstractTensorMap{S,1,1}) where {S}
    trees = collect(fusiontrees(C))
    refs = BlockRef[]
    for ql in secs(space(C, 1)), (f1, f2) in trees
        f1.uncoupled == (ql,) || continue
        sv = C[f1, f2]; st = strides(sv)
        push!(refs, BlockRef(pointer(sv), st[1], st[2]))
    end
    return refs
end
mcharges(::AbstractTensorMap{S,1,1}) where {S} = Int32[0]

# transfer-matrix vectors v[l; e r] with a one-dimensional charged leg e in the domain (transfer_spectrum, toolbox.jl:44-46):
# one label; conservation q_l = q_e + q_r with the domain sectors as they appear in the fusion tree f2, i.e. the ABI label
# charge (q_r = q_l + m) is -charge2 of that sector
function blockrefs(v::AbstractTensorMap{S,1,2}) where {S}
    trees = collect(fusiontrees(v))
    refs = BlockRef[]
    for ql in secs(space(v, 1)), (f1, f2) in trees
        f1.uncoupled == (ql,) || continue
        sv = v[f1, f2]; st = strides(sv)                      # size (n_l, 1, n_r)
        push!(refs, BlockRef(pointer(sv), st[1], st[3]))
    end
    return refs
end
mcharges(v::AbstractTensorMap{S,1,2}) where {S} = Int32[-charge2(first(fusiontrees(v))[2].uncoupled[1])]

# two-site tensors x[l p1; r p2] (dmrg.jl:92): labels (p1, p2) with p2 fastest; the p2 leg sits in the domain, where its
# sectors appear conjugated (P'), so its contribution to the label charge is -charge2 of the sector in f2.uncoupled[2]
function blockrefs(x::AbstractTensorMap{S,2,2}) where {S}
    l1 = labels(space(x, 2)); l2 = labels(space(x, 4)); ls = secs(space(x, 1))
    l2 = sort(l2; by=t -> (-charge2(t[1]), t[2]))             # ascending physical charge s_p2 = -charge2(sector of P')
    trees = collect(fusiontrees(x))
    refs = BlockRef[]
    for (c1, k1) in l1, (c2, k2) in l2, ql in ls
        for (f1, f2) in trees
            (f1.uncoupled == (ql, c1) && f2.uncoupled[2] == c2) || continue
            sv = x[f1, f2]                                    # size (n_l, dim p1, n_r, dim p2)
            st = strides(sv)
            push!(refs, BlockRef(pointer(sv) + ((k1 - 1) * st[2] + (k2 - 1) * st[4]) * sizeof(eltype(x)), st[1], st[3]))
        end
    end
    return refs
end
function mcharges(x::AbstractTensorMap{S,2,2}) where {S}
    l1 = labels(space(x, 2)); l2 = sort(labels(space(x, 4)); by=t -> (-charge2(t[1]), t[2]))
    return Int32[charge2(c1) - charge2(c2) for (c1, _) in l1 for (c2, _) in l2]
end

left_space(t) = space(t, 1)
right_space(t::AbstractTensorMap{S,2,1}) where {S} = space(t, 3)'      # domain leg: undo the dual of `space(t, i)`
right_space(t::AbstractTensorMap{S,1,1}) where {S} = space(t, 2)'
right_space(t::AbstractTensorMap{S,2,2}) where {S} = space(t, 3)'
right_space(t::AbstractTensorMap{S,1,2}) where {S} = space(t, 3)'

function upload(t::AbstractTensorMap)
    ql, nl = cspace(left_space(t)); qr, nr = cspace(right_space(t))
    refs = blockrefs(t); mch = mcharges(t)
    out = Ref{Ptr{Cvoid}}()
    GC.@preserve t ql nl qr nr refs mch begin
        sl = CSpace(length(ql), pointer(ql), pointer(nl), 0); sr = CSpace(length(qr), pointer(qr), pointer(nr), 0)
        check(ccall((:mpsk_tensor_upload, LIB), Int32,
                    (Ptr{Cvoid}, Ref{CSpace}, Int32, Ptr{Int32}, Ref{CSpace}, Int32, Ptr{BlockRef}, Int32, Ref{Ptr{Cvoid}}),
                    CTX[], sl, length(mch), mch, sr, scalartype(t) <: Complex ? 1 : 0, refs, length(refs), out))
    end
    return Handle(out[])
end
resident(t::AbstractTensorMap) = get!(() -> upload(t), RESIDENT, t)
function download!(t::AbstractTensorMap, h::Handle)
    refs = blockrefs(t)
    GC.@preserve t refs check(ccall((:mpsk_tensor_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{BlockRef}, Int32),
                                    CTX[], h.p, refs, length(refs)))
    return t
end
function zeros_like(h::Handle)
    out = Ref{Ptr{Cvoid}}()
    check(ccall((:mpsk_tensor_copy, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), CTX[], h.p, out))
    check(ccall((:mpsk_scale, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, Ptr{Cvoid}), CTX[], 0.0, 0.0, out[]))
    return Handle(out[])
end

# environments: MPSKit keeps a Vector of per-level tensors (FinEnv.jl:9-37); the matvec wants ONE tensor whose labels are
# the MPO levels -> upload every level (cached) and stack them on the device
function upload_env(levels::AbstractVector)
    return get!(RESIDENT, levels) do
        hs = [resident(l) for l in levels]
        ptrs = Ptr{Cvoid}[h.p for h in hs]
        out = Ref{Ptr{Cvoid}}()
        GC.@preserve hs ptrs check(ccall((:mpsk_tensor_stack_labels, LIB), Int32, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int32, Ref{Ptr{Cvoid}}),
                                         CTX[], ptrs, length(ptrs), out))
        Handle(out[])
    end
end
function download_env(h::Handle, like::AbstractVector)
    outv = similar.(like)
    for (i, t) in enumerate(outv)
        sel = Ref{Ptr{Cvoid}}(); lab = Int32[i - 1]
        check(ccall((:mpsk_tensor_select_labels, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), CTX[], h.p, lab, 1, sel))
        download!(t, Handle(sel[]))
    end
    return outv
end

# SparseMPOSlice -> mpsk_mposlice (kinds: 0 absent / 1 scalar (Os[i,j] isa E) / 2 operator), sparseslice.jl:40-106.
# Width-1 levels: level charge = the single sector of domspaces[i]; imspaces are the adjoints (sparsempo.jl:16-17).
const SLICES = WeakKeyDict{Any,Ptr{Cvoid}}()
function slice_handle(h::SparseMPOSlice)
    return get!(SLICES, h) do
        od = h.odim
        P = h.pspace; lab = labels(P); d = length(lab)
        all(dim(V) == 1 for V in h.domspaces) || error("mpskit_b200: MPO levels wider than one must be flattened first")
        cl = Int32[charge2(first(sectors(V))) for V in h.domspaces]
        cr = Int32[-charge2(first(sectors(V))) for V in h.imspaces]
        pch = Int32[charge2(c) for (c, _) in lab]
        kind = zeros(Int32, od * od); scal = zeros(Float64, 2 * od * od); ops = Float64[]
        for i in 1:od, j in 1:od
            e = h.Os[i, j]
            idx = (i - 1) * od + j
            if e isa Number
                iszero(e) && continue
                kind[idx] = 1; scal[2idx - 1] = real(e); scal[2idx] = imag(e)
            else
                kind[idx] = 2
                O = zeros(ComplexF64, d, d)                    # O[p_out, p_in] in the label basis
                for (f1, f2) in fusiontrees(e)                 # e[wl p_out; p_in wr]
                    blk = e[f1, f2]                            # size (1, dim(c_out), dim(c_in), 1)
                    co, ci = f1.uncoupled[2], f2.uncoupled[1]
                    ro = findall(t -> t[1] == co, lab); ri = findall(t -> t[1] == ci, lab)
                    O[ro, ri] .= reshape(blk, length(ro), length(ri))
                end
                append!(ops, reinterpret(Float64, vec(permutedims(O))))     # row-major d x d, interleaved re/im
            end
        end
        out = Ref{Ptr{Cvoid}}()
        check(ccall((:mpsk_mposlice_create, LIB), Int32,
                    (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                    CTX[], od, cl, cr, d, pch, kind, scal, isempty(ops) ? zeros(2) : ops, out))
        out[]
    end
end

# ---- the overloads ------------------------------------------------------------------------------------------------
which_code(w::Symbol) = w === :SR ? Int32(0) : w === :LM ? Int32(1) : Int32(2)

function device_op(kind::Integer, o1, o2, leftenv, rightenv, x::Handle)
    GL = upload_env(leftenv); GR = upload_env(rightenv)
    op = Ref{Ptr{Cvoid}}()
    check(ccall((:mpsk_op_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                CTX[], kind, o1 === nothing ? C_NULL : slice_handle(o1), o2 === nothing ? C_NULL : slice_handle(o2), GL.p, GR.p, x.p, op))
    return op[]
end
free_op(op) = ccall((:mpsk_op_free, LIB), Cvoid, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], op)

function device_eigsolve(op, x::Handle, which::Symbol, alg)
    vec = zeros_like(x); re = Ref(0.0); im_ = Ref(0.0); info = Ref(CInfo(0, 0, 0, 0.0))
    herm = alg isa Lanczos ? Int32(1) : Int32(0)          # Arnoldi (Defaults.eigsolver, defaults.jl:33) -> Krylov-Schur on the device
    check(ccall((:mpsk_eigsolve, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Int32, Int32, Float64, Ref{Float64}, Ref{Float64}, Ptr{Cvoid}, Ref{CInfo}),
                CTX[], op, x.p, which_code(which), herm, alg.krylovdim, alg.maxiter, alg.tol, re, im_, vec.p, info))
    info[].converged == 0 && @warn "fixedpoint not converged after $(info[].numiter) iterations: normres = $(info[].normres)"   # fixedpoint.jl:12-14
    return complex(re[], im_[]), vec
end

# fixedpoint(A, x0, which, alg) -- src/algorithms/fixedpoint.jl:9-30, called from dmrg.jl:35,94,110, vumps.jl:81-89
function MPSKit.fixedpoint(A::MPO_∂∂AC2, x0::MPOTensor, which::Symbol, alg::Union{Lanczos,Arnoldi})
    x = upload(x0); op = device_op(2, A.o1, A.o2, A.leftenv, A.rightenv, x)
    val, vec = device_eigsolve(op, x, which, alg); free_op(op)
    return val, download!(similar(x0), vec)
end
function MPSKit.fixedpoint(A::MPO_∂∂AC, x0::MPSTensor, which::Symbol, alg::Union{Lanczos,Arnoldi})
    x = upload(x0); op = device_op(1, A.o, nothing, A.leftenv, A.rightenv, x)
    val, vec = device_eigsolve(op, x, which, alg); free_op(op)
    return val, download!(similar(x0), vec)
end
function MPSKit.fixedpoint(A::MPO_∂∂C, x0::MPSBondTensor, which::Symbol, alg::Union{Lanczos,Arnoldi})
    x = upload(x0); op = device_op(0, nothing, nothing, A.leftenv, A.rightenv, x)
    val, vec = device_eigsolve(op, x, which, alg); free_op(op)
    return val, download!(similar(x0), vec)
end

# (h::MPO_∂∂AC2)(x) -- derivatives.jl:26-31 (what KrylovKit calls when the Krylov loop stays in Julia)
function (h::MPO_∂∂AC2)(x0::MPOTensor)
    x = upload(x0); op = device_op(2, h.o1, h.o2, h.leftenv, h.rightenv, x); y = zeros_like(x)
    check(ccall((:mpsk_op_apply, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), CTX[], op, x.p, y.p)); free_op(op)
    return download!(similar(x0), y)
end

# integrate(f, y0, t, dt, alg) -> exponentiate(f, -im*dt, y0) -- src/algorithms/timestep/integrators.jl:20-25
function device_expv(op, y0::Handle, dt, alg)
    y = zeros_like(y0); info = Ref(CInfo(0, 0, 0, 0.0)); tau = -im * dt
    check(ccall((:mpsk_expv, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Ptr{Cvoid}, Int32, Int32, Float64, Ptr{Cvoid}, Ref{CInfo}),
                CTX[], op, real(tau), imag(tau), y0.p, alg.krylovdim, alg.maxiter, alg.tol, y.p, info))
    info[].converged == 0 && @warn "integrate not converged"                                           # integrators.jl:22-23
    return y
end
function MPSKit.integrate(f::MPO_∂∂AC2, y0::MPOTensor, t::Number, dt::Number, alg::Lanczos)
    x = upload(y0); op = device_op(2, f.o1, f.o2, f.leftenv, f.rightenv, x); y = device_expv(op, x, dt, alg); free_op(op)
    return download!(similar(y0), y)
end
function MPSKit.integrate(f::MPO_∂∂AC, y0::MPSTensor, t::Number, dt::Number, alg::Lanczos)
    x = upload(y0); op = device_op(1, f.o, nothing, f.leftenv, f.rightenv, x); y = device_expv(op, x, dt, alg); free_op(op)
    return download!(similar(y0), y)
end
function MPSKit.integrate(f::MPO_∂∂C, y0::MPSBondTensor, t::Number, dt::Number, alg::Lanczos)
    x = upload(y0); op = device_op(0, nothing, nothing, f.leftenv, f.rightenv, x); y = device_expv(op, x, dt, alg); free_op(op)
    return download!(similar(y0), y)
end

# transfer_left / transfer_right over a SparseMPOSlice -- src/transfermatrix/transfer.jl:137-140, FinEnv.jl:122,138
function MPSKit.transfer_left(vec::AbstractVector{V}, ham::SparseMPOSlice, A::V, Ab::V) where {V<:MPSTensor}
    out = Ref{Ptr{Cvoid}}()
    check(ccall((:mpsk_transfer_left, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                CTX[], slice_handle(ham), resident(A).p, resident(Ab).p, upload_env(vec).p, out))
    res = download_env(Handle(out[]), vec)
    return res
end
function MPSKit.transfer_right(vec::AbstractVector{V}, ham::SparseMPOSlice, A::V, Ab::V) where {V<:MPSTensor}
    out = Ref{Ptr{Cvoid}}()
    check(ccall((:mpsk_transfer_right, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                CTX[], slice_handle(ham), resident(A).p, resident(Ab).p, upload_env(vec).p, out))
    return download_env(Handle(out[]), vec)
end

# DMRG2 step tail -- dmrg.jl:96-97: al, c, ar, eps = tsvd!(ac2; trunc, alg=SVD()); normalize!(c).
# Two-phase size query: the kept sector dimensions come back from the library (mpsk_tensor_structure / _spaces).
trunc_code(t::TensorKit.NoTruncation) = CTrunc(0, 0.0)
trunc_code(t::TensorKit.TruncationDimension) = CTrunc(1, Float64(t.dim))
trunc_code(t::TensorKit.TruncationError) = CTrunc(2, Float64(t.ϵ))
trunc_code(t::TensorKit.TruncationCutoff) = CTrunc(3, Float64(t.ϵ))
function device_space(h::Handle, side::Symbol)
    nl = Ref{Int32}(); nr = Ref{Int32}(); nm = Ref{Int32}(); dt = Ref{Int32}()
    ccall((:mpsk_tensor_structure, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Int32}), h.p, nl, nr, nm, dt)
    ql = zeros(Int32, nl[]); dl = zeros(Int64, nl[]); qr = zeros(Int32, nr[]); dr = zeros(Int64, nr[]); mc = zeros(Int32, nm[])
    ccall((:mpsk_tensor_spaces, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int64}, Ptr{Int32}, Ptr{Int64}, Ptr{Int32}), h.p, ql, dl, qr, dr, mc)
    q, n = side === :left ? (ql, dl) : (qr, dr)
    return U1Space(U1Irrep(c // 2) => k for (c, k) in zip(q, n))
end
function device_tsvd(ac2::MPOTensor, trunc)
    x = upload(ac2)
    P1 = space(ac2, 2); pch = Int32[charge2(c) for (c, _) in labels(P1)]
    U = Ref{Ptr{Cvoid}}(); S = Ref{Ptr{Cvoid}}(); Vh = Ref{Ptr{Cvoid}}(); eps = Ref(0.0)
    check(ccall((:mpsk_tsvd, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, CTrunc, Ref{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}, Ref{Float64}),
                CTX[], x.p, length(pch), pch, trunc_code(trunc), U, S, Vh, eps))
    hU, hS, hV = Handle(U[]), Handle(S[]), Handle(Vh[])
    Vn = device_space(hU, :right)                              # the truncated bond space decided on the device
    al = TensorMap(undef, scalartype(ac2), space(ac2, 1) ⊗ P1, Vn)
    c = TensorMap(undef, scalartype(ac2), Vn, Vn)
    ar = TensorMap(undef, scalartype(ac2), Vn ⊗ space(ac2, 4)', space(ac2, 3)')     # [k p2; r] = _transpose_front(V⁺)
    download!(al, hU); download!(c, hS); download!(ar, hV)
    return al, c, ar, eps[]
end

# Handle-level wrappers of the entries the bond-changing algorithms use (changebonds/optimalexpand.jl:26-28,84-95 and the
# edge updates of IDMRG2, idmrg.jl:126,139,168,174): unary / binary device operations returning a new handle.
function unary(sym::Symbol, a::Handle)
    out = Ref{Ptr{Cvoid}}()
    check(ccall((sym, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), CTX[], a.p, out))
    return Handle(out[])
end
device_leftnull(a::Handle) = unary(:mpsk_leftnull, a)                      # leftnull(AL)
device_rightnull(a::Handle) = unary(:mpsk_rightnull, a)                    # rightnull!(_transpose_tail(AR))
device_inv(c::Handle) = unary(:mpsk_bond_inverse, c)                       # inv(ψ.CR[i])
device_adjoint(a::Handle) = unary(:mpsk_tensor_adjoint, a)
function device_adjoint_left(n::Handle, x2::Handle)                        # adjoint(VL) * AC2
    out = Ref{Ptr{Cvoid}}()
    check(ccall((:mpsk_contract_adjoint_left, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), CTX[], n.p, x2.p, out))
    return Handle(out[])
end
function device_cat(a::Handle, b::Handle, axis::Integer)                   # catcodomain (axis 0) / catdomain (axis 1)
    out = Ref{Ptr{Cvoid}}()
    check(ccall((:mpsk_tensor_cat, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), CTX[], a.p, b.p, axis, out))
    return Handle(out[])
end
function TensorKit.leftnull(A::MPSTensor{<:GradedSpace{U1Irrep}})
    h = device_leftnull(resident(A))
    N = TensorMap(undef, scalartype(A), codomain(A), device_space(h, :right))
    return download!(N, h)
end

# VectorInterface on handles (KrylovKit vectors that stay on the device between calls): inner, norm, add!!, scale!!
function device_dot(a::Handle, b::Handle)
    re = Ref(0.0); im_ = Ref(0.0)
    check(ccall((:mpsk_dot, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), CTX[], a.p, b.p, re, im_))
    return complex(re[], im_[])
end
function device_norm(a::Handle)
    out = Ref(0.0)
    check(ccall((:mpsk_norm, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Float64}), CTX[], a.p, out))
    return out[]
end
function device_axpy!(y::Handle, α::Number, x::Handle)                       # y += α x
    check(ccall((:mpsk_axpy, LIB), Int32, (Ptr{Cvoid}, Float64, Float64, Ptr{Cvoid}, Ptr{Cvoid}), CTX[], real(α), imag(α), x.p, y.p))
    return y
end

# leftorth!(A; alg = QRpos()) / rightorth!(A; alg = LQpos()) -- orthoview.jl:49-60, ortho.jl:127-138, finitemps.jl:150
function device_qrpos(a::Handle)
    q = Ref{Ptr{Cvoid}}(); r = Ref{Ptr{Cvoid}}()
    check(ccall((:mpsk_qrpos, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}), CTX[], a.p, q, r))
    return Handle(q[]), Handle(r[])
end
function device_lqpos(a::Handle)
    l = Ref{Ptr{Cvoid}}(); q = Ref{Ptr{Cvoid}}()
    check(ccall((:mpsk_lqpos, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}), CTX[], a.p, l, q))
    return Handle(l[]), Handle(q[])
end
device_mul_right(a::Handle, c::Handle) = binary(:mpsk_mul_bond_right, a, c)   # AC = AL * C   (orthoview.jl:95-106)
device_mul_left(c::Handle, a::Handle) = binary(:mpsk_mul_bond_left, c, a)     # AC = C * AR
device_ac2(a1::Handle, a2::Handle) = binary(:mpsk_contract_ac2, a1, a2)       # dmrg.jl:92,108
function binary(sym::Symbol, a::Handle, b::Handle)
    out = Ref{Ptr{Cvoid}}()
    check(ccall((sym, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), CTX[], a.p, b.p, out))
    return Handle(out[])
end

# linsolve(flip(tm), b, x0, GMRES, 1, -1) -- src/environments/mpohaminfenv.jl:95,113,146,164: the regularised MPO-free
# transfer matrix of the identity levels as ONE device operator (regularize: transfermatrix.jl:68-76) and the library's GMRES.
# `left` = true is flip(tm) (v -> v * tm).  The levels with an operator on the diagonal pass their 1 x 1 slices in `slices`.
function device_transfer_linsolve(A::AbstractVector, Ab::AbstractVector, left::Bool, b::MPSBondTensor, x0::MPSBondTensor, alg,
                                  a0::Number, a1::Number; lvec=nothing, rvec=nothing, slices=nothing)
    n = length(A)
    pa = [resident(A[i]).p for i in 1:n]; pb = [resident(Ab[i]).p for i in 1:n]
    ps = slices === nothing ? C_NULL : [slice_handle(s) for s in slices]
    hb = upload(b); hx = upload(x0)
    # regularisation pair: lvec_adj = adjoint(lvec) in the vector space of b, rvec the subtracted direction
    hl = lvec === nothing ? C_NULL : device_adjoint(upload(lvec)).p
    hr = rvec === nothing ? C_NULL : upload(rvec).p
    op = Ref{Ptr{Cvoid}}()
    GC.@preserve pa pb ps check(ccall((:mpsk_tmop_create, LIB), Int32,
                                      (Ptr{Cvoid}, Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                                      CTX[], left ? 1 : 0, n, ps, pa, pb, hb.p, hl, hr, op))
    x = zeros_like(hx); info = Ref(CInfo(0, 0, 0, 0.0))
    check(ccall((:mpsk_linsolve, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Float64, Float64, Int32, Int32, Float64, Ptr{Cvoid}, Ref{CInfo}),
                CTX[], op[], hb.p, hx.p, real(a0), imag(a0), real(a1), imag(a1), alg.krylovdim, alg.maxiter, alg.tol, x.p, info))
    free_op(op[])
    info[].converged == 0 && @warn "linsolve failed to converge: normres = $(info[].normres)"          # mpohaminfenv.jl:97-98
    return download!(similar(x0), x), info[]
end

# transfer_spectrum(above; below, tol, num_vals, sector) -- src/algorithms/toolbox.jl:42-58: the reference calls
# eigsolve(flip(TransferMatrix(above.AL, below.AL)), init, num_vals, :LM; tol); here the unit-cell transfer operator is one
# device operator and the Krylov-Schur solve for all `num_vals` values runs inside the library.
function device_transfer_spectrum(above::InfiniteMPS, below::InfiniteMPS, init::AbstractTensorMap, num_vals::Integer, tol::Real;
                                  krylovdim::Integer=max(30, 2 * num_vals), maxiter::Integer=100)
    n = length(above)
    A = [resident(above.AL[i]).p for i in 1:n]; Ab = [resident(below.AL[i]).p for i in 1:n]
    v = upload(init)
    op = Ref{Ptr{Cvoid}}()
    GC.@preserve A Ab check(ccall((:mpsk_tmop_create, LIB), Int32,
                                  (Ptr{Cvoid}, Int32, Int32, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                                  CTX[], 1, n, C_NULL, A, Ab, v.p, C_NULL, C_NULL, op))       # left = 1: v -> v * tm (flip)
    vals = zeros(Float64, 2 * num_vals); nvals = Ref{Int32}(0); nconv = Ref{Int32}(0); info = Ref(CInfo(0, 0, 0, 0.0))
    check(ccall((:mpsk_eigsolve_values, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Int32, Int32, Float64, Ptr{Float64}, Ref{Int32}, Ref{Int32}, Ref{CInfo}),
                CTX[], op[], v.p, which_code(:LM), num_vals, krylovdim, maxiter, tol, vals, nvals, nconv, info))
    free_op(op[])
    nconv[] < num_vals && @warn "correlation length failed to converge: normres = $(info[].normres)"      # toolbox.jl:54-55
    return [complex(vals[2j - 1], vals[2j]) for j in 1:nvals[]]
end
function MPSKit.transfer_spectrum(above::InfiniteMPS{<:MPSTensor{<:GradedSpace{U1Irrep}}}; below=above, tol=MPSKit.Defaults.tol,
                                  num_vals=20, sector=first(sectors(oneunit(left_virtualspace(above, 1)))))
    init = randomize!(similar(above.AL[1], left_virtualspace(below, 0),
                              ℂ[typeof(sector)](sector => 1)' * left_virtualspace(above, 0)))     # toolbox.jl:44-46
    transferspace = fuse(left_virtualspace(above, 0) * left_virtualspace(below, 0)')
    num_vals = min(dim(transferspace, sector), num_vals)                                          # toolbox.jl:48-49
    # the one-dimensional sector leg becomes the label charge of the device vector (blockrefs of a 1 <- 2 tensor map)
    return device_transfer_spectrum(above, below, init, num_vals, tol)
end

# Whole sweeps can stay on the device (state, environments and Krylov basis resident between calls): the sweep body of
# dmrg.jl:80-137 with the calls above on handles is exactly mpskit.jl_b200/dmrg.py (find_groundstate_inplace); a Julia
# `find_groundstate!(ψ::FiniteMPS, H::MPOHamiltonian, alg::DMRG2, envs)` method built from `device_op`, `device_eigsolve`,
# `device_tsvd` and `mpsk_contract_ac2` uploads ψ once and downloads it at the end.
end # module
