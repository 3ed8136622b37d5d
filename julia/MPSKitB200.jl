# MPSKitB200.jl -- thin ccall shim that routes MPSKit.jl's hot path through libmpskit_b200 (include/mpskit_b200.h).
# Written against MPSKit v0.10.2 / TensorKit 0.12; NOT runnable in the build image (no Julia there) -- see INTEGRATION.md.
# No algorithm file of MPSKit is edited: the shim only adds more specific methods (multiple dispatch is the seam).
module MPSKitB200

using MPSKit, TensorKit, KrylovKit
using MPSKit: MPO_∂∂C, MPO_∂∂AC, MPO_∂∂AC2, SparseMPOSlice, MPSTensor, MPOTensor, MPSBondTensor

const LIB = joinpath(@__DIR__, "..", "mpskit.jl_b200", "libmpskit_b200.so")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

struct CSpace
    nsec::Int32; charge::Ptr{Int32}; dim::Ptr{Int64}; modulus::Int32
end
struct BlockRef
    ptr::Ptr{Cvoid}; stride_l::Int64; stride_r::Int64
end
struct CInfo
    converged::Int32; numiter::Int32; numops::Int32; normres::Float64
end

function check(st)
    st == 0 && return
    msg = unsafe_string(ccall((:mpsk_last_error, LIB), Cstring, (Ptr{Cvoid},), CTX[]))
    st == -2 ? throw(SpaceMismatch(msg)) : error("mpskit_b200 error $st: $msg")
end

function __init__()
    h = Ref{Ptr{Cvoid}}()
    st = ccall((:mpsk_create, LIB), Int32, (Int32, UInt64, Ref{Ptr{Cvoid}}), 0, 0, h)
    st == 0 || error("mpskit_b200: no B200 available (status $st); there is no CPU fallback")
    CTX[] = h[]
end

# ---- TensorKit adapter: the ONLY place that knows TensorKit's storage conventions -------------------------------
# U1Irrep charge -> Int32 (2q); sectors sorted ascending for the ABI (TensorKit iterates |q| ascending, + before -).
charge2(c::U1Irrep) = Int32(round(2 * c.charge)); charge2(::Trivial) = Int32(0)
function cspace(V)
    secs = sort(collect(sectors(V)); by=charge2)
    q = Int32[isdual(V) ? -charge2(c) : charge2(c) for c in secs]
    p = sortperm(q)
    return q[p], Int64[dim(V, c) for c in secs][p]
end

# A TensorMap is handed over block by block.  TensorKit 0.12 stores one column-major matrix per coupled sector;
# `t[f1, f2]` is a StridedView of the sub-block of the fusion-tree pair (f1, f2) into that matrix, so its pointer and
# strides are exactly what the ABI wants -- no data is copied or permuted on the Julia side.
#
# MPS tensor A[l p; r]:  f1.uncoupled = (q_l, s_p), f2.uncoupled = (q_r,).  The ABI labels are the physical *basis states*
# (a charge sector of degeneracy k contributes k labels), blocks are ordered label-major, l-sector-minor.
function blockrefs(A::MPSTensor)
    P = space(A, 2)
    labels = Tuple{eltype(sectors(P)),Int}[]                  # (charge, index inside the degeneracy space)
    for c in sort(collect(sectors(P)); by=charge2), k in 1:dim(P, c)
        push!(labels, (c, k))
    end
    Vl = space(A, 1)
    lsecs = sort(collect(sectors(Vl)); by=charge2)
    refs = BlockRef[]
    for (c, k) in labels, ql in lsecs
        for (f1, f2) in fusiontrees(A)
            f1.uncoupled == (ql, c) || continue
            sv = A[f1, f2]                                    # size (n_l, dim(P, c), n_r)
            st = strides(sv)
            push!(refs, BlockRef(pointer(sv) + (k - 1) * st[2] * sizeof(eltype(A)), st[1], st[3]))
        end
    end
    return refs
end
# two-site tensors x[l p1; r p2] (labels (p1, p2), stride_r taken from the r leg of the domain), bond tensors C[l; r]
# (one label) and environment levels GL_i[l_bra w; l_ket] (label = MPO level, w one-dimensional) follow the same pattern;
# MPSKit keeps environments as Vectors of such tensors, which map onto `mpsk_tensor_stack_labels`.
function blockrefs(x::MPOTensor) end
function blockrefs(C::MPSBondTensor) end

function upload(t::AbstractTensorMap, mch::Vector{Int32}, Vl, Vr)
    ql, nl = cspace(Vl); qr, nr = cspace(Vr)
    refs = blockrefs(t)
    out = Ref{Ptr{Cvoid}}()
    GC.@preserve t ql nl qr nr refs mch begin
        sl = CSpace(length(ql), pointer(ql), pointer(nl), 0); sr = CSpace(length(qr), pointer(qr), pointer(nr), 0)
        check(ccall((:mpsk_tensor_upload, LIB), Int32,
                    (Ptr{Cvoid}, Ref{CSpace}, Int32, Ptr{Int32}, Ref{CSpace}, Int32, Ptr{BlockRef}, Int32, Ref{Ptr{Cvoid}}),
                    CTX[], sl, length(mch), mch, sr, scalartype(t) <: Complex ? 1 : 0, refs, length(refs), out))
    end
    return out[]
end
download!(t::AbstractTensorMap, h) = (refs = blockrefs(t); GC.@preserve t refs check(ccall((:mpsk_tensor_download, LIB), Int32,
                                      (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{BlockRef}, Int32), CTX[], h, refs, length(refs))); t)

# SparseMPOSlice -> mpsk_mposlice (kinds: 0 absent / 1 scalar (Os[i,j] isa E) / 2 operator), sparseslice.jl:40-106
function slice_handle(h::SparseMPOSlice) end

# ---- the overloads ------------------------------------------------------------------------------------------------
# fixedpoint(A, x0, which, alg) -- src/algorithms/fixedpoint.jl:9-30, called from dmrg.jl:35,94,110, vumps.jl:81-89
function MPSKit.fixedpoint(A::MPO_∂∂AC2, x0::MPOTensor, which::Symbol, alg::Union{Lanczos,Arnoldi})
    GL = upload_env(A.leftenv); GR = upload_env(A.rightenv); x = upload_ac2(x0)
    op = Ref{Ptr{Cvoid}}()
    check(ccall((:mpsk_op_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                CTX[], 2, slice_handle(A.o1), slice_handle(A.o2), GL, GR, x, op))
    vec = zeros_like(x); re = Ref(0.0); im_ = Ref(0.0); info = Ref(CInfo(0, 0, 0, 0.0))
    check(ccall((:mpsk_eigsolve, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Int32, Int32, Float64, Ref{Float64}, Ref{Float64}, Ptr{Cvoid}, Ref{CInfo}),
                CTX[], op[], x, which === :SR ? 0 : which === :LM ? 1 : 2, 1, alg.krylovdim, alg.maxiter, alg.tol, re, im_, vec, info))
    info[].converged == 0 && @warn "fixedpoint not converged after $(info[].numiter) iterations: normres = $(info[].normres)"
    return complex(re[], im_[]), download!(similar(x0), vec)
end
# identical methods for MPO_∂∂AC (kind 1) and MPO_∂∂C (kind 0); MPSKit.integrate(f::DerivativeOperator, y0, t, dt, alg)
# -> mpsk_expv (integrators.jl:20-25); (h::MPO_∂∂AC2)(x) -> mpsk_op_apply (derivatives.jl:30).

# transfer_left / transfer_right over a SparseMPOSlice -- src/transfermatrix/transfer.jl:137-140, FinEnv.jl:122,138
function MPSKit.transfer_left(vec::AbstractVector{V}, ham::SparseMPOSlice, A::V, Ab::V) where {V<:MPSTensor}
    out = Ref{Ptr{Cvoid}}()
    check(ccall((:mpsk_transfer_left, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                CTX[], slice_handle(ham), upload_mps(A), upload_mps(Ab), upload_env(vec), out))
    return download_env(out[], vec)
end

# DMRG2 step tail -- dmrg.jl:96-103: tsvd!(ac2; trunc, alg=SVD()), normalize!(c)
function device_tsvd(ac2h, d1, p1charges, trunc) end   # -> mpsk_tsvd + mpsk_tensor_structure/_spaces (two-phase size query)

# Whole sweeps stay on the device (state, environments and Krylov basis are device-resident between calls):
# MPSKit.find_groundstate!(ψ::FiniteMPS, H::MPOHamiltonian, alg::DMRG2, envs) uploads ψ once, runs the sweep body of
# dmrg.jl:80-137 with the calls above on handles (exactly what mpskit.jl_b200/dmrg.py does), and downloads ψ at the end.
end # module
