"""Restatement of the reference contractions (Appendix A).

Dense index orders (C-order numpy arrays): A[l,p,r], O[wl,p_out,p_in,wr], GL[l_bra,w,l_ket],
GR[r_ket,w,r_bra], C[l,r], x2[l,p1,r,p2].  Two independent implementations:

 * ``*_dense``  : one einsum per reference expression on dense embeddings;
 * ``*_sparse`` : the reference's own loop structure over ``keys`` of the SparseMPOSlice with
   per-entry pairwise contractions in the NCON label order of the source
   (derivatives.jl:77-104,119-154,171-193; transfer.jl:166-259), on dense arrays, and
 * ``dAC2_blocks``: the same reference-order algorithm on symmetry blocks (what TensorKit
   executes: per coupled sector one gemm) -- this is the CPU baseline that bench.py times.
"""
from __future__ import annotations
import numpy as np
from .blocks import BlockTensor
from .mpo import SparseMPOSlice


# ------------------------------------------------------------------ dense single expressions
def transfer_left_dense(v, O, A, Ab):
    """transfer.jl:105-107  v'[a' w'; b'] = v[1 2;4] A[4 5;b'] O[2 3;5 w'] conj(Ab[1 3;a'])"""
    return np.einsum('abd,dec,bfeg,afh->hgc', v, A, O, np.conj(Ab), optimize=True)


def transfer_right_dense(v, O, A, Ab):
    """transfer.jl:108-110  v'[a w; b] = A[a 2;1] O[w 4;2 3] conj(Ab[b 4;5]) v[1 3;5]"""
    return np.einsum('xba,ydbc,zde,ace->xyz', A, O, np.conj(Ab), v, optimize=True)


def transfer_left_bond(v, A, Ab):
    """transfer.jl:18-25 (N=1): v[bra;ket] -> v'[bra';ket']"""
    return np.einsum('ab,bpc,apd->dc', v, A, np.conj(Ab), optimize=True)


def transfer_right_bond(v, A, Ab):
    """transfer.jl:38-45 (N=1): v[ket;bra]"""
    return np.einsum('apb,bc,dpc->ad', A, v, np.conj(Ab), optimize=True)


def dAC_dense(x, O, GL, GR):
    """derivatives.jl:95-99"""
    return np.einsum('aed,dbx,esbc,xcz->asz', GL, x, O, GR, optimize=True)


def dAC2_dense(x2, O1, O2, GL, GR):
    """derivatives.jl:155-158"""
    return np.einsum('Agf,fexc,gsed,dtcb,xbz->Aszt', GL, x2, O1, O2, GR, optimize=True)


def dC_dense(C, GL, GR):
    """derivatives.jl:171-193 summed over levels"""
    return np.einsum('acb,bd,dcz->az', GL, C, GR, optimize=True)


# ------------------------------------------------------------------ reference loop structure
def transfer_left_sparse(vec, ham: SparseMPOSlice, A, Ab):
    """transfer.jl:166-211 (serial branch). vec: GL[l_bra, w, l_ket] with width-1 levels."""
    out = np.zeros((Ab.shape[2], ham.odim, A.shape[2]), dtype=np.result_type(vec, A, np.complex128))
    for k in range(ham.odim):
        for j in ham.keys_col(k):
            if ham.isscal(j, k):
                out[:, k, :] += ham.Os[j][k] * transfer_left_bond(vec[:, j, :], A, Ab)
            else:
                out[:, k, :] += np.einsum('ad,dec,fe,afh->hc', vec[:, j, :], A, ham.op(j, k),
                                          np.conj(Ab), optimize=True)
        # empty column: the reference contracts a zero tensor (transfer.jl:182-184) -> zeros
    return out


def transfer_right_sparse(vec, ham: SparseMPOSlice, A, Ab):
    """transfer.jl:212-259 (serial branch). vec: GR[r_ket, w, r_bra]."""
    out = np.zeros((A.shape[0], ham.odim, Ab.shape[0]), dtype=np.result_type(vec, A, np.complex128))
    for j in range(ham.odim):
        for k in ham.keys_row(j):
            if ham.isscal(j, k):
                out[:, j, :] += ham.Os[j][k] * transfer_right_bond(vec[:, k, :], A, Ab)
            else:
                out[:, j, :] += np.einsum('xba,db,zde,ae->xz', A, ham.op(j, k), np.conj(Ab),
                                          vec[:, k, :], optimize=True)
    return out


def dAC_sparse(x, H: SparseMPOSlice, GL, GR):
    """derivatives.jl:77-93 (serial branch) with the per-entry contraction order of 95-104:
    x*GR_j (label 1), then the operator (2,3), then GL_i (4,5)."""
    y = np.zeros((GL.shape[0], x.shape[1], GR.shape[2]), dtype=np.result_type(x, GL, np.complex128))   # bra spaces (rectangular for ac_proj)
    for (i, j) in H.keys():
        t = np.einsum('dbx,xz->dbz', x, GR[:, j, :])
        if H.isscal(i, j):
            t = H.Os[i][j] * t
        else:
            t = np.einsum('sb,dbz->dsz', H.op(i, j), t)
        y += np.einsum('ad,dsz->asz', GL[:, i, :], t)
    return y


def dAC2_sparse(x2, h1: SparseMPOSlice, h2: SparseMPOSlice, GL, GR):
    """derivatives.jl:119-154: stage 1 builds hl[j] over keys(h1,:,j); stage 2 sums over keys(h2).
    Per-entry NCON order: label 1 first (GL with the operator), then x (derivatives.jl:128-132)."""
    Dl, d1, Dr, d2 = x2.shape
    hl = [None] * h1.odim
    for j in range(h1.odim):
        cur = None
        for i in h1.keys_col(j):
            if h1.isscal(i, j):
                t = np.einsum('ab,bucq->aucq', h1.Os[i][j] * GL[:, i, :], x2)
            else:
                glo = np.einsum('ab,su->asbu', GL[:, i, :], h1.op(i, j))
                t = np.einsum('asbu,bucq->ascq', glo, x2)
            cur = t if cur is None else cur + t
        hl[j] = cur
    y = None
    for (j, k) in h2.keys():
        if hl[j] is None:
            continue
        if h2.isscal(j, k):
            t = np.einsum('asct,cz->aszt', h2.Os[j][k] * hl[j], GR[:, k, :])
        else:
            ho = np.einsum('asct,Tt->ascT', hl[j], h2.op(j, k))
            t = np.einsum('ascT,cz->aszT', ho, GR[:, k, :])
        y = t if y is None else y + t
    return y


# ------------------------------------------------------------------ block-sparse reference order
def env_blocks(G: BlockTensor, level):
    """{l-sector index: block} of one level of an environment BlockTensor."""
    return {jl: b for (m, jl), b in G.blocks.items() if m == level}


def dAC2_blocks(x: BlockTensor, h1: SparseMPOSlice, h2: SparseMPOSlice, GL: BlockTensor, GR: BlockTensor):
    """Block-sparse dAC2 in the reference's entry loop and contraction order (derivatives.jl:119-154).

    x : [l, (p1 p2), r] blocks;  GL: [l_bra, level, l_ket] (mch = -c_i);  GR: [r_ket, level, r_bra]
    (mch = +c_k).  For every h1 entry the operator is first fused with the env block (label 1),
    then multiplied with x (labels 2, 3): one gemm per coupled sector of size
    (n_a * d) x (n_b * d) -- the d-times-more-than-optimal cost the survey notes.  Returns y with
    the structure of x.  Used as the CPU baseline (`cpu_baseline.kind = "port"`).
    """
    d = h1.d
    Vl, Vr = x.Vl, x.Vr
    Vb = GL.Vl                      # bra space: x.Vl, or a row panel of it (multi-GPU sharding tests)
    lpos = {c: i for i, c in enumerate(Vl.charges)}
    rpos = {c: i for i, c in enumerate(Vr.charges)}
    # hl[j][(ja, s, u_t... )]: keep as dict keyed by (ja, s, t, jc) -> (n_a, n_c) arrays
    hl = [None] * h1.odim
    for j in range(h1.odim):
        cur = None
        for i in h1.keys_col(j):
            t = {}
            scal = h1.isscal(i, j)
            O = None if scal else h1.op(i, j)
            for (lev, ja), g in GL.blocks.items():
                if lev != i:
                    continue
                qa = Vb.charges[ja]
                qb = qa + GL.mch[i]
                jb = lpos.get(qb)
                if jb is None:
                    continue
                if scal:
                    gs = h1.Os[i][j] * g
                    for u in range(d):
                        for tt in range(d):
                            xb = x.blocks.get((u * d + tt, jb))
                            if xb is None:
                                continue
                            jc = rpos[qb + x.mch[u * d + tt]]
                            key = (ja, u, tt, jc)
                            t[key] = t.get(key, 0) + gs @ xb
                else:
                    # TensorKit fuses the env block with the operator first (label 1) and then does one
                    # gemm per coupled sector; for abelian sectors that is one gemm per symmetry-allowed
                    # non-zero operator element (s,u):  (O[s,u] * GL_i) @ x[u,t]
                    for s_ in range(d):
                        for u in range(d):
                            if O[s_, u] == 0:
                                continue
                            gs = O[s_, u] * g
                            for tt in range(d):
                                xb = x.blocks.get((u * d + tt, jb))
                                if xb is None:
                                    continue
                                jc = rpos[qb + x.mch[u * d + tt]]
                                key = (ja, s_, tt, jc)
                                t[key] = t.get(key, 0) + gs @ xb
            if cur is None:
                cur = t
            else:
                for key, v in t.items():
                    cur[key] = cur.get(key, 0) + v
        hl[j] = cur
    y = BlockTensor(Vb, x.mch, Vr, np.result_type(x.dtype, GL.dtype, np.complex128))
    for (j, k) in h2.keys():
        if hl[j] is None:
            continue
        scal = h2.isscal(j, k)
        O = None if scal else h2.op(j, k)
        for (ja, s, tt, jc), blk in hl[j].items():
            g = GR.blocks.get((k, jc))
            if g is None:
                continue
            jz = rpos[Vr.charges[jc] + GR.mch[k]]
            if scal:
                key = (s * d + tt, ja)
                if key in y.blocks and y.blocks[key].shape[1] == g.shape[1]:
                    y.blocks[key] += (h2.Os[j][k] * blk) @ g
            else:
                for T in range(d):
                    if O[T, tt] == 0:
                        continue
                    key = (s * d + T, ja)
                    if key in y.blocks:
                        y.blocks[key] += (O[T, tt] * blk) @ g
    return y
