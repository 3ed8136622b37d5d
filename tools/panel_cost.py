"""offline cost model of the sharded matvec: tile-quantised GEMM work per rank for a given panel table"""
import sys, math
sys.path.insert(0, ".")
import mpskit_b200 as m
from mpskit_b200.workloads import gaussian_bond_space

def rank_cost(V, h, mch, bounds):
    """bounds: {sector j: (lo, hi)} rows of the bra space owned by this rank -> (tile work A, tile work C, exact flops/8)"""
    q = dict(zip(V.charges, V.dims)); idx = {c: j for j, c in enumerate(V.charges)}
    na = {V.charges[j]: hi - lo for j, (lo, hi) in bounds.items() if hi > lo}
    cl, cr = h.domcharges, h.imcharges
    tA = tC = fl = 0
    for qb, nb in q.items():
        M = sum(na.get(qb + c, 0) for c in cl)
        N = sum(q.get(qb + s, 0) for s in mch)
        if M and N:
            tA += math.ceil(M / 128) * math.ceil(N / 64) * 128 * 64 * nb; fl += M * N * nb
    for qa, n in na.items():
        for s in mch:
            nz = q.get(qa + s, 0)
            if not nz: continue
            K = sum(q.get(qa + s - c, 0) for c in cr)
            tC += math.ceil(nz / 128) * math.ceil(n / 64) * 128 * 64 * K; fl += nz * n * K
    return tA, tC, fl

def even_table(V, N):
    return {(r, j): ((n * r) // N, (n * (r + 1)) // N) for r in range(N) for j, n in enumerate(V.dims)}

if __name__ == "__main__":
    V = gaussian_bond_space(4096, 0)
    h = m.heisenberg_longrange_u1(2, r=6)[0]
    p = h.pcharges; mch = [p[a] + p[b] for a in range(2) for b in range(2)]
    full = rank_cost(V, h, mch, {j: (0, n) for j, n in enumerate(V.dims)})
    print("full: tile work", (full[0] + full[1]) / 1e9, "flops", full[2] / 1e9, "quantisation", (full[0] + full[1]) / full[2])
    for N in (2, 4, 8):
        for name, tab in (("even", even_table(V, N)), ("chunk", m.panel_table(V, N))):
            costs = []
            for r in range(N):
                b = {j: tab.get((r, j), (0, 0)) for j in range(len(V.dims))}
                tA, tC, fl = rank_cost(V, h, mch, b)
                costs.append((tA + tC, fl))
            mx = max(c[0] for c in costs)
            print(N, name, "max tile work / (full/N) =", round(mx / ((full[0] + full[1]) / N), 3), "flop shares", [round(c[1] / (full[2] / N), 2) for c in costs])
