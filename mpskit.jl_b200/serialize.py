"""State (de)serialisation — SURVEY.md §8 row (f4): the reference has no on-disk format (users `serialize` Julia objects),
so this one is ours: a single `.npz` holding, per site, whichever gauge forms of the state are materialised
(`AL / AR / AC / CL`, the four lists behind `orthoview.jl:1-143`), each as its block list in ABI order plus the sector
tables needed to rebuild it.  Blocks move through the same `mpsk_tensor_download / _upload` calls as every other host
transfer; nothing is recomputed on load, so a state written mid-sweep resumes with the gauge centre where it was and the
reloaded tensors are bit-identical.

    save_mps(path, psi)            FiniteMPS -> file
    load_mps(path, ctx=None)       file -> FiniteMPS on `ctx`

`pack_tensor / unpack_tensor` are the host-side halves (numpy only) and are unit-tested on the CPU."""
from __future__ import annotations
import json
import numpy as np
from .spaces import Space

FORMAT_VERSION = 1
_LISTS = ("ALs", "ARs", "ACs", "CLs")


def pack_tensor(prefix, Vl: Space, mch, Vr: Space, dtype, blocks, table, store):
    """flatten one block tensor into `store` (name -> ndarray): header arrays + one contiguous data vector in ABI block
    order (`table` = [(m, l-sector, r-sector)])"""
    store[prefix + "/Vl"] = np.array([Vl.charges, Vl.dims], dtype=np.int64).reshape(2, -1)
    store[prefix + "/Vr"] = np.array([Vr.charges, Vr.dims], dtype=np.int64).reshape(2, -1)
    store[prefix + "/mch"] = np.array(list(mch), dtype=np.int64)
    parts = [np.ascontiguousarray(blocks[(m, jl)], dtype=dtype).reshape(-1) for (m, jl, jr) in table]
    store[prefix + "/data"] = np.concatenate(parts) if parts else np.zeros(0, dtype=dtype)


def unpack_tensor(prefix, store, table_fn, modulus=0):
    """inverse of pack_tensor: returns (Vl, mch, Vr, dtype, {(m, l-sector): ndarray}); `table_fn(Vl, mch, Vr)` gives the
    ABI block order"""
    vl, vr = store[prefix + "/Vl"], store[prefix + "/Vr"]
    Vl = Space(vl[0].tolist(), vl[1].tolist(), modulus)
    Vr = Space(vr[0].tolist(), vr[1].tolist(), modulus)
    mch = [int(c) for c in store[prefix + "/mch"]]
    data = store[prefix + "/data"]
    blocks, pos = {}, 0
    for (m, jl, jr) in table_fn(Vl, mch, Vr):
        n = Vl.dims[jl] * Vr.dims[jr]
        blocks[(m, jl)] = data[pos:pos + n].reshape(Vl.dims[jl], Vr.dims[jr])
        pos += n
    if pos != data.size:
        raise ValueError(f"{prefix}: {data.size} stored elements, block structure expects {pos}")
    return Vl, mch, Vr, data.dtype, blocks


def save_mps(path, psi):
    from .tensor import block_table
    store = {}
    present = {}
    for name in _LISTS:
        lst = getattr(psi, name)
        present[name] = [i for i, t in enumerate(lst) if t is not None]
        for i in present[name]:
            t = lst[i]
            table = block_table(t.Vl, t.mch, t.Vr)
            pack_tensor(f"{name}/{i}", t.Vl, t.mch, t.Vr, t.dtype, t.to_blocks(), table, store)
    meta = dict(version=FORMAT_VERSION, L=psi.L, dtype=str(psi.dtype), modulus=psi.pspace.modulus,
                pspace=[list(psi.pspace.charges), list(psi.pspace.dims)],
                V=[[list(v.charges), list(v.dims)] for v in psi.V], present=present)
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez(path, **store)


def load_mps(path, ctx=None):
    from .tensor import DeviceTensor, block_table
    from .finitemps import FiniteMPS
    from ._lib import default_context
    ctx = ctx or default_context()
    with np.load(path) as store:
        meta = json.loads(bytes(store["meta"]).decode())
        if meta["version"] != FORMAT_VERSION:
            raise ValueError(f"unsupported state file version {meta['version']}")
        mod = meta["modulus"]
        psi = object.__new__(FiniteMPS)
        psi.ctx, psi.L = ctx, meta["L"]
        psi.pspace = Space(meta["pspace"][0], meta["pspace"][1], mod)
        psi.pq = tuple(int(c) for c in psi.pspace.index_charges())
        psi.dtype = np.dtype(meta["dtype"])
        psi.V = [Space(c, d, mod) for c, d in meta["V"]]
        for name in _LISTS:
            lst = [None] * (psi.L + 1 if name == "CLs" else psi.L)
            for i in meta["present"][name]:
                Vl, mch, Vr, dt, blocks = unpack_tensor(f"{name}/{i}", store, block_table, mod)
                t = DeviceTensor.from_blocks(ctx, Vl, mch, Vr, blocks, dt)
                if name == "CLs":
                    t._is_bond = True
                lst[i] = t
            setattr(psi, name, lst)
    psi._views()
    return psi
