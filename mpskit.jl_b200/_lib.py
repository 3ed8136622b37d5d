"""ctypes binding of libmpskit_b200.so (include/mpskit_b200.h).  There is no CPU fallback: if the
shared library is missing or no sm_100 GPU is present every product call fails loudly."""
from __future__ import annotations
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpskit_b200.so")

OK, EINVAL, ESPACE, ECUDA, ENOMEM, EINTERNAL, ENODEVICE = 0, -1, -2, -3, -4, -5, -6
F64, C128 = 0, 1
SR, LM, LR = 0, 1, 2
NOTRUNC, TRUNCDIM, TRUNCERR, TRUNCBELOW = 0, 1, 2, 3
OP_DC, OP_DAC, OP_DAC2 = 0, 1, 2
FLAG_DEBUG_GEMM, FLAG_EXACT_SVD, FLAG_FAST_SVD_ALL, FLAG_REPLICATE_LOCAL, FLAG_SVD_RELATIVE, FLAG_GEMM_4M = 1, 2, 4, 8, 16, 32
OP_SHARDED = 0x100


class MpskError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mpskit_b200 error {code}: {msg}")
        self.code = code


class SpaceMismatch(MpskError):
    """mirror of TensorKit.SpaceMismatch (MPSK_ESPACE)"""


class c_space(C.Structure):
    _fields_ = [("nsec", C.c_int32), ("charge", C.POINTER(C.c_int32)), ("dim", C.POINTER(C.c_int64)), ("modulus", C.c_int32)]


class c_blockref(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride_l", C.c_int64), ("stride_r", C.c_int64)]


class c_trunc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("value", C.c_double)]


class c_info(C.Structure):
    _fields_ = [("converged", C.c_int32), ("numiter", C.c_int32), ("numops", C.c_int32), ("normres", C.c_double)]


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_SP = C.POINTER(c_space)
_I32P = C.POINTER(C.c_int32)
_I64P = C.POINTER(C.c_int64)
_DP = C.POINTER(C.c_double)

# name -> (restype, argtypes); must list every symbol the header declares (tests check the exports)
PROTOTYPES = {
    "mpsk_version": (C.c_int32, []),
    "mpsk_create": (C.c_int32, [C.c_int32, C.c_uint64, _PP]),
    "mpsk_destroy": (None, [_P]),
    "mpsk_last_error": (C.c_char_p, [_P]),
    "mpsk_set_stream": (C.c_int32, [_P, _P]),
    "mpsk_synchronize": (C.c_int32, [_P]),
    "mpsk_launch_count": (C.c_int64, [_P]),
    "mpsk_bytes_in_use": (C.c_int64, [_P]),
    "mpsk_block_table": (C.c_int32, [_SP, C.c_int32, _I32P, _SP, _I32P, _I32P, _I32P, _I32P]),
    "mpsk_tensor_upload": (C.c_int32, [_P, _SP, C.c_int32, _I32P, _SP, C.c_int32, C.POINTER(c_blockref), C.c_int32, _PP]),
    "mpsk_tensor_zeros": (C.c_int32, [_P, _SP, C.c_int32, _I32P, _SP, C.c_int32, _PP]),
    "mpsk_tensor_download": (C.c_int32, [_P, _P, C.POINTER(c_blockref), C.c_int32]),
    "mpsk_tensor_stack_labels": (C.c_int32, [_P, _PP, C.c_int32, _PP]),
    "mpsk_tensor_select_labels": (C.c_int32, [_P, _P, _I32P, C.c_int32, _PP]),
    "mpsk_tensor_adjoint": (C.c_int32, [_P, _P, _PP]),
    "mpsk_tensor_fill_random": (C.c_int32, [_P, _P, C.c_uint64]),
    "mpsk_tensor_copy": (C.c_int32, [_P, _P, _PP]),
    "mpsk_tensor_free": (None, [_P, _P]),
    "mpsk_tensor_structure": (C.c_int32, [_P, _I32P, _I32P, _I32P, _I32P]),
    "mpsk_tensor_spaces": (C.c_int32, [_P, _I32P, _I64P, _I32P, _I64P, _I32P]),
    "mpsk_dot": (C.c_int32, [_P, _P, _P, _DP, _DP]),
    "mpsk_norm": (C.c_int32, [_P, _P, _DP]),
    "mpsk_axpy": (C.c_int32, [_P, C.c_double, C.c_double, _P, _P]),
    "mpsk_scale": (C.c_int32, [_P, C.c_double, C.c_double, _P]),
    "mpsk_mposlice_create": (C.c_int32, [_P, C.c_int32, _I32P, _I32P, C.c_int32, _I32P, _I32P, _DP, _DP, _PP]),
    "mpsk_mposlice_create_zn": (C.c_int32, [_P, C.c_int32, _I32P, _I32P, C.c_int32, _I32P, _I32P, _DP, _DP, C.c_int32, _PP]),
    "mpsk_mposlice_free": (None, [_P, _P]),
    "mpsk_mposlice_keys": (C.c_int32, [_P, _I32P, _I32P]),
    "mpsk_op_create": (C.c_int32, [_P, C.c_int32, _P, _P, _P, _P, _P, _PP]),
    "mpsk_op_free": (None, [_P, _P]),
    "mpsk_op_flops": (C.c_double, [_P]),
    "mpsk_op_time_gemms": (C.c_int32, [_P, _P, C.c_int32, _DP]),
    "mpsk_op_time_mix": (C.c_int32, [_P, _P, C.c_int32, _DP, _DP]),
    "mpsk_op_apply": (C.c_int32, [_P, _P, _P, _P]),
    "mpsk_eigsolve": (C.c_int32, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, _DP, _DP, _P, C.POINTER(c_info)]),
    "mpsk_eigsolve_values": (C.c_int32, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, _DP, _I32P, _I32P,
                                         C.POINTER(c_info)]),
    "mpsk_expv": (C.c_int32, [_P, _P, C.c_double, C.c_double, _P, C.c_int32, C.c_int32, C.c_double, _P, C.POINTER(c_info)]),
    "mpsk_linsolve": (C.c_int32, [_P, _P, _P, _P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_double, _P,
                                  C.POINTER(c_info)]),
    "mpsk_tmop_create": (C.c_int32, [_P, C.c_int32, C.c_int32, _PP, _PP, _PP, _P, _P, _P, _PP]),
    "mpsk_multiop_create": (C.c_int32, [_P, C.c_int32, _PP, C.c_int32, _PP]),
    "mpsk_transfer_left": (C.c_int32, [_P, _P, _P, _P, _P, _PP]),
    "mpsk_transfer_right": (C.c_int32, [_P, _P, _P, _P, _P, _PP]),
    "mpsk_contract_ac2": (C.c_int32, [_P, _P, _P, _PP]),
    "mpsk_mul_bond_right": (C.c_int32, [_P, _P, _P, _PP]),
    "mpsk_mul_bond_left": (C.c_int32, [_P, _P, _P, _PP]),
    "mpsk_tsvd": (C.c_int32, [_P, _P, C.c_int32, _I32P, c_trunc, _PP, _PP, _PP, _DP]),
    "mpsk_tsvd_spectrum": (C.c_int32, [_P, _I32P, _I32P, _I32P, _I32P, _DP]),
    "mpsk_bond_inverse": (C.c_int32, [_P, _P, _PP]),
    "mpsk_leftnull": (C.c_int32, [_P, _P, _PP]),
    "mpsk_rightnull": (C.c_int32, [_P, _P, _PP]),
    "mpsk_contract_adjoint_left": (C.c_int32, [_P, _P, _P, _PP]),
    "mpsk_tensor_cat": (C.c_int32, [_P, _P, _P, C.c_int32, _PP]),
    "mpsk_qrpos": (C.c_int32, [_P, _P, _PP, _PP]),
    "mpsk_lqpos": (C.c_int32, [_P, _P, _PP, _PP]),
    "mpsk_comm_unique_id": (C.c_int32, [C.POINTER(C.c_uint8)]),
    "mpsk_comm_init": (C.c_int32, [_P, C.POINTER(C.c_uint8), C.c_int32, C.c_int32]),
    "mpsk_allreduce_sum": (C.c_int32, [_P, _P]),
    "mpsk_tensor_row_panel": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, _PP]),
    "mpsk_tensor_scatter_panel": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, _P]),
    "mpsk_op_apply_local": (C.c_int32, [_P, _P, _P, _P]),
    "mpsk_op_apply_panel": (C.c_int32, [_P, _P, _P, _P]),
    "mpsk_tensor_panel": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _PP]),
    "mpsk_tensor_allgather": (C.c_int32, [_P, _P, _P]),
    "mpsk_host_schur": (C.c_int32, [C.c_int32, _DP, _DP, C.c_int32, C.c_int32]),
    "mpsk_host_alloc_selftest": (C.c_int32, [C.c_uint64, C.c_int32]),
    "mpsk_matvec_stream_create": (C.c_int32, [_P, _P, _P, _P, C.c_int32, C.c_int32, _PP]),
    "mpsk_matvec_stream_submit": (C.c_int32, [_P, _P, C.POINTER(c_blockref), C.c_int32, C.POINTER(c_blockref), C.c_int32]),
    "mpsk_matvec_stream_wait": (C.c_int32, [_P, _P]),
    "mpsk_matvec_stream_free": (None, [_P, _P]),
}

_lib = None


def load():
    """dlopen the in-tree shared library and attach prototypes (no GPU needed for this)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class Context:
    """one library context per process / rank (mpsk_create)."""

    def __init__(self, device=0, flags=0):
        self.lib = load()
        h = C.c_void_p()
        st = self.lib.mpsk_create(int(device), int(flags), C.byref(h))
        if st == ENODEVICE:
            raise MpskError(st, "no sm_100 GPU available -- mpskit_b200 has no CPU fallback")
        if st != OK:
            raise MpskError(st, "mpsk_create failed")
        self.h = h
        self.device = device
        self.rank, self.nranks = 0, 1            # set by sharding.init_communicator

    def check(self, st):
        if st != OK:
            msg = self.lib.mpsk_last_error(self.h).decode()
            raise (SpaceMismatch if st == ESPACE else MpskError)(st, msg)

    def set_stream(self, cuda_stream_ptr):
        self.check(self.lib.mpsk_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self.check(self.lib.mpsk_synchronize(self.h))

    def launch_count(self):
        return int(self.lib.mpsk_launch_count(self.h))

    def bytes_in_use(self):
        return int(self.lib.mpsk_bytes_in_use(self.h))

    def close(self):
        if self.h:
            self.lib.mpsk_destroy(self.h)
            self.h = None


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        dev = int(os.environ.get("LOCAL_RANK", "0"))
        flags = FLAG_DEBUG_GEMM if os.environ.get("MPSK_DEBUG_GEMM") == "1" else 0
        if os.environ.get("MPSK_GEMM_4M") == "1":
            flags |= FLAG_GEMM_4M
        _default_ctx = Context(dev, flags)
    return _default_ctx
