"""ctypes binding of libmdiff.so (the C ABI declared in include/mdiff.h).

The library is the product: there is no Python/torch fallback.  Importing this module without the built
shared object raises, and every wrapper raises ``MdiffError`` with the library's message on a non-zero return.
"""
import ctypes as C
import os
from pathlib import Path

_ROOT = Path(__file__).resolve().parent
LIB_PATH = _ROOT / "_lib" / "libmdiff.so"


class MdiffError(RuntimeError):
    pass


def _load():
    if not LIB_PATH.exists():
        raise MdiffError(
            f"{LIB_PATH} is missing: build it with `python -m morphablediffusion_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / PyTorch fallback for the hot path.")
    return C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)


lib = _load()


class ConvGemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p),
        ("B", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("Cin", C.c_int), ("Cpitch", C.c_int),
        ("Wt", C.c_void_p),
        ("N", C.c_int),
        ("ntaps", C.c_int),
        ("tap", (C.c_int * 3) * 27),
        ("OD", C.c_int), ("OH", C.c_int), ("OW", C.c_int),
        ("os", C.c_int * 3), ("op", C.c_int * 3),
        ("bias", C.c_void_p),
        ("rowvec", C.c_void_p),
        ("rowvec_ld", C.c_int),
        ("res_f32", C.c_void_p),
        ("res_bf16", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("out_bf16", C.c_void_p),
        ("ldo", C.c_int),
        ("act", C.c_int),
        ("out_scale", C.c_float),
        ("BN", C.c_int),
    ]


lib.md_version.restype = C.c_int
lib.md_last_error.restype = C.c_char_p
lib.md_launch_count.restype = C.c_longlong
lib.md_reset_launch_count.restype = None
lib.md_op_conv_gemm.argtypes = [C.POINTER(ConvGemmArgs), C.c_void_p]
lib.md_op_conv_gemm.restype = C.c_int


def check(rc, what=""):
    if rc != 0:
        raise MdiffError(f"{what}: {lib.md_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "libmdiff takes contiguous CUDA tensors"
    return t.data_ptr()


def cur_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


ACT = {"none": 0, "silu": 1, "relu": 2, "geglu": 3, "gelu": 4}


def conv_gemm(A, Wt, *, B, D, H, W, Cin, N, taps, bias=None, rowvec=None, res_f32=None, res_bf16=None,
              out_f32=None, out_bf16=None, act="none", Cpitch=0, out_dims=None, os_=None, op=None, ldo=0,
              out_scale=1.0, BN=0):
    a = ConvGemmArgs()
    a.A = ptr(A); a.B, a.D, a.H, a.W = B, D, H, W
    a.Cin, a.Cpitch = Cin, Cpitch
    a.Wt = ptr(Wt); a.N = N
    a.ntaps = len(taps)
    for i, t in enumerate(taps):
        for j in range(3):
            a.tap[i][j] = t[j]
    if out_dims:
        a.OD, a.OH, a.OW = out_dims
    if os_:
        for j in range(3):
            a.os[j] = os_[j]
    if op:
        for j in range(3):
            a.op[j] = op[j]
    a.bias = ptr(bias); a.rowvec = ptr(rowvec)
    a.rowvec_ld = rowvec.shape[-1] if rowvec is not None else 0
    a.res_f32 = ptr(res_f32); a.res_bf16 = ptr(res_bf16)
    a.out_f32 = ptr(out_f32); a.out_bf16 = ptr(out_bf16)
    a.ldo = ldo; a.act = ACT[act]; a.out_scale = out_scale; a.BN = BN
    check(lib.md_op_conv_gemm(C.byref(a), cur_stream()), "md_op_conv_gemm")
