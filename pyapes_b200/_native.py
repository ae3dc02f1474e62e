"""ctypes binding of the C ABI declared in include/pyapes_b200.h.

The shared library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a) into
`pyapes_b200/lib/libpyapes_b200.so`.  There is no CPU fallback: if the library is missing, or
no CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PA_LIB") or os.path.join(_HERE, "lib", "libpyapes_b200.so")

PA_F32, PA_F64 = 0, 1
PA_MAX_OPS, PA_MAX_FACES = 4, 6
BC_KIND = {"dirichlet": 1, "neumann": 2, "symmetry": 3, "periodic": 4}
OP_STAR, OP_DIV_CENTRAL_FIELD, OP_DIV_UPWIND_FIELD, OP_DIV_UPWINDFD_FIELD = 0, 1, 2, 3
METHOD = {"cg": 0, "bicgstab": 1, "jacobi": 2}
RUNNING, CONVERGED, MAXIT, BAD_TOL = 0, 1, 2, 3
FLAG_CONTRACT = 1


class FaceBC(C.Structure):
    _fields_ = [
        ("axis", C.c_int32),
        ("side", C.c_int32),
        ("kind", C.c_int32),
        ("reserved", C.c_int32),
        ("value", C.c_double),
        ("values", C.c_void_p),
    ]


class Grid(C.Structure):
    _fields_ = [
        ("n", C.c_int32 * 3),
        ("lo", C.c_int32 * 3),
        ("hi", C.c_int32 * 3),
        ("gn0", C.c_int32),
        ("goff0", C.c_int32),
        ("olo0", C.c_int32),
        ("ohi0", C.c_int32),
        ("ndim", C.c_int32),
        ("reserved", C.c_int32),
    ]


class Op(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("has_param", C.c_int32),
        ("sign", C.c_double),
        ("param", C.c_double),
        ("coef", C.c_double * 3 * 3 * 3),
        ("adv", C.c_void_p),
        ("two_dx", C.c_double * 3),
        ("dx", C.c_double * 3),
        ("zero_am_lo", C.c_int32 * 3),
        ("zero_ap_hi", C.c_int32 * 3),
        ("param_field", C.c_void_p),
        ("edge", C.c_int32),
        ("adv_is_iterate", C.c_int32),
        ("adv_const", C.c_double),
        ("coef_tab", C.c_void_p * 3),
    ]


class Equation(C.Structure):
    _fields_ = [("nops", C.c_int32), ("reserved", C.c_int32), ("ops", Op * PA_MAX_OPS)]


class Report(C.Structure):
    _fields_ = [
        ("itr", C.c_int32),
        ("status", C.c_int32),
        ("tol", C.c_double),
        ("result_in_alt", C.c_int32),
        ("launches", C.c_int32),
        ("swaps", C.c_int32),
        ("reserved", C.c_int32),
    ]


class SolverCfg(C.Structure):
    _fields_ = [
        ("tol", C.c_double),
        ("max_it", C.c_int32),
        ("check_every", C.c_int32),
        ("use_graph", C.c_int32),
        ("variant", C.c_int32),
        ("flags", C.c_int32),
        ("reserved", C.c_int32),
    ]


# every symbol include/pyapes_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "pa_last_error": (C.c_char_p, []),
    "pa_abi_version": (C.c_int, []),
    "pa_device_count": (C.c_int, []),
    "pa_stencil_apply": (C.c_int, [C.POINTER(Grid), C.POINTER(Equation), C.c_int, _P, _P, _P]),
    "pa_grad_apply": (C.c_int, [C.POINTER(Grid), C.POINTER(Op), C.c_int, _P, _P, _P]),
    "pa_bc_apply": (C.c_int, [C.POINTER(Grid), C.c_int, C.POINTER(FaceBC), C.c_int, _P, _P]),
    "pa_solver_workspace_bytes": (C.c_size_t, [C.POINTER(Grid), C.c_int, C.c_int]),
    "pa_cg_solve": (C.c_int, [C.POINTER(Grid), C.POINTER(Equation), C.c_int, C.POINTER(FaceBC), C.c_int,
                              _P, _P, _P, C.POINTER(SolverCfg), _P, C.c_size_t, C.POINTER(Report), _P]),
    "pa_comm_unique_id": (C.c_int, [_P]),
    "pa_comm_create": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "pa_comm_destroy": (C.c_int, [_P]),
    "pa_cg_solve_dist": (C.c_int, [C.POINTER(Grid), C.POINTER(Equation), C.c_int, C.POINTER(FaceBC), C.c_int,
                                   _P, _P, _P, C.POINTER(SolverCfg), _P, C.c_size_t, _P, C.c_int, C.c_int,
                                   C.POINTER(Report), _P]),
    "pa_p2p_local_handle": (C.c_int, [_P]),
    "pa_p2p_attach": (C.c_int, [_P, C.c_int, C.c_int]),
    "pa_p2p_enabled": (C.c_int, []),
    "pa_p2p_halo_cap": (C.c_longlong, []),
    "pa_p2p_set_halo_cap": (C.c_int, [C.c_longlong]),
    "pa_p2p_disable": (C.c_int, []),
    "pa_solve_dist": (C.c_int, [C.c_int, C.POINTER(Grid), C.POINTER(Equation), C.c_int, C.POINTER(FaceBC), C.c_int,
                                _P, _P, _P, C.POINTER(SolverCfg), _P, C.c_size_t, _P, C.c_int, C.c_int,
                                C.POINTER(Report), _P]),
    "pa_euler_steps_dist": (C.c_int, [C.POINTER(Grid), C.POINTER(Equation), C.c_int, C.POINTER(FaceBC), C.c_int,
                                      _P, _P, _P, C.c_double, C.c_int, C.POINTER(C.c_int), _P, C.c_int, C.c_int, _P]),
    "pa_halo_exchange": (C.c_int, [C.POINTER(Grid), C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P]),
    "pa_cg_profile": (C.c_int, [C.POINTER(Grid), C.POINTER(Equation), C.c_int, C.POINTER(FaceBC), C.c_int,
                                _P, _P, _P, C.c_int, C.c_int, _P, C.c_size_t, C.POINTER(C.c_double), _P]),
    "pa_bicgstab_solve": (C.c_int, [C.POINTER(Grid), C.POINTER(Equation), C.c_int, C.POINTER(FaceBC), C.c_int,
                                    _P, _P, _P, C.POINTER(SolverCfg), _P, C.c_size_t, C.POINTER(Report), _P]),
    "pa_jacobi_solve": (C.c_int, [C.POINTER(Grid), C.POINTER(Equation), C.c_int, C.POINTER(FaceBC), C.c_int,
                                  _P, _P, _P, C.POINTER(SolverCfg), _P, C.c_size_t, C.POINTER(Report), _P]),
    "pa_euler_step": (C.c_int, [C.POINTER(Grid), C.POINTER(Equation), C.c_int, C.POINTER(FaceBC), C.c_int,
                                _P, _P, _P, C.c_double, _P]),
    "pa_euler_steps": (C.c_int, [C.POINTER(Grid), C.POINTER(Equation), C.c_int, C.POINTER(FaceBC), C.c_int,
                                 _P, _P, _P, C.c_double, C.c_int, C.POINTER(C.c_int), _P]),
    "pa_axpy": (C.c_int, [C.c_int, C.c_longlong, C.c_double, _P, _P, _P, _P]),
    "pa_cg_solve_host": (C.c_int, [C.POINTER(Grid), C.POINTER(Equation), C.c_int, C.POINTER(FaceBC), C.c_int,
                                   _P, _P, C.POINTER(SolverCfg), C.POINTER(Report)]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the CUDA library once; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"pyapes_b200: {LIB_PATH} is missing. Build it with `python -c 'import "
                "__graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        if handle.pa_abi_version() != 2:
            raise NativeError("pyapes_b200: ABI version mismatch, rebuild the library")
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().pa_last_error().decode()
        if rc == -3:
            raise NotImplementedError(f"pyapes_b200: {msg}")
        raise NativeError(f"pyapes_b200 native call failed ({rc}): {msg}")


def dtype_code(t) -> int:
    import torch

    if t == torch.float64:
        return PA_F64
    if t == torch.float32:
        return PA_F32
    raise TypeError(f"pyapes_b200: unsupported dtype {t}; only float32/float64 (backend.py:28-42)")


def require_cuda(t, what: str = "tensor") -> None:
    if not t.is_cuda:
        raise NativeError(
            f"pyapes_b200: {what} lives on {t.device}; the finite-difference path only runs on CUDA "
            "(no CPU fallback). Create the Mesh with device='cuda'."
        )
    if not t.is_contiguous():
        raise NativeError(f"pyapes_b200: {what} must be contiguous")


def current_stream(device) -> C.c_void_p:
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
