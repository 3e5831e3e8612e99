"""ctypes binding of libmhdb200.so -- the same C ABI a Julia host reaches with `ccall`
(include/mhdb200.h; INTEGRATION.md shows the Julia side).  No torch types cross this boundary: only raw
pointers and sizes.  There is NO CPU fallback: if the library or a CUDA device is missing, calls raise."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MHDB200_LIBRARY") or os.path.join(_HERE, "csrc", "libmhdb200.so")  # same variable as the Julia binding

MHD_OK = 0
ERRORS = {-1: "MHD_E_INVALID", -2: "MHD_E_CUDA", -3: "MHD_E_STATE", -4: "MHD_E_CAPACITY", -5: "MHD_E_COMM", -6: "MHD_E_NOTCONV"}
FIELD_IDS = {"u": 0, "p": 1, "j": 2, "phi": 3}
CONVECTION = {"none": 0, "picard": 1, "newton": 2}
PRECOND = {"none": 0, "jacobi": 1, "block_tri": 2, "h1h1_blocks": 3}
UJ_SOLVER = {"gmres_jacobi": 0, "dense_lu": 1, "gmres_patch": 2}


class MhdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class mhd_mesh_t(C.Structure):
    _fields_ = [("nnodes", C.c_int64), ("coords", C.c_void_p), ("ncells", C.c_int64), ("cell_nodes", C.c_void_p),
                ("index_base", C.c_int32), ("cell_solid", C.c_void_p), ("cell_sigma", C.c_void_p)]


class mhd_tables_t(C.Structure):
    _fields_ = [("nq", C.c_int32), ("w", C.c_void_p), ("geo_grad", C.c_void_p), ("u_val", C.c_void_p),
                ("u_grad", C.c_void_p), ("p_val", C.c_void_p), ("j_val", C.c_void_p), ("j_div", C.c_void_p),
                ("phi_val", C.c_void_p)]


class mhd_layout_t(C.Structure):
    _fields_ = [("cell_dofs", C.c_void_p * 4), ("j_sign", C.c_void_p), ("nfree", C.c_int64 * 4),
                ("nowned", C.c_int64 * 4), ("ndir", C.c_int64 * 4), ("dir_values", C.c_void_p * 4),
                ("field_order", C.c_int32 * 4)]


class mhd_tables_h1h1_t(C.Structure):
    _fields_ = [("nq", C.c_int32), ("w", C.c_void_p), ("geo_grad", C.c_void_p), ("u_val", C.c_void_p),
                ("u_grad", C.c_void_p), ("p_val", C.c_void_p), ("phi_grad", C.c_void_p)]


class mhd_layout_h1h1_t(C.Structure):
    _fields_ = [("cell_dofs", C.c_void_p * 3), ("nfree", C.c_int64 * 3), ("nowned", C.c_int64 * 3),
                ("ndir", C.c_int64 * 3), ("dir_values", C.c_void_p * 3), ("field_order", C.c_int32 * 3)]


class mhd_params_t(C.Structure):
    _fields_ = [("alpha", C.c_double), ("beta", C.c_double), ("gamma", C.c_double), ("sigma", C.c_double),
                ("zeta_u", C.c_double), ("zeta_j", C.c_double), ("B", C.c_double * 3), ("f", C.c_double * 3),
                ("g", C.c_double * 3), ("convection", C.c_int32)]


class mhd_hunt_post_t(C.Structure):
    _fields_ = [("a", C.c_double), ("mu", C.c_double), ("sigma", C.c_double), ("grad_pz", C.c_double), ("Ha", C.c_double),
                ("nsums", C.c_int32), ("reserved", C.c_int32), ("u0", C.c_double), ("jscale", C.c_double)]


class mhd_solver_opts_t(C.Structure):
    _fields_ = [("m", C.c_int32), ("maxiter", C.c_int32), ("rtol", C.c_double), ("atol", C.c_double),
                ("precond", C.c_int32), ("uj_inner_its", C.c_int32), ("uj_inner_restart", C.c_int32),
                ("alpha_p", C.c_double), ("alpha_phi", C.c_double), ("uj_solver", C.c_int32), ("patch_its", C.c_int32),
                ("patch_omega", C.c_double)]


# every exported symbol of include/mhdb200.h with its signature (tests check the .so exports all of them)
_P = C.c_void_p
SIGNATURES = {
    "mhd_init": (C.c_int, [C.c_int]),
    "mhd_finalize": (C.c_int, []),
    "mhd_set_stream": (C.c_int, [_P]),
    "mhd_device_synchronize": (C.c_int, []),
    "mhd_last_error_string": (C.c_char_p, []),
    "mhd_comm_get_unique_id": (C.c_int, [_P]),
    "mhd_comm_init": (C.c_int, [C.c_int, C.c_int, _P]),
    "mhd_comm_finalize": (C.c_int, []),
    "mhd_operator_create": (C.c_int, [C.POINTER(mhd_mesh_t), C.POINTER(mhd_tables_t), C.POINTER(mhd_layout_t),
                                      C.POINTER(mhd_params_t), C.POINTER(_P)]),
    "mhd_operator_destroy": (C.c_int, [_P]),
    "mhd_operator_set_params": (C.c_int, [_P, C.POINTER(mhd_params_t)]),
    "mhd_operator_get_kernel_version": (C.c_int, [_P, _P]),
    "mhd_operator_set_deterministic": (C.c_int, [_P, C.c_int32, _P]),
    "mhd_operator_set_halo": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P]),
    "mhd_operator_halo_ipc_export": (C.c_int, [_P, _P]),
    "mhd_operator_halo_ipc_connect": (C.c_int, [_P, _P, _P, _P, _P]),
    "mhd_operator_halo_status": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mhd_operator_symbolic": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mhd_operator_get_csr": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "mhd_operator_get_scatter_stats": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mhd_jacobian": (C.c_int, [_P, _P, _P]),
    "mhd_residual": (C.c_int, [_P, _P, _P]),
    "mhd_residual_and_jacobian": (C.c_int, [_P, _P, _P]),
    "mhd_get_nzval": (C.c_int, [_P, _P]),
    "mhd_set_nzval": (C.c_int, [_P, _P]),
    "mhd_spmv": (C.c_int, [_P, _P, _P]),
    "mhd_dot": (C.c_int, [_P, _P, _P, _P]),
    "mhd_axpy": (C.c_int, [_P, C.c_double, _P, _P]),
    "mhd_multi_dot_axpy": (C.c_int, [_P, C.c_int32, _P, C.c_int64, _P, _P]),
    "mhd_solver_default_opts": (C.c_int, [C.POINTER(mhd_solver_opts_t)]),
    "mhd_solver_create": (C.c_int, [_P, C.POINTER(mhd_solver_opts_t), C.POINTER(_P)]),
    "mhd_solver_set_patches": (C.c_int, [_P, C.c_int64, _P, _P]),
    "mhd_solver_set_phi_patches": (C.c_int, [_P, C.c_int64, _P, _P]),
    "mhd_solver_setup": (C.c_int, [_P]),
    "mhd_solver_patch_apply": (C.c_int, [_P, _P, _P, C.c_double]),
    "mhd_solve": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_double), _P]),
    "mhd_solver_destroy": (C.c_int, [_P]),
    "mhd_operator_device_ptrs": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "mhd_kernel_launch_count": (C.c_int, [C.POINTER(C.c_int64)]),
    "mhd_fp64_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
    "mhd_hunt_error_norms": (C.c_int, [_P, _P, C.POINTER(mhd_tables_t), C.POINTER(mhd_hunt_post_t), C.POINTER(C.c_double)]),
    "mhd_map_entry_order": (C.c_int, [C.POINTER(C.c_uint16), C.POINTER(C.c_int64)]),
    "mhd_h1h1_operator_create": (C.c_int, [C.POINTER(mhd_mesh_t), C.POINTER(mhd_tables_h1h1_t), C.POINTER(mhd_layout_h1h1_t),
                                           C.POINTER(mhd_params_t), C.POINTER(_P)]),
    "mhd_h1h1_entry_order": (C.c_int, [C.POINTER(C.c_uint16), C.POINTER(C.c_int64)]),
    "mhd_profile_enable": (C.c_int, [C.c_int]),
    "mhd_profile_get": (C.c_int, [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mhd_profile_reset": (C.c_int, []),
}

_lib = None


def load():
    """dlopen the in-tree library (raises if it has not been built: `python __graft_entry__.py` builds it)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not built; run __graft_entry__.build() (no CPU fallback exists)")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != MHD_OK:
        raise MhdError(rc, load().mhd_last_error_string().decode())


_initialised = False


def init(device: int = 0):
    global _initialised
    check(load().mhd_init(device))
    _initialised = True


def finalize():
    global _initialised
    if _initialised:
        load().mhd_finalize()
        _initialised = False


def ptr(a) -> int:
    """Raw address of a numpy array, a torch tensor (host or CUDA) or an int address."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr()
    raise TypeError(type(a))


def profile_get(name: str):
    """(total device ms, launches) of the named kernel since the last reset."""
    ms, n = C.c_double(), C.c_int64()
    check(load().mhd_profile_get(name.encode(), C.byref(ms), C.byref(n)))
    return ms.value, n.value


def launch_count() -> int:
    n = C.c_int64()
    check(load().mhd_kernel_launch_count(C.byref(n)))
    return n.value
