"""ctypes wrapper of the C oracle (oracle/mhd_oracle.c). TEST INFRASTRUCTURE ONLY: imported by tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libmhd_oracle.so")
_FIELDS = ("u", "p", "j", "phi")
_LO = {"u": 0, "p": 81, "j": 85, "phi": 121}


class oracle_params_t(C.Structure):
    _fields_ = [("alpha", C.c_double), ("beta", C.c_double), ("gamma", C.c_double), ("sigma", C.c_double),
                ("zeta_u", C.c_double), ("zeta_j", C.c_double), ("B", C.c_double * 3), ("f", C.c_double * 3),
                ("g", C.c_double * 3), ("convection", C.c_int32)]


class oracle_tables_t(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w", "geo_grad", "u_val", "u_grad", "p_val", "j_val", "j_div", "phi_val")]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            import subprocess

            subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = C.CDLL(_PATH)
        _lib.oracle_dot.restype = C.c_double
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


class COracle:
    """Holds contiguous copies of the FE tables of a `FESpaces` and calls the C restatement."""

    def __init__(self, fes, prm, cells=None, threads=None):
        """`cells`: optional index array -- the oracle then only knows these cells (bounded parity samples).
        `threads`: OpenMP threads (default: the runtime's choice, i.e. OMP_NUM_THREADS or all cores)."""
        self.lib = load()
        if threads:
            self.lib.oracle_set_num_threads(int(threads))
        T = fes.tables
        self._t = [np.ascontiguousarray(a, dtype=np.float64) for a in (T.w, T.geo_grad, T.nu, T.dnu, T.pp, T.psi, T.dpsi, T.chi)]
        self.tab = oracle_tables_t(*[a.ctypes.data for a in self._t])
        p = oracle_params_t()
        p.alpha, p.beta, p.gamma, p.sigma, p.zeta_u, p.zeta_j = prm.alpha, prm.beta, prm.gamma, prm.sigma, prm.zeta_u, prm.zeta_j
        for i in range(3):
            p.B[i], p.f[i], p.g[i] = prm.B[i], prm.f[i], prm.g[i]
        p.convection = {"none": 0, "picard": 1, "newton": 2}[prm.convection]
        self.prm = p
        self.fes = fes
        self.coords = np.ascontiguousarray(fes.mesh.coords, dtype=np.float64)
        sel = slice(None) if cells is None else np.asarray(cells, dtype=np.int64)
        self.cell_nodes = np.ascontiguousarray(fes.mesh.cell_nodes[sel], dtype=np.int32)
        self.ncells = len(self.cell_nodes)
        gids = fes.cell_global_ids()[sel]
        # Dirichlet entries: -(index into the concatenated Dirichlet array + 1)
        dir_off, o = {}, 0
        for f in _FIELDS:
            dir_off[f] = o
            o += fes.ndir[f]
        g = gids.copy()
        for f in _FIELDS:
            ids = fes.cell_dofs[f][sel]
            sl = slice(_LO[f], _LO[f] + ids.shape[1])
            g[:, sl] = np.where(ids > 0, gids[:, sl], -(dir_off[f] + (-ids - 1)) - 1)
        self.gids = np.ascontiguousarray(g, dtype=np.int32)
        self.dirv = np.ascontiguousarray(np.concatenate([fes.dirichlet_values[f] for f in _FIELDS] + [np.zeros(1)]))
        self.jsign = np.ascontiguousarray(fes.j_sign[sel], dtype=np.int8)
        self.threads = self.lib.oracle_num_threads()

    def jacobian_values(self, x, rowptr, colval, c0=0, c1=None, out=None):
        c1 = self.ncells if c1 is None else c1
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        colval = np.ascontiguousarray(colval, dtype=np.int64)
        nz = np.zeros(len(colval)) if out is None else out
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.lib.oracle_assemble_jacobian(C.byref(self.tab), C.byref(self.prm), C.c_int64(c0), C.c_int64(c1), _p(self.coords),
                                          _p(self.cell_nodes), _p(self.gids), _p(self.jsign), _p(self.dirv), _p(x),
                                          _p(rowptr), _p(colval), _p(nz))
        return nz

    def residual(self, x, c0=0, c1=None):
        c1 = self.ncells if c1 is None else c1
        r = np.zeros(self.fes.ndofs)
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.lib.oracle_assemble_residual(C.byref(self.tab), C.byref(self.prm), C.c_int64(c0), C.c_int64(c1), _p(self.coords),
                                          _p(self.cell_nodes), _p(self.gids), _p(self.jsign), _p(self.dirv), _p(x), _p(r))
        return r

    def spmv(self, rowptr, colval, nzval, x):
        y = np.empty(len(rowptr) - 1)
        self.lib.oracle_spmv(C.c_int64(len(rowptr) - 1), _p(rowptr), _p(colval), _p(nzval), _p(x), _p(y))
        return y
