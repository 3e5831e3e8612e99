"""Parity sampler (TEST INFRASTRUCTURE: tests/ and the post-timing check of bench.py only, never the product path).

Compares what the device assembled -- CSR structure, matrix values, residual -- with the C oracle on a bounded block of
cells at ANY mesh size (the benchmark configuration included), and the device SpMV with a host product on the device's
own matrix.  A matrix row is checked when every cell around its dof lies inside the sampled block ("complete row"): the
oracle then holds the full row although it only integrated the block.

Numbering: the oracle works in the numbering of the FE spaces (per-field offsets, `FESpaces.cell_global_ids`), the
library in [owned fields | ghost fields]; `lib_of_fes` maps the former to the latter (identity on one GPU).
"""
from __future__ import annotations

import numpy as np

from . import mhd_oracle as O
from .c_oracle import COracle


def sample_block(cell_nodes: np.ndarray, target: int = 2048, seed_cell: int = 0) -> np.ndarray:
    """A connected block of >= `target` cells grown ring by ring (cells sharing a node) from `seed_cell`."""
    nc = len(cell_nodes)
    target = min(target, nc)
    nn = int(cell_nodes.max()) + 1
    in_set = np.zeros(nc, dtype=bool)
    in_set[seed_cell] = True
    while in_set.sum() < target:
        node_mark = np.zeros(nn, dtype=bool)
        node_mark[cell_nodes[in_set].ravel()] = True
        grown = node_mark[cell_nodes].any(axis=1)
        if grown.sum() == in_set.sum():  # disconnected mesh: restart from the first cell outside
            grown[np.nonzero(~grown)[0][0]] = True
        in_set = grown
    return np.nonzero(in_set)[0]


def lib_of_fes_map(fes, nowned: dict | None = None) -> np.ndarray:
    """library vector index of every index of the FE-space numbering (see module docstring)."""
    n = fes.ndofs
    if nowned is None:
        return np.arange(n, dtype=np.int64)
    out = np.empty(n, dtype=np.int64)
    off = fes.offsets
    own, gh, o = {}, {}, 0
    for f in fes.field_order:
        own[f] = o
        o += nowned[f]
    for f in fes.field_order:
        gh[f] = o
        o += fes.nfree[f] - nowned[f]
    for f in fes.field_order:
        no, nf = nowned[f], fes.nfree[f]
        out[off[f] : off[f] + no] = own[f] + np.arange(no)
        out[off[f] + no : off[f] + nf] = gh[f] + np.arange(nf - no)
    return out


def assembly_parity(fes, prm, x_lib, rowptr, colval, nzval, r_lib, nrows, nowned=None, ncells=2048, seed_cell=0, gather=None, v_lib=None):
    """Device CSR (library numbering, owned rows) and residual against the C oracle on a block of >= ncells cells.
    Returns {"jac_rel", "res_rel", "csr_bitexact", "rows_checked", "entries_checked", "cells"}.
    `gather(idx) -> (colval[idx], nzval[idx])`: for matrices too large to copy to the host the caller keeps colval / nzval on
    the device and hands over only the entries of the sampled rows (colval and nzval arguments are then ignored).  With `v_lib`
    the result also carries "rows_lib" and "y_rows" = (A v) on the sampled rows from those entries (for the SpMV check)."""
    cells = sample_block(fes.mesh.cell_nodes, ncells, seed_cell)
    lof = lib_of_fes_map(fes, nowned)
    x_fes = np.ascontiguousarray(np.asarray(x_lib)[lof])
    co = COracle(fes, prm, cells=cells)
    gall = fes.cell_global_ids()
    gs = gall[cells]
    rp, cv = O.symbolic_csr(gs, fes.ndofs)
    nz = co.jacobian_values(x_fes, rp, cv)
    r = co.residual(x_fes)
    full = np.bincount(gall[gall >= 0], minlength=fes.ndofs)
    part = np.bincount(gs[gs >= 0], minlength=fes.ndofs)
    rows_fes = np.nonzero((full > 0) & (full == part) & (lof < nrows))[0]
    rows_lib = lof[rows_fes]
    # oracle entries of the complete rows, in library numbering, sorted by (row, col)
    cnt_o = (rp[rows_fes + 1] - rp[rows_fes]).astype(np.int64)
    idx_o = np.repeat(rp[rows_fes] - np.concatenate([[0], np.cumsum(cnt_o)[:-1]]), cnt_o) + np.arange(cnt_o.sum())
    ro = np.repeat(rows_lib, cnt_o)
    co_ = lof[cv[idx_o]]
    vo = nz[idx_o]
    order = np.lexsort((co_, ro))
    ro, co_, vo = ro[order], co_[order], vo[order]
    # device entries of the same rows
    rowptr = np.asarray(rowptr, dtype=np.int64)
    rl = np.sort(rows_lib)
    cnt_d = rowptr[rl + 1] - rowptr[rl]
    idx_d = np.repeat(rowptr[rl] - np.concatenate([[0], np.cumsum(cnt_d)[:-1]]), cnt_d) + np.arange(cnt_d.sum())
    rd = np.repeat(rl, cnt_d)
    if gather is not None:
        cd, vd = gather(idx_d)
        cd, vd = np.asarray(cd).astype(np.int64), np.asarray(vd)
    else:
        cd = np.asarray(colval)[idx_d].astype(np.int64)
        vd = np.asarray(nzval)[idx_d]
    bitexact = bool(len(rd) == len(ro) and np.array_equal(rd, ro) and np.array_equal(cd, co_))
    scale = float(np.abs(vo).max()) if len(vo) else 1.0
    jac_rel = float(np.abs(vd - vo).max() / scale) if bitexact and len(vo) else float("inf")
    worst = None
    if bitexact and len(vo):  # where the largest deviation sits (field block of the entry, both values): diagnostics only
        k = int(np.abs(vd - vo).argmax())
        inv = np.empty(len(lof), dtype=np.int64)
        inv[lof] = np.arange(len(lof))
        names = list(fes.offsets.keys())
        offs = np.array([fes.offsets[f] for f in names])
        fld = lambda g: names[int(np.searchsorted(offs, g, side="right") - 1)]
        worst = {"row_field": fld(inv[ro[k]]), "col_field": fld(inv[co_[k]]), "device": float(vd[k]), "oracle": float(vo[k]), "scale": scale}
    rr = np.asarray(r_lib)[rows_lib]
    res_rel = float(np.abs(rr - r[rows_fes]).max() / np.abs(r[rows_fes]).max()) if len(rows_fes) else float("inf")
    out = {"jac_rel": jac_rel, "res_rel": res_rel, "csr_bitexact": bitexact, "rows_checked": int(len(rows_fes)),
           "entries_checked": int(len(vo)), "cells": int(len(cells)), "worst_entry": worst}
    if v_lib is not None:  # host product of the sampled rows with the device's own entries
        seg = np.concatenate([[0], np.cumsum(cnt_d)])
        out["rows_lib"] = rl
        out["y_rows"] = np.add.reduceat(vd * np.asarray(v_lib)[cd], seg[:-1]) if len(vd) else np.zeros(0)
    return out


def spmv_parity(rowptr, colval, nzval, v_lib, y_dev):
    """max |y_dev - A v| / max |A v| with A = the device's own CSR multiplied on the host (all owned rows)."""
    import scipy.sparse as sp

    n = len(rowptr) - 1
    A = sp.csr_matrix((np.asarray(nzval), np.asarray(colval), np.asarray(rowptr)), shape=(n, len(v_lib)))
    y = A @ np.asarray(v_lib)
    return float(np.abs(np.asarray(y_dev) - y).max() / np.abs(y).max())
