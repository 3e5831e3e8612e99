"""The parity sampler of bench.py / the GPU tests (oracle/parity.py) checked on the CPU: the "device" matrix is the NumPy
oracle's global assembly, the sampler re-derives complete rows from a block of cells with the C oracle."""
import numpy as np

from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
from gridapmhd_jl_b200.host.partition import hunt_cell_partition, partition_fespaces
from oracle import mhd_oracle as O
from oracle import parity as P


def _case(nc=(6, 5)):
    p = hunt_params(nc=nc, B=(0.0, 20.0, 0.0), solver="badia2024", zeta_u=1.0, zeta_j=2.0)
    fes = setup_spaces(p)
    fl = p["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    x = np.random.default_rng(0).random(fes.ndofs)
    return fes, prm, x


def test_sampler_accepts_the_oracle_and_detects_a_wrong_entry():
    fes, prm, x = _case()
    A = O.jacobian(fes, x, prm)
    r = O.residual(fes, x, prm)
    out = P.assembly_parity(fes, prm, x, A.indptr, A.indices, A.data, r, fes.ndofs, ncells=30)
    assert out["csr_bitexact"] and out["jac_rel"] < 1e-13 and out["res_rel"] < 1e-13
    assert out["rows_checked"] > 500 and out["cells"] >= 30
    bad = A.data.copy()
    rows = np.repeat(np.arange(fes.ndofs), np.diff(A.indptr))
    # first entry of a checked row (every dof of the seed cell is complete once one ring around it is in the block)
    g0 = fes.cell_global_ids()[0]
    k = np.nonzero(rows == g0[g0 >= 0][0])[0][0]
    bad[k] += 1e-6 * np.abs(A.data).max()
    out = P.assembly_parity(fes, prm, x, A.indptr, A.indices, bad, r, fes.ndofs, ncells=30)
    assert out["jac_rel"] > 1e-7
    cols = A.indices.copy()
    cols[k] += 1 if k + 1 < len(cols) and rows[k + 1] == rows[k] and cols[k + 1] > cols[k] + 1 else -1
    assert not P.assembly_parity(fes, prm, x, A.indptr, cols, A.data, r, fes.ndofs, ncells=30)["csr_bitexact"]


def test_sampler_on_a_partition_uses_the_library_numbering():
    fes, prm, x = _case((6, 4))
    Ag = O.jacobian(fes, x, prm)
    rg = O.residual(fes, x, prm)
    part = hunt_cell_partition(fes.mesh, (2, 1))
    for rank in (0, 1):
        ps = partition_fespaces(fes, part, rank)
        gl = ps.local_vector_ids()
        Aloc = Ag[gl[: ps.nrows]][:, gl].tocsr()
        Aloc.sort_indices()
        out = P.assembly_parity(ps.fes, prm, x[gl], Aloc.indptr, Aloc.indices, Aloc.data, rg[gl[: ps.nrows]], ps.nrows,
                                nowned=ps.nowned, ncells=20)
        assert out["csr_bitexact"] and out["jac_rel"] < 1e-13 and out["res_rel"] < 1e-13 and out["rows_checked"] > 200
        y = Aloc @ x[gl]
        assert P.spmv_parity(Aloc.indptr, Aloc.indices, Aloc.data, x[gl], y) < 1e-15


def test_sampler_with_a_gather_callback_and_the_row_product():
    """Large matrices stay on the device: the sampler then receives only the entries of its rows through `gather` and returns
    the product of those rows (bench.py --nc-global)."""
    fes, prm, x = _case()
    A = O.jacobian(fes, x, prm)
    r = O.residual(fes, x, prm)
    v = np.cos(0.37 * np.arange(fes.ndofs))
    calls = []

    def gather(idx):
        calls.append(len(idx))
        return A.indices[idx], A.data[idx]

    out = P.assembly_parity(fes, prm, x, A.indptr, None, None, r, fes.ndofs, ncells=30, gather=gather, v_lib=v)
    ref = P.assembly_parity(fes, prm, x, A.indptr, A.indices, A.data, r, fes.ndofs, ncells=30)
    assert calls == [out["entries_checked"]] and out["csr_bitexact"]
    assert out["jac_rel"] == ref["jac_rel"] and out["res_rel"] == ref["res_rel"]
    y = A @ v
    assert np.abs(out["y_rows"] - y[out["rows_lib"]]).max() <= 1e-13 * np.abs(y).max()


def test_bench_parses_the_ncu_dram_csv_of_a_committed_capture():
    """bench.py measures `roofline.traffic` through an ncu child run; its CSV parser is checked here on a capture kept in profiles/"""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    out = bench.parse_ncu_dram_csv(os.path.join(root, "profiles", "r2_dram_bytes_bfs_order.csv"))
    assert out == {"jacobian": 952693760 + 1194133248}
