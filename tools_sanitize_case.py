import numpy as np, sys
sys.path.insert(0,".")
import gridapmhd_jl_b200
from gridapmhd_jl_b200 import lib as L
from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
from gridapmhd_jl_b200.feoperator import B200FEOperator
L.init(0)
for kw in (dict(zeta_u=0.0, zeta_j=0.0, convection="newton"), dict(zeta_u=2.0, zeta_j=3.0, convection="picard")):
    params = hunt_params(nc=(3,2), B=(0.0,10.0,0.0), **kw); fes=setup_spaces(params); op=B200FEOperator(fes, params["fluid"])
    x = np.random.default_rng(1).random(fes.ndofs)
    A = op.allocate_jacobian(); b=np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, x); op.jacobian(x); r=op.residual(x); y=op.spmv(x)
    print("ok", kw, float(np.abs(r-b).max()))
    op.destroy()
L.finalize()
