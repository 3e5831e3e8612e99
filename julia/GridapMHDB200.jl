# GridapMHDB200.jl -- Julia side of the drop-in boundary (NOT executed in the build container: no Julia there).
#
# This is the binding a GridapMHD maintainer adds next to src/Solvers/petsc.jl: it keeps GridapMHD.main(params),
# the params Dict, setup_fe_spaces and the solver entry points, and reaches CUDA only through `ccall` into
# libmhdb200.so (include/mhdb200.h).  It plugs into the two seams of the reference:
#   seam 1  _fe_operator(U,V,params)            src/main.jl:207-233   -> B200FEOperator <: FEOperator
#   seam 2  _solver(::Val{:b200},op,params)     src/main.jl:181-190   -> NewtonSolver(B200LinearSolver(...))
#           get_block_solver(::Val{:b200},...)  src/Solvers/gridap.jl:2-3
# and needs the matching methods of _multi_field_style / default_solver_params / uses_petsc / space_uses_multigrid
# (src/fespaces.jl:4-9, src/parameters.jl:221-303).
module GridapMHDB200

using Gridap, Gridap.FESpaces, Gridap.Algebra, Gridap.ReferenceFEs, Gridap.Geometry, Gridap.MultiField
using SparseMatricesCSR

const libmhd = get(ENV, "MHDB200_LIBRARY", "libmhdb200.so")   # like JULIA_PETSC_LIBRARY (ci_mpi.yml:9)

# ---- error convention: every call returns Cint, message via mhd_last_error_string (cf. @check_error_code,
#      src/Solvers/petsc.jl:16-27)
macro check(ex)
  quote
    rc = $(esc(ex))
    rc == 0 || error("libmhdb200: ", unsafe_string(ccall((:mhd_last_error_string, libmhd), Cstring, ())))
  end
end

struct MhdMesh
  nnodes::Int64; coords::Ptr{Float64}; ncells::Int64; cell_nodes::Ptr{Int32}; index_base::Int32
  cell_solid::Ptr{UInt8}; cell_sigma::Ptr{Float64}   # C_NULL without params[:solid]
end
struct MhdTables
  nq::Int32; w::Ptr{Float64}; geo_grad::Ptr{Float64}; u_val::Ptr{Float64}; u_grad::Ptr{Float64}
  p_val::Ptr{Float64}; j_val::Ptr{Float64}; j_div::Ptr{Float64}; phi_val::Ptr{Float64}
end
struct MhdLayout
  cell_dofs::NTuple{4,Ptr{Int32}}; j_sign::Ptr{Int8}
  nfree::NTuple{4,Int64}; nowned::NTuple{4,Int64}; ndir::NTuple{4,Int64}
  dir_values::NTuple{4,Ptr{Float64}}; field_order::NTuple{4,Int32}
end
struct MhdParams
  alpha::Float64; beta::Float64; gamma::Float64; sigma::Float64; zeta_u::Float64; zeta_j::Float64
  B::NTuple{3,Float64}; f::NTuple{3,Float64}; g::NTuple{3,Float64}; convection::Int32
end

# library scope, like GridapPETSc.with(args=...) do ... end (src/Applications/hunt.jl:202-206)
function with(f; device=0)
  @check ccall((:mhd_init, libmhd), Cint, (Cint,), device)
  try
    return f()
  finally
    ccall((:mhd_finalize, libmhd), Cint, ())
  end
end

# ---------------------------------------------------------------------------------------------------------------
# seam 1: the FE operator
mutable struct B200FEOperator <: FEOperator
  trial; test
  handle::Ptr{Cvoid}
  nrows::Int; nnz::Int
  rowptr::Vector{Int64}; colval::Vector{Int64}     # fetched once (0-based, SparseMatrixCSR{0})
end
FESpaces.get_trial(op::B200FEOperator) = op.trial
FESpaces.get_test(op::B200FEOperator) = op.test

"""
    B200FEOperator(U,V,params)

Built where `_fe_operator(mfs,U,V,params)` builds `FEOperator(res,jac,U,V,assem)` (src/main.jl:218-233).
Tables handed over (all borrowed for the call only):
  * node coordinates / cell node ids of `params[:model]` (get_node_coordinates, get_cell_node_ids; 1-based -> index_base=1)
  * reference tables at `Quadrature(HEX,params[:fespaces][:q])` points: evaluate(get_shapefuns(reffe),x), gradients, ...
    for reffe_u (scalar part), reffe_p, reffe_j (+ divergence), reffe_φ  (src/parameters.jl:436-441,521-525)
  * `get_cell_dof_ids(V_f)` per field (signed, 1-based), `get_dirichlet_dof_values(U_f)`, RT sign flips
    (`get_sign_flip`), free/Dirichlet counts, and the field order of `_multi_field_style(params)`
  * fluid parameters α β γ σ ζᵤ ζⱼ B f g convection from `params[:fluid]` (src/weakforms.jl:71-83)
"""
function B200FEOperator(U, V, params; tables, mesh, layout, fluid)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  @check ccall((:mhd_operator_create, libmhd), Cint,
               (Ref{MhdMesh}, Ref{MhdTables}, Ref{MhdLayout}, Ref{MhdParams}, Ref{Ptr{Cvoid}}), mesh, tables, layout, fluid, h)
  nr = Ref{Int64}(0); nc = Ref{Int64}(0); nnz = Ref{Int64}(0)
  @check ccall((:mhd_operator_symbolic, libmhd), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), h[], nr, nc, nnz)
  rowptr = Vector{Int64}(undef, nr[] + 1); colval = Vector{Int64}(undef, nnz[])
  @check ccall((:mhd_operator_get_csr, libmhd), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Cint, Cint), h[], rowptr, colval, 8, 0)
  op = B200FEOperator(U, V, h[], nr[], nnz[], rowptr, colval)
  finalizer(o -> nothing, op)   # explicit destroy only (collective objects must not be freed from GC; hunt.jl:204)
  return op
end
destroy!(op::B200FEOperator) = (ccall((:mhd_operator_destroy, libmhd), Cint, (Ptr{Cvoid},), op.handle); op.handle = C_NULL)

# ---- H1-H1 formulation (params[:fespaces][:current_disc] = :H1 => U = (U_u,U_p,U_φ), src/fespaces.jl:32-41;
#      weak_form_h1_h1, src/weakforms.jl:344-355).  Same operator type and methods; only the creation call differs.
struct MhdTablesH1H1
  nq::Int32; w::Ptr{Float64}; geo_grad::Ptr{Float64}; u_val::Ptr{Float64}; u_grad::Ptr{Float64}
  p_val::Ptr{Float64}; phi_grad::Ptr{Float64}     # phi_grad: ∇(shape functions of reffe_φ = LagrangianRefFE(Float64,HEX,3)) at the points
end
struct MhdLayoutH1H1
  cell_dofs::NTuple{3,Ptr{Int32}}                  # get_cell_dof_ids of V_u, V_p, V_φ
  nfree::NTuple{3,Int64}; nowned::NTuple{3,Int64}; ndir::NTuple{3,Int64}
  dir_values::NTuple{3,Ptr{Float64}}; field_order::NTuple{3,Int32}
end
function B200H1H1FEOperator(U, V, params; tables::MhdTablesH1H1, mesh::MhdMesh, layout::MhdLayoutH1H1, fluid::MhdParams)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  @check ccall((:mhd_h1h1_operator_create, libmhd), Cint,
               (Ref{MhdMesh}, Ref{MhdTablesH1H1}, Ref{MhdLayoutH1H1}, Ref{MhdParams}, Ref{Ptr{Cvoid}}), mesh, tables, layout, fluid, h)
  nr = Ref{Int64}(0); nc = Ref{Int64}(0); nnz = Ref{Int64}(0)
  @check ccall((:mhd_operator_symbolic, libmhd), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), h[], nr, nc, nnz)
  rowptr = Vector{Int64}(undef, nr[] + 1); colval = Vector{Int64}(undef, nnz[])
  @check ccall((:mhd_operator_get_csr, libmhd), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Cint, Cint), h[], rowptr, colval, 8, 0)
  B200FEOperator(U, V, h[], nr[], nnz[], rowptr, colval)
end

# Gridap NonlinearOperator API used by solve!(xh,solver,op) (src/main.jl:275) and by main.jl:158,163
function Algebra.allocate_residual(op::B200FEOperator, x::AbstractVector)
  zeros(Float64, op.nrows)
end
function Algebra.allocate_jacobian(op::B200FEOperator, x::AbstractVector)
  SparseMatrixCSR{0}(op.nrows, op.nrows, op.rowptr, op.colval, zeros(Float64, op.nnz))
end
function Algebra.residual!(b::AbstractVector, op::B200FEOperator, x::AbstractVector)
  @check ccall((:mhd_residual, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), op.handle, x, b)
  b
end
function Algebra.jacobian!(A::SparseMatrixCSR, op::B200FEOperator, x::AbstractVector)
  # values are copied out only when the caller wants them on the host (direct solvers); the B200 linear solver
  # below passes C_NULL and keeps the matrix on the device behind the handle
  @check ccall((:mhd_jacobian, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), op.handle, x, A.nzval)
  A
end
function Algebra.residual_and_jacobian!(b, A::SparseMatrixCSR, op::B200FEOperator, x)
  @check ccall((:mhd_residual_and_jacobian, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), op.handle, x, b)
  @check ccall((:mhd_get_nzval, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}), op.handle, A.nzval)
  b, A
end

# ---------------------------------------------------------------------------------------------------------------
# seam 2: the linear solver (FGMRES + block-triangular preconditioner, device resident)
# ---- post-processing of _hunt (src/Applications/hunt.jl:239-260): replaces the ∫(...)dΩ_phys block between
#      tic!(t) and toc!(t,"post_process"); `tables6` are the reference tables of Measure(Ω,2*(order+1))
struct MhdHuntPost
  a::Float64; mu::Float64; sigma::Float64; grad_pz::Float64; Ha::Float64
  nsums::Int32; reserved::Int32
  u0::Float64; jscale::Float64
end
function hunt_error_norms(op::B200FEOperator, x::AbstractVector, tables6::MhdTables; L, μ, σ, grad_pz, Ha, nsums, u0, jscale)
  out = zeros(6)
  prm = MhdHuntPost(L, μ, σ, grad_pz, Ha, nsums, 0, u0, jscale)
  @check ccall((:mhd_hunt_error_norms, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{MhdTables}, Ref{MhdHuntPost}, Ptr{Float64}),
               op.handle, x, tables6, prm, out)
  (; eu_l2=out[1], eu_h1=out[2], ej_l2=out[3], uh_l2=out[4], uh_h1=out[5], jh_l2=out[6])
end

struct MhdSolverOpts
  m::Int32; maxiter::Int32; rtol::Float64; atol::Float64; precond::Int32
  uj_inner_its::Int32; uj_inner_restart::Int32; alpha_p::Float64; alpha_phi::Float64; uj_solver::Int32; patch_its::Int32
  patch_omega::Float64
end

# uj_solver = 2 (MHD_UJ_GMRES_PATCH): hand the vertex-star dof lists of the (u,j) block to the solver before the first
# numerical_setup! -- what gmg_block_jacobi_smoothers (src/Solvers/gmg.jl:62-81) gives PatchBasedSmoothers.BlockJacobiSolver:
#   ptopo = Geometry.PatchTopology(ReferenceFE{0}, model); patch_ptr / patch_dofs = free (u,j) dof ids per patch, 0-based, sorted
function set_patches!(ns, patch_ptr::Vector{Int64}, patch_dofs::Vector{Int32})
  @check ccall((:mhd_solver_set_patches, libmhd), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int32}),
               ns.handle, length(patch_ptr) - 1, patch_ptr, patch_dofs)
  ns
end
# precond = 3 (MHD_PC_H1H1_BLOCKS, src/Solvers/h1h1blocks.jl): set_patches! carries the u patches, this one the phi patches
function set_phi_patches!(ns, patch_ptr::Vector{Int64}, patch_dofs::Vector{Int32})
  @check ccall((:mhd_solver_set_phi_patches, libmhd), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int32}),
               ns.handle, length(patch_ptr) - 1, patch_ptr, patch_dofs)
  ns
end

struct B200LinearSolver <: Algebra.LinearSolver
  op::B200FEOperator
  opts::MhdSolverOpts
end
struct B200SymbolicSetup <: Algebra.SymbolicSetup
  solver::B200LinearSolver
end
mutable struct B200NumericalSetup <: Algebra.NumericalSetup
  solver::B200LinearSolver
  handle::Ptr{Cvoid}
end
Algebra.symbolic_setup(s::B200LinearSolver, A::AbstractMatrix) = B200SymbolicSetup(s)
function Algebra.numerical_setup(ss::B200SymbolicSetup, A::AbstractMatrix)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  @check ccall((:mhd_solver_create, libmhd), Cint, (Ptr{Cvoid}, Ref{MhdSolverOpts}, Ref{Ptr{Cvoid}}), ss.solver.op.handle, ss.solver.opts, h)
  ns = B200NumericalSetup(ss.solver, h[])
  Algebra.numerical_setup!(ns, A)
end
function Algebra.numerical_setup!(ns::B200NumericalSetup, A::AbstractMatrix)
  @check ccall((:mhd_solver_setup, libmhd), Cint, (Ptr{Cvoid},), ns.handle)   # the matrix is already on the device
  ns
end
function Algebra.solve!(x::AbstractVector, ns::B200NumericalSetup, b::AbstractVector)
  iters = Ref{Int32}(0); res = Ref{Float64}(0.0)
  @check ccall((:mhd_solve, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}, Ref{Float64}, Ptr{Float64}),
               ns.handle, b, x, iters, res, C_NULL)
  x
end

# what GridapMHD itself needs (one more symbol, SURVEY.md 5.6):
#   _multi_field_style(::Val{:b200}) = BlockMultiFieldStyle(3,(2,1,1),(1,3,2,4))            # ([u,j],p,φ), fespaces.jl:8
#   uses_petsc(::Val{:b200}) = false ; space_uses_multigrid(::Val{:b200},solver) = fill(false,4)
#   default_solver_params(::Val{:b200}) = Dict(:solver=>:b200, :matrix_type=>SparseMatrixCSR{0,Float64,Int64},
#        :vector_type=>Vector{Float64}, :niter=>20, :niter_ls=>15, :rtol=>1e-6, :atol=>1e-8, ...)   # parameters.jl:259-271
#   _fe_operator(::BlockMultiFieldStyle,U,V,params)  -> B200FEOperator(U,V,params;...) when params[:solver][:solver]==:b200
#   _solver(::Val{:b200},op,params) = GridapSolvers.NewtonSolver(B200LinearSolver(op,opts);
#        maxiter=params[:solver][:niter], atol=params[:solver][:atol], rtol=params[:solver][:rtol])      # badia2024.jl:46

end # module
