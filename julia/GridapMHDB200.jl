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
using LinearAlgebra

const libmhd = get(ENV, "MHDB200_LIBRARY", "libmhdb200.so")   # like JULIA_PETSC_LIBRARY (ci_mpi.yml:9)

# ---- error convention: every call returns Cint, message via mhd_last_error_string (cf. @check_error_code,
#      src/Solvers/petsc.jl:16-27)
macro check(ex)
  quote
    rc = $(esc(ex))
    rc == 0 || error("libmhdb200: ", unsafe_string(ccall((:mhd_last_error_string, libmhd), Cstring, ())))
  end
end
const MHD_E_NOTCONV = Cint(-6)   # mhd_solve reached maxiter: a RESULT, not a failure (see solve! below)

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
  device_resident::Bool                            # allocate_jacobian returns a B200DeviceMatrix (solver :b200) or a host CSR
  keep::Any                                        # the host arrays the create call borrowed (kept only for debugging)
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
function B200FEOperator(U, V, params; tables, mesh, layout, fluid, device_resident=(params[:solver][:solver] == :b200), keep=nothing)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  @check ccall((:mhd_operator_create, libmhd), Cint,
               (Ref{MhdMesh}, Ref{MhdTables}, Ref{MhdLayout}, Ref{MhdParams}, Ref{Ptr{Cvoid}}), mesh, tables, layout, fluid, h)
  nr = Ref{Int64}(0); nc = Ref{Int64}(0); nnz = Ref{Int64}(0)
  @check ccall((:mhd_operator_symbolic, libmhd), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), h[], nr, nc, nnz)
  rowptr = Vector{Int64}(undef, nr[] + 1); colval = Vector{Int64}(undef, nnz[])
  @check ccall((:mhd_operator_get_csr, libmhd), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Cint, Cint), h[], rowptr, colval, 8, 0)
  op = B200FEOperator(U, V, h[], nr[], nnz[], rowptr, colval, device_resident, keep)
  finalizer(o -> nothing, op)   # explicit destroy only (collective objects must not be freed from GC; hunt.jl:204)
  return op
end
destroy!(op::B200FEOperator) = (ccall((:mhd_operator_destroy, libmhd), Cint, (Ptr{Cvoid},), op.handle); op.handle = C_NULL)

# ---- Gridap -> C structs: the part of the seam that decides whether Gridap's own entries come out (src/main.jl:207-233,
#      src/fespaces.jl:13-46).  Everything the kernels integrate with is READ OFF Gridap's objects -- reference tables from the
#      reffes at the quadrature points, cell dof ids and RT sign flips from the FE spaces -- never re-derived.
"""
    b200_operator(U, V, params) -> B200FEOperator

Drop-in for `_fe_operator(mfs,U,V,params)` of the H1-HDiv formulation on a sequential (or one rank's local) model.
"""
function b200_operator(U, V, params)
  model = params[:model]
  grid  = get_grid(model)
  # -- mesh: HEX8 cells in Gridap's lexicographic vertex order (what the geometry tables below assume)
  xs    = get_node_coordinates(grid)
  coords = collect(Float64, reinterpret(reshape, Float64, collect(xs)))            # 3 x nnodes, column major = [node][3]
  cn    = get_cell_node_ids(grid)
  ncells = length(cn)
  cell_nodes = Matrix{Int32}(undef, 8, ncells)
  for (c, ids) in enumerate(cn); cell_nodes[:, c] .= ids; end
  # -- solid cells (params[:solid], src/Applications/hunt.jl:168-176): u, p live on Ωf only => dof id 0 on solid cells
  Ωf = params[:Ωf]
  fluid_cells = Ωf === nothing ? collect(1:ncells) : get_cell_to_bgcell(Ωf) # Triangulation -> background cell ids (Gridap.Geometry; older: get_glue(Ωf,Val(3)).tface_to_mface)
  is_solid = ones(UInt8, ncells); is_solid[fluid_cells] .= 0
  has_solid = any(!iszero, is_solid)
  cell_sigma = has_solid ? _cellwise_sigma(params, ncells, is_solid) : Float64[]
  # -- reference tables at the points of Quadrature(HEX, q) (src/parameters.jl:381-389: q = 5 -> 27 Gauss points)
  q     = params[:fespaces][:q]
  quad  = Quadrature(HEX, q)
  xq, wq = get_coordinates(quad), get_weights(quad)
  @assert length(wq) == 27 "libmhdb200 integrates with the 27-point rule (order_u = order_j = 2)"
  reffe_u, reffe_p = _reffe(params, :reffe_u), _reffe(params, :reffe_p)
  reffe_j, reffe_φ = _reffe(params, :reffe_j), _reffe(params, :reffe_φ)
  geo   = LagrangianRefFE(Float64, HEX, 1)                                             # trilinear geometry map
  dgeo  = evaluate(Broadcasting(∇)(get_shapefuns(geo)), xq)                            # [q, v] VectorValue{3}
  geo_grad = Float64[dgeo[iq, v][k] for k in 1:3, v in 1:8, iq in 1:27]                # C layout [q][v][k]
  # velocity: the SCALAR Q2 basis (the library expands N_a e_c itself; Gridap's vector Lagrangian dofs are component major,
  # dof = node + nnodes*(comp-1): checked below on the cell dof ids)
  su    = get_shapefuns(LagrangianRefFE(Float64, HEX, get_order(reffe_u)))
  uval  = evaluate(su, xq); ugrad = evaluate(Broadcasting(∇)(su), xq)
  u_val = Float64[uval[iq, a] for a in 1:27, iq in 1:27]                               # [q][a]
  u_grad = Float64[ugrad[iq, a][k] for k in 1:3, a in 1:27, iq in 1:27]                # [q][a][k]
  pval  = evaluate(get_shapefuns(reffe_p), xq)
  p_val = Float64[pval[iq, k] for k in 1:4, iq in 1:27]
  sj    = get_shapefuns(reffe_j)
  jval  = evaluate(sj, xq); jdiv = evaluate(Broadcasting(divergence)(sj), xq)
  j_val = Float64[jval[iq, m][k] for k in 1:3, m in 1:36, iq in 1:27]                  # reference (un-mapped) RT basis
  j_div = Float64[jdiv[iq, m] for m in 1:36, iq in 1:27]
  fval  = evaluate(get_shapefuns(reffe_φ), xq)
  phi_val = Float64[fval[iq, l] for l in 1:8, iq in 1:27]
  w     = collect(Float64, wq)
  # -- dof ids: signed, 1-based, negative = Dirichlet (Gridap's own convention); 0 = absent on a solid cell
  V_u, V_p, V_j, V_φ = V[1], V[2], V[3], V[4]
  U_u, U_j = U[1], U[3]
  cd_u = _cell_dofs(V_u, 81, ncells, fluid_cells); cd_p = _cell_dofs(V_p, 4, ncells, fluid_cells)
  cd_j = _cell_dofs(V_j, 36, ncells, 1:ncells);    cd_φ = _cell_dofs(V_φ, 8, ncells, 1:ncells)
  # RT sign flips: Gridap keeps them with the cell map of the dof basis ("a cell flips the dofs of a facet iff it is the
  # second cell around it"); get_sign_flip(model, cell_reffes) returns, per cell, a Bool per local dof
  flips = ReferenceFEs.get_sign_flip(model, Fill(reffe_j, ncells))                    # Gridap.FESpaces (DivConformingFESpaces.jl)
  j_sign = Matrix{Int8}(undef, 36, ncells)
  for (c, f) in enumerate(flips); j_sign[:, c] .= ifelse.(f, Int8(-1), Int8(1)); end
  dir_u = collect(Float64, get_dirichlet_dof_values(U_u)); dir_j = collect(Float64, get_dirichlet_dof_values(U_j))
  nfree = (num_free_dofs(V_u), num_free_dofs(V_p), num_free_dofs(V_j), num_free_dofs(V_φ))
  ndir  = (num_dirichlet_dofs(V_u), 0, num_dirichlet_dofs(V_j), 0)
  order = _field_order(params)                                                       # (0,1,2,3) | (0,2,1,3) | (2,0,1,3): src/fespaces.jl:4-9
  fl    = params[:fluid]
  conv  = Dict(:none => 0, :picard => 1, :newton => 2)[get(fl, :convection, :newton)]
  c3(v) = (Float64(v[1]), Float64(v[2]), Float64(v[3]))
  fluid = MhdParams(fl[:α], fl[:β], fl[:γ], fl[:σ], fl[:ζᵤ], fl[:ζⱼ], c3(fl[:B]), c3(fl[:f]), c3(get(fl, :g, (0, 0, 0))), conv)
  keep  = (coords, cell_nodes, is_solid, cell_sigma, w, geo_grad, u_val, u_grad, p_val, j_val, j_div, phi_val, cd_u, cd_p, cd_j, cd_φ,
           j_sign, dir_u, dir_j)
  GC.@preserve keep begin
    mesh = MhdMesh(size(coords, 2), pointer(coords), ncells, pointer(cell_nodes), 1,
                   has_solid ? pointer(is_solid) : C_NULL, has_solid ? pointer(cell_sigma) : C_NULL)
    tables = MhdTables(27, pointer(w), pointer(geo_grad), pointer(u_val), pointer(u_grad), pointer(p_val), pointer(j_val),
                       pointer(j_div), pointer(phi_val))
    layout = MhdLayout((pointer(cd_u), pointer(cd_p), pointer(cd_j), pointer(cd_φ)), pointer(j_sign), nfree, nfree, ndir,
                       (isempty(dir_u) ? C_NULL : pointer(dir_u), C_NULL, isempty(dir_j) ? C_NULL : pointer(dir_j), C_NULL), order)
    return B200FEOperator(U, V, params; tables, mesh, layout, fluid)   # the library copies everything during the call
  end
end
_reffe(params, key) = (r = params[:fespaces][key]; r isa Tuple ? ReferenceFE(HEX, r[1], r[2]...; r[3]...) : r)
function _cell_dofs(V, n, ncells, cells)
  ids = get_cell_dof_ids(V)                          # on the space's own triangulation: one entry per cell of `cells`
  out = zeros(Int32, n, ncells)                      # 0 = the dof does not exist on this (solid) cell
  for (i, c) in enumerate(cells); out[:, c] .= ids[i]; end
  out
end
function _cellwise_sigma(params, ncells, is_solid)
  σs = params[:solid] === nothing ? 0.0 : params[:solid][:σ]
  Float64[is_solid[c] != 0 ? σs : params[:fluid][:σ] for c in 1:ncells]
end
function _field_order(params)
  s = params[:solver][:solver]
  s === :li2019 ? (Int32(2), Int32(0), Int32(1), Int32(3)) :                          # (j,u,p,φ)
  (s === :badia2024 || s === :b200) ? (Int32(0), Int32(2), Int32(1), Int32(3)) :      # ([u,j],p,φ)
  (Int32(0), Int32(1), Int32(2), Int32(3))                                           # Consecutive
end

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
  B200FEOperator(U, V, h[], nr[], nnz[], rowptr, colval, params[:solver][:solver] == :b200, nothing)
end

# Gridap NonlinearOperator API used by solve!(xh,solver,op) (src/main.jl:275) and by main.jl:158,163
function Algebra.allocate_residual(op::B200FEOperator, x::AbstractVector)
  zeros(Float64, op.nrows)
end
function Algebra.allocate_jacobian(op::B200FEOperator, x::AbstractVector)
  op.device_resident ? B200DeviceMatrix(op) : SparseMatrixCSR{0}(op.nrows, op.nrows, op.rowptr, op.colval, zeros(Float64, op.nnz))
end
function Algebra.residual!(b::AbstractVector, op::B200FEOperator, x::AbstractVector)
  @check ccall((:mhd_residual, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), op.handle, x, b)
  b
end
# Two matrix types (chosen by params[:solver][:matrix_type], src/main.jl:220):
#  * B200DeviceMatrix: the values never leave the GPU -- the matrix IS the operator handle (1.2 GB per assembly stay where the
#    device-resident Krylov solver reads them).  This is what :b200 solvers use; `nzval(A)` fetches a host copy on demand.
#  * SparseMatrixCSR{0,Float64,Int64}: for host-side solvers (:julia LU, PETSc): the values are copied out after every assembly
#    (bench.py reports both modes: `e2e` and `e2e_with_matrix_d2h`).
struct B200DeviceMatrix <: AbstractMatrix{Float64}
  op::B200FEOperator
end
Base.size(A::B200DeviceMatrix) = (A.op.nrows, A.op.nrows)
Base.getindex(A::B200DeviceMatrix, i::Int, j::Int) = error("B200DeviceMatrix is device resident: use host_copy(A)")
function host_copy(A::B200DeviceMatrix)
  v = Vector{Float64}(undef, A.op.nnz)
  @check ccall((:mhd_get_nzval, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}), A.op.handle, v)
  SparseMatrixCSR{0}(A.op.nrows, A.op.nrows, A.op.rowptr, A.op.colval, v)
end
function LinearAlgebra.mul!(y::AbstractVector, A::B200DeviceMatrix, x::AbstractVector)
  @check ccall((:mhd_spmv, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), A.op.handle, x, y)
  y
end
function Algebra.jacobian!(A::B200DeviceMatrix, op::B200FEOperator, x::AbstractVector)
  @check ccall((:mhd_jacobian, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), op.handle, x, C_NULL)   # no D2H of nzval
  A
end
function Algebra.residual_and_jacobian!(b, A::B200DeviceMatrix, op::B200FEOperator, x)
  @check ccall((:mhd_residual_and_jacobian, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), op.handle, x, b)
  b, A
end
function Algebra.jacobian!(A::SparseMatrixCSR, op::B200FEOperator, x::AbstractVector)
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
  verbose::Bool
end
B200LinearSolver(op, opts) = B200LinearSolver(op, opts, false)
struct B200SymbolicSetup <: Algebra.SymbolicSetup
  solver::B200LinearSolver
end
mutable struct B200NumericalSetup <: Algebra.NumericalSetup
  solver::B200LinearSolver
  handle::Ptr{Cvoid}
  last_iters::Int32
  last_residual::Float64
end
Algebra.symbolic_setup(s::B200LinearSolver, A::AbstractMatrix) = B200SymbolicSetup(s)
function Algebra.numerical_setup(ss::B200SymbolicSetup, A::AbstractMatrix)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  @check ccall((:mhd_solver_create, libmhd), Cint, (Ptr{Cvoid}, Ref{MhdSolverOpts}, Ref{Ptr{Cvoid}}), ss.solver.op.handle, ss.solver.opts, h)
  ns = B200NumericalSetup(ss.solver, h[], 0, 0.0)
  Algebra.numerical_setup!(ns, A)
end
function Algebra.numerical_setup!(ns::B200NumericalSetup, A::AbstractMatrix)
  @check ccall((:mhd_solver_setup, libmhd), Cint, (Ptr{Cvoid},), ns.handle)   # the matrix is already on the device
  ns
end
# GridapSolvers' FGMRESSolver(m,P;maxiter=m,...) (src/Solvers/badia2024.jl:40) RETURNS when it hits maxiter and the NewtonSolver
# around it carries on as an inexact Newton method; with m = maxiter = 15 and rtol = nl_rtol/10 that is the common case.
# So MHD_E_NOTCONV must not raise: it is reported through the (optional) log and the iteration count.
function Algebra.solve!(x::AbstractVector, ns::B200NumericalSetup, b::AbstractVector)
  iters = Ref{Int32}(0); res = Ref{Float64}(0.0)
  rc = ccall((:mhd_solve, libmhd), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}, Ref{Float64}, Ptr{Float64}),
             ns.handle, b, x, iters, res, C_NULL)
  if rc == MHD_E_NOTCONV
    ns.solver.verbose && @info "B200LinearSolver: maxiter reached" iterations=iters[] residual=res[]
  elseif rc != 0
    error("libmhdb200: ", unsafe_string(ccall((:mhd_last_error_string, libmhd), Cstring, ())))
  end
  ns.last_iters = iters[]; ns.last_residual = res[]
  x
end

# ---------------------------------------------------------------------------------------------------------------
# distributed runs: one MPI rank per GPU, as the reference runs under PartitionedArrays (src/Applications/hunt.jl:18-25,
# with_mpi).  The operator of a rank is built by b200_operator on its LOCAL spaces (own + ghost cells, local dof numbering with the
# owned dofs first -- `layout.nowned` says how many per field); what the library needs on top is (1) a communicator, (2) the
# ghost-exchange plan of the dof partition and (3), on one node, the peer-memory connection of the inboxes.  The plan is what
# PartitionedArrays keeps in the cache of the column PRange of the assembled matrix (`assembly_neighbors`, `assembly_local_indices`):
# for every neighbouring rank the local ids this rank sends (owned here, ghost there) and receives (ghost here), the two lists of a
# pair of ranks in the same order.  Python counterpart, executed in this repository on 2-8 GPUs: host/partition.py:distribute_operator.
import MPI

function comm_init(comm::MPI.Comm)
  rank, nranks = MPI.Comm_rank(comm), MPI.Comm_size(comm)
  id = zeros(UInt8, 128)                                       # ncclUniqueId travels through the host
  rank == 0 && @check ccall((:mhd_comm_get_unique_id, libmhd), Cint, (Ptr{UInt8},), id)
  MPI.Bcast!(id, 0, comm)
  @check ccall((:mhd_comm_init, libmhd), Cint, (Cint, Cint, Ptr{UInt8}), rank, nranks, id)
end
comm_finalize() = ccall((:mhd_comm_finalize, libmhd), Cint, ())

# neigh: ranks of the neighbours; send_ptr/recv_ptr: [nneigh+1] offsets into send_idx/recv_idx (0-based local dof ids)
function set_halo!(op::B200FEOperator, neigh::Vector{Int32}, send_ptr::Vector{Int64}, send_idx::Vector{Int32},
                   recv_ptr::Vector{Int64}, recv_idx::Vector{Int32})
  @check ccall((:mhd_operator_set_halo, libmhd), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Int32}, Ptr{Int64}, Ptr{Int32}),
               op.handle, length(neigh), neigh, send_ptr, send_idx, recv_ptr, recv_idx)
end

# SpMV + ghost exchange over NVLink peer memory (ranks of ONE node; without this call the exchange runs through NCCL send/recv):
# every rank exports its inbox (CUDA IPC handle, 64 bytes), the handles and the receive lists are gathered through MPI, and each
# rank tells the library, for every value it sends, the ghost slot that value fills on its neighbour.
# every rank's vector, as a vector of vectors (counts first, then MPI.Allgatherv!)
function _allgatherv(v::Vector{T}, comm::MPI.Comm) where T
  counts = MPI.Allgather(Int32[length(v)], comm)
  out = Vector{T}(undef, sum(counts))
  MPI.Allgatherv!(v, MPI.VBuffer(out, counts), comm)
  ptr = cumsum([0; Int.(counts)])
  [out[ptr[r]+1:ptr[r+1]] for r in 1:length(counts)]
end

function connect_peer_memory!(op::B200FEOperator, comm::MPI.Comm, neigh::Vector{Int32}, send_ptr::Vector{Int64},
                              recv_ptr::Vector{Int64}, recv_idx::Vector{Int32}, nrows::Int, ncols::Int)
  rank = MPI.Comm_rank(comm)
  handle = zeros(UInt8, 64)
  isempty(neigh) || @check ccall((:mhd_operator_halo_ipc_export, libmhd), Cint, (Ptr{Cvoid}, Ptr{UInt8}), op.handle, handle)
  # small host-side metadata, gathered once per mesh (collective: ranks without neighbours take part too)
  handles_all = reshape(MPI.Allgather(handle, comm), 64, :)
  sizes_all   = reshape(MPI.Allgather(Int64[nrows, ncols], comm), 2, :)
  neigh_all, rptr_all, ridx_all = _allgatherv(neigh, comm), _allgatherv(recv_ptr, comm), _allgatherv(recv_idx, comm)
  isempty(neigh) && return
  handles = reduce(vcat, (handles_all[:, s+1] for s in neigh))
  slot = Int32[findfirst(==(Int32(rank)), neigh_all[s+1]) - 1 for s in neigh]          # my position in the neighbour's list
  nghost = Int64[sizes_all[2, s+1] - sizes_all[1, s+1] for s in neigh]
  send_dst = Int32[]
  for (k, s) in enumerate(neigh)
    kk = slot[k] + 1
    seg = ridx_all[s+1][rptr_all[s+1][kk]+1:rptr_all[s+1][kk+1]] .- Int32(sizes_all[1, s+1])  # ghost slot = ghost id - row count, over there
    @assert length(seg) == send_ptr[k+1] - send_ptr[k]
    append!(send_dst, seg)
  end
  @check ccall((:mhd_operator_halo_ipc_connect, libmhd), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}),
               op.handle, handles, send_dst, slot, nghost)
  MPI.Barrier(comm)
end

# bit-reproducible assembly (coloured launches; SURVEY 5.2) and which assembly kernel the tables selected (7 = sum-factorised)
function set_deterministic!(op::B200FEOperator, on::Bool=true)
  ncolors = Ref{Int32}(0)
  @check ccall((:mhd_operator_set_deterministic, libmhd), Cint, (Ptr{Cvoid}, Int32, Ref{Int32}), op.handle, on ? 1 : 0, ncolors)
  Int(ncolors[])
end
function kernel_version(op::B200FEOperator)
  v = Ref{Int32}(0)
  @check ccall((:mhd_operator_get_kernel_version, libmhd), Cint, (Ptr{Cvoid}, Ref{Int32}), op.handle, v)
  Int(v[])
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
