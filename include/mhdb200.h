/* mhdb200.h -- C ABI of libmhdb200.so: B200 (sm_100a) assembly + Krylov kernels for the
 * inductionless-MHD H1-HDiv hot path of GridapMHD.jl.
 *
 * This is the drop-in boundary a Julia host reaches with `ccall` (see INTEGRATION.md and
 * julia/GridapMHDB200.jl).  Each entry point cites the reference interface it stands behind
 * (paths relative to the GridapMHD.jl checkout).
 *
 * Conventions
 *   - every function returns int: 0 = OK, <0 = error class (MHD_E_*); text via mhd_last_error_string().
 *     The library never aborts/exits (GridapPETSc precedent: src/Solvers/petsc.jl:16-27 @check_error_code).
 *   - handles are opaque, created and destroyed explicitly (no finalizer-driven frees; hunt.jl:204).
 *   - host arrays are BORROWED for the duration of the call only; the library copies what it keeps.
 *   - every `double*` / `const double*` vector argument of the compute calls may be a HOST pointer or a
 *     DEVICE pointer (detected with cudaPointerGetAttributes).  Host pointers: H2D copy, compute, D2H copy,
 *     stream synchronised on return.  Device pointers: work is enqueued on the library stream and the call
 *     returns without synchronising.
 *   - cell->dof ids are Gridap's: per field, 1-based, signed; id<0 is a Dirichlet dof and -id (1-based)
 *     indexes the field's Dirichlet-value array; id 0 marks a dof that does not exist on that cell (u, p on solid
 *     cells, whose spaces live on the fluid triangulation only).  Vertex ids in `cell_nodes` use `index_base`.
 *   - global vector layout: fields concatenated in `field_order` (src/fespaces.jl:4-9 _multi_field_style).
 *   - emitted CSR: sorted column indices, explicit zeros kept, Dirichlet rows/cols dropped
 *     (Gridap SparseMatrixAssembler semantics, src/main.jl:222-223).
 */
#ifndef MHDB200_H
#define MHDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MHD_OK 0
#define MHD_E_INVALID (-1)   /* bad argument / inconsistent tables */
#define MHD_E_CUDA (-2)      /* CUDA runtime error (sticky async errors surface at the next sync) */
#define MHD_E_STATE (-3)     /* call order violated (e.g. jacobian before symbolic) */
#define MHD_E_CAPACITY (-4)  /* a size limit of the implementation was exceeded */
#define MHD_E_COMM (-5)      /* NCCL / communicator error */
#define MHD_E_NOTCONV (-6)   /* iterative solve hit maxiter without meeting the tolerance */

enum { MHD_FIELD_U = 0, MHD_FIELD_P = 1, MHD_FIELD_J = 2, MHD_FIELD_PHI = 3 };
enum { MHD_CONV_NONE = 0, MHD_CONV_PICARD = 1, MHD_CONV_NEWTON = 2 }; /* src/parameters.jl:721 */

typedef struct mhd_operator mhd_operator_t;
typedef struct mhd_solver mhd_solver_t;

/* Mesh: what Gridap's UnstructuredGrid holds (node coordinates + cell node ids, HEX8, lexicographic local
 * vertex order x fastest).  Geometry is the trilinear map of the 8 vertices. */
typedef struct {
  int64_t nnodes;
  const double* coords;      /* [nnodes*3] */
  int64_t ncells;
  const int32_t* cell_nodes; /* [ncells*8] */
  int32_t index_base;        /* 0 or 1 */
  /* solid sub-domain (params[:solid], src/weakforms.jl:314-338; NULL when there is none): cells flagged 1 carry only
   * j and phi -- their u and p dof ids are 0 (= absent) -- and use cell_sigma[cell] as conductivity. */
  const uint8_t* cell_solid; /* [ncells] or NULL */
  const double* cell_sigma;  /* [ncells] or NULL */
} mhd_mesh_t;

/* Reference-element tables at the nq cell quadrature points (src/parameters.jl:436-441,521-525,617-639:
 * Q2 vector Lagrangian, P1disc, RT1, Q1disc, Quadrature(HEX,5) => nq = 27).  The kernels are basis-agnostic:
 * they only consume these numbers.  u is vector-valued with local dof = a + 27*c (component-major). */
typedef struct {
  int32_t nq;             /* must be 27 */
  const double* w;        /* [nq] reference weights */
  const double* geo_grad; /* [nq*8*3]  d(trilinear vertex function v)/d xi_k */
  const double* u_val;    /* [nq*27]   scalar Q2 basis */
  const double* u_grad;   /* [nq*27*3] reference gradients */
  const double* p_val;    /* [nq*4] */
  const double* j_val;    /* [nq*36*3] reference RT basis (contravariant Piola applied by the kernel) */
  const double* j_div;    /* [nq*36]   reference divergence */
  const double* phi_val;  /* [nq*8] */
} mhd_tables_t;

/* FE-space layout: what setup_fe_spaces (src/fespaces.jl:13-46) produces. */
typedef struct {
  const int32_t* cell_dofs[4]; /* per field [ncells*ndofs_f] signed 1-based; ndofs = 81,4,36,8 */
  const int8_t* j_sign;        /* [ncells*36] +1/-1 sign flips of the RT face dofs */
  int64_t nfree[4];            /* free dofs per field (local numbering: owned first, then ghosts) */
  int64_t nowned[4];           /* owned free dofs per field; == nfree on a single GPU */
  int64_t ndir[4];             /* Dirichlet dofs per field */
  const double* dir_values[4]; /* [ndir_f] Dirichlet values (may be NULL when ndir_f == 0) */
  int32_t field_order[4];      /* field ids in the order they appear in the global vector */
} mhd_layout_t;

/* retrieve_fluid_params (src/weakforms.jl:71-83), constant coefficients. */
typedef struct {
  double alpha, beta, gamma, sigma, zeta_u, zeta_j;
  double B[3], f[3], g[3];
  int32_t convection; /* MHD_CONV_* */
} mhd_params_t;

/* FGMRES + preconditioner options (src/Solvers/badia2024.jl:32-45, src/parameters.jl:259-271). */
/* MHD_PC_H1H1_BLOCKS: H1H1BlockSolver (src/Solvers/h1h1blocks.jl:2-43) for operators of mhd_h1h1_operator_create with layout
 * (u,p,phi): upper block-triangular, coefficients [1 1 1; 0 1 0; 0 0 1], p block = alpha_p x mass (alpha_p = -1/(beta+zeta_u)),
 * u and phi blocks = inner GMRES (uj_inner_its / uj_inner_restart) on the assembled diagonal blocks with Jacobi
 * (MHD_UJ_GMRES_JACOBI) or their vertex-patch solvers (MHD_UJ_GMRES_PATCH: mhd_solver_set_patches = u patches,
 * mhd_solver_set_phi_patches = phi patches; the reference puts GMG with the same patch smoothers there, gmg.jl:107-140). */
enum { MHD_PC_NONE = 0, MHD_PC_JACOBI = 1, MHD_PC_BLOCK_TRI = 2, MHD_PC_H1H1_BLOCKS = 3 };
/* (u,j)-block solver of the block-triangular preconditioner: inner Jacobi-GMRES (any size, any rank count) or an
 * exact dense LU on the device (cuSOLVER getrf/getrs; single GPU, n_uj <= 24576) -- the device stand-in for the
 * reference's direct block solver on small problems. */
/* MHD_UJ_GMRES_PATCH: inner GMRES preconditioned by the vertex-patch block-Jacobi smoother of
 * gmg_block_jacobi_smoothers (src/Solvers/gmg.jl:62-81): z = Richardson(patch_its sweeps, damping patch_omega) of the additive
 * Schwarz operator sum_p R_p' inv(A_p) R_p over the vertex stars; needs mhd_solver_set_patches.  Any size, one GPU. */
enum { MHD_UJ_GMRES_JACOBI = 0, MHD_UJ_DENSE_LU = 1, MHD_UJ_GMRES_PATCH = 2 };
typedef struct {
  int32_t m;            /* restart length (niter_ls, default 15) */
  int32_t maxiter;      /* total Krylov iterations (reference: == m) */
  double rtol, atol;    /* relative (to ||b - A x0||) and absolute tolerance */
  int32_t precond;      /* MHD_PC_* */
  int32_t uj_inner_its; /* MHD_PC_BLOCK_TRI: inner GMRES iterations on the (u,j) block */
  int32_t uj_inner_restart;
  double alpha_p, alpha_phi; /* scalings of the p / phi mass blocks (badia2024.jl:11-12) */
  int32_t uj_solver;    /* MHD_UJ_*: how block_solvers[1] (LU/MUMPS in the reference, badia2024.jl:22) is realised */
  int32_t patch_its;    /* MHD_UJ_GMRES_PATCH: Richardson sweeps per preconditioner application (reference: niter = 10; default 1) */
  double patch_omega;   /* damping of the sweeps (reference: w = 0.2, gmg.jl:62); irrelevant for patch_its = 1 under GMRES */
} mhd_solver_opts_t;

/* ---- library lifetime (GridapPETSc.with(args=...) do ... end; src/Applications/hunt.jl:202-206) ---- */
int mhd_init(int device_ordinal);
int mhd_finalize(void);
int mhd_set_stream(void* cuda_stream); /* stream all later calls enqueue on (default: legacy stream 0) */
int mhd_device_synchronize(void);
const char* mhd_last_error_string(void);

/* ---- multi-GPU: one process per GPU; rank 0 creates the id, the host broadcasts it (MPI / torch.distributed).
 * Replaces PartitionedArrays' MPI backend (with_mpi, src/Applications/hunt.jl:22-24). ---- */
int mhd_comm_get_unique_id(void* id128 /* 128 bytes out */);
int mhd_comm_init(int rank, int nranks, const void* id128);
int mhd_comm_finalize(void);

/* ---- operator: stands behind _fe_operator / FEOperator(res,jac,U,V,assem) (src/main.jl:207-233) ---- */
int mhd_operator_create(const mhd_mesh_t*, const mhd_tables_t*, const mhd_layout_t*, const mhd_params_t*,
                        mhd_operator_t** out);
int mhd_operator_destroy(mhd_operator_t*);
int mhd_operator_set_params(mhd_operator_t*, const mhd_params_t*); /* continuation: src/main.jl:243-260 */

/* Which Jacobian kernel the operator runs: 7 = fully sum-factorised kernel (csrc/hdiv_v7.cu; chosen automatically at
 * mhd_operator_create when the tables of mhd_tables_t turn out to be tensor products of 1-D factors on the tensor Gauss rule --
 * true for the reference's HEX elements, src/parameters.jl:436-441,521-525 -- unless MHD_JAC_V7=0), 5 = generic tensor-core
 * kernel (csrc/assembly.cu, any tables).  H1-H1 operators report 0. */
int mhd_operator_get_kernel_version(mhd_operator_t*, int32_t* version);
/* Deterministic assembly ("offer a deterministic (coloured) mode"): cells are greedily coloured so that no two
 * cells of a colour share a dof and the Jacobian / residual kernels run one launch per colour, which fixes the order in which
 * contributions to a shared nnz / row are summed => bit-identical results from run to run (the reference's sequential
 * assembly loop is deterministic by construction).  Needs kernel version 7.  on = 0 switches back to one launch. */
int mhd_operator_set_deterministic(mhd_operator_t*, int32_t on, int32_t* ncolors /* nullable out */);

/* Ghost exchange plan of the operator's vectors (PartitionedArrays consistent!/PVector):
 * for neighbour k: send x[send_idx[send_ptr[k]..send_ptr[k+1])] (owned local ids, 0-based) and receive into
 * local ids recv_idx[recv_ptr[k]..) (ghost ids). */
int mhd_operator_set_halo(mhd_operator_t*, int32_t nneigh, const int32_t* neigh_ranks, const int64_t* send_ptr,
                          const int32_t* send_idx, const int64_t* recv_ptr, const int32_t* recv_idx);

/* Fused SpMV + ghost exchange over NVLink peer memory (CUDA IPC; same-node ranks): every rank exports the handle of its
 * inbox, the host exchanges the 64-byte handles (MPI / torch.distributed) and hands each rank its neighbours' handles
 * plus: for every entry of the send list the ghost slot it fills on its neighbour (the neighbour's recv_idx entry minus
 * the neighbour's row count), and per neighbour k this rank's position in k's neighbour list and k's ghost count.  Afterwards mhd_spmv / mhd_solve run every
 * product as two launches: a product kernel whose first CTAs push the interface values into the neighbours' inboxes while the
 * others multiply the local columns, and a tail kernel over the interface rows that waits for the arrivals and adds the ghost part.
 * Without it the exchange runs as pack -> ncclSend/ncclRecv -> unpack. */
int mhd_operator_halo_ipc_export(mhd_operator_t*, void* handle64 /* 64 bytes out */);
int mhd_operator_halo_ipc_connect(mhd_operator_t*, const void* handles /* nneigh x 64 B */, const int32_t* send_dst /* [nsend] */,
                                  const int32_t* peer_slot, const int64_t* peer_nghost);
int mhd_operator_halo_status(mhd_operator_t*, int32_t* fused, int32_t* timed_out);

/* symbolic phase: symbolic_loop_matrix! + nz_allocation of Gridap's assembler (called through
 * allocate_jacobian; src/main.jl:222,163).  Builds rowptr/colval and the cell-entry -> nnz scatter map. */
int mhd_operator_symbolic(mhd_operator_t*, int64_t* nrows, int64_t* ncols, int64_t* nnz);
/* copy the pattern out: index_bytes 4|8, base 0|1 (SparseMatrixCSR{0,Float64,PetscInt} / 1-based Int64;
 * src/parameters.jl:224,235) */
int mhd_operator_get_csr(mhd_operator_t*, void* rowptr, void* colval, int index_bytes, int base);
int mhd_operator_get_scatter_stats(mhd_operator_t*, int64_t* nentries, int64_t* nexclusive);
/* The enumeration of a cell's touched entries used by the scatter map = the order in which the Jacobian kernel
 * consumes it: order[e] = (row slot << 8 | col slot) in the permuted local numbering (u: c*27+s, p: 81+k, j: 85+s,
 * phi: 121+l), 0xFFFF for the codes that pad a warp job to a multiple of 32.  order == NULL: only *n is returned.
 * Needs no device (used by the CPU tests to pin the symbolic/numeric contract). */
int mhd_map_entry_order(uint16_t* order, int64_t* n);

/* numeric phase: jacobian!(A,op,x) / residual!(b,op,x) (Gridap NonlinearOperator API used by
 * solve!(xh,solver,op), src/main.jl:275; and jacobian(op,xh)/residual(op,xh), src/main.jl:158,163).
 * x: [n local] free values.  nzval_out: [nnz] or NULL (values stay on the device behind the handle). */
int mhd_jacobian(mhd_operator_t*, const double* x, double* nzval_out);
int mhd_residual(mhd_operator_t*, const double* x, double* r_out);
/* residual_and_jacobian!(b,A,op,x) (Gridap NonlinearOperator API): one fused kernel, the cell preparation is shared */
int mhd_residual_and_jacobian(mhd_operator_t*, const double* x, double* r_out);
int mhd_get_nzval(mhd_operator_t*, double* nzval_out);          /* D2H (or D2D) copy of the current values */
int mhd_set_nzval(mhd_operator_t*, const double* nzval);        /* tests: load values assembled elsewhere */

/* ---- Krylov building blocks: mul!(y,A,x), dot, axpy! on PVector/PSparseMatrix as used by
 * GridapSolvers FGMRES (src/Solvers/badia2024.jl:40).  n = local length (rows for dot/axpy). ---- */
int mhd_spmv(mhd_operator_t*, const double* x, double* y); /* y = A x (halo exchange first when ranks>1) */
int mhd_dot(mhd_operator_t*, const double* x, const double* y, double* result /* host or device */);
int mhd_axpy(mhd_operator_t*, double a, const double* x, double* y); /* y += a x */
/* fused Gram-Schmidt step of FGMRES: h[i] = <w, V_i>, i<k (one pass), then w -= sum h[i] V_i.
 * V: k vectors of leading dimension ldv (device or host); h: k doubles out. */
int mhd_multi_dot_axpy(mhd_operator_t*, int32_t k, const double* V, int64_t ldv, double* w, double* h);

/* ---- linear solver: symbolic_setup/numerical_setup/numerical_setup!/solve! of a Gridap LinearSolver
 * (seam: _solver / get_block_solver, src/main.jl:181-190, src/Solvers/gridap.jl:2-3). ---- */
int mhd_solver_default_opts(mhd_solver_opts_t*);
int mhd_solver_create(mhd_operator_t*, const mhd_solver_opts_t*, mhd_solver_t** out);
/* Vertex-star patches for MHD_UJ_GMRES_PATCH: what Geometry.PatchTopology(ReferenceFE{0},model) + the FE space give the
 * reference's BlockJacobiSolver (src/Solvers/gmg.jl:69-73).  Patch k owns the rows patch_dofs[patch_ptr[k]..patch_ptr[k+1])
 * (0-based local row ids inside the (u,j) block, strictly increasing, at most 256 per patch; empty patches allowed).
 * Call before mhd_solver_setup; the lists are copied. */
int mhd_solver_set_patches(mhd_solver_t*, int64_t npatch, const int64_t* patch_ptr, const int32_t* patch_dofs);
/* MHD_PC_H1H1_BLOCKS: vertex-star patches of the phi block (global row ids, i.e. >= n_u + n_p); the u patches go through
 * mhd_solver_set_patches */
int mhd_solver_set_phi_patches(mhd_solver_t*, int64_t npatch, const int64_t* patch_ptr, const int32_t* patch_dofs);
int mhd_solver_setup(mhd_solver_t*); /* numerical_setup!: refresh preconditioner data after mhd_jacobian */
/* solve!(z, ns::BlockJacobiSolver, r) of the patch solver alone: z = omega * sum_p R_p' inv(A_p) R_p r on the (u,j) block
 * (r, z: [n_uj], host or device).  After mhd_solver_set_patches + mhd_solver_setup. */
int mhd_solver_patch_apply(mhd_solver_t*, const double* r, double* z, double omega);
int mhd_solve(mhd_solver_t*, const double* b, double* x /* in: x0, out: solution */, int32_t* iters,
              double* resnorm, double* res_history /* nullable, maxiter+1 doubles */);
int mhd_solver_destroy(mhd_solver_t*);

/* ---- post-processing of GridapMHD.hunt (src/Applications/hunt.jl:239-260): norms of the discrete solution and its
 * errors against the analytical Hunt series analytical_hunt_u / analytical_hunt_j (hunt.jl:372-457, square duct b = a),
 * integrated with the tables `tab6` of the degree 2*(order+1) rule (nq <= 64; phi_val = the 8 vertex functions of the
 * geometry map).  x: [n local] free values of the dimensionless solution; uh = u0 * ubar_h, jh = jscale * jbar_h
 * (hunt.jl:212-217).  out[6] = { eu_l2, eu_h1, ej_l2, uh_l2, uh_h1, jh_l2 } over the LOCAL cells (square roots taken;
 * with several ranks the caller sums the squares of the owned-cell results). */
typedef struct {
  double a, mu, sigma, grad_pz, Ha; /* semi-width, viscosity rho*nu, conductivity, pressure gradient -f_z/rho, Hartmann number */
  int32_t nsums, reserved;          /* series terms k = 0..nsums */
  double u0, jscale;
} mhd_hunt_post_t;
int mhd_hunt_error_norms(mhd_operator_t*, const double* x, const mhd_tables_t* tab6, const mhd_hunt_post_t* prm, double* out6);

/* ---- H1-H1 formulation (u Q2, p P1disc, phi Q3 continuous; j = sigma (u x B - grad phi) eliminated):
 * weak_form_h1_h1 -> jac_fluid_h1_h1 / res_fluid_h1_h1 (src/weakforms.jl:344-355,415-466), solid cells :468-478,
 * spaces of the `formulation in (:H1H1,:HDivH1)` branch of setup_fe_spaces (src/fespaces.jl:32-41) with
 * reffe_phi = LagrangianRefFE(Float64,HEX,k+1;space=:Q), conformity :H1 (src/parameters.jl:528-532),
 * layout _multi_field_style(::Val{:h1h1blocks}) = (u,p,phi) (src/fespaces.jl:9).
 * The handle is an mhd_operator_t: mhd_operator_symbolic / get_csr / mhd_jacobian / mhd_residual /
 * mhd_residual_and_jacobian / mhd_spmv / mhd_dot / mhd_axpy / mhd_multi_dot_axpy / mhd_solver_* (MHD_PC_NONE and
 * MHD_PC_JACOBI) / halo calls work on it as documented above.  Local dofs of a cell: u (a + 27 c) | p | phi (64, in
 * whatever local order the host tabulated phi_grad), 149 in total; touched blocks uu, up, u-phi, pu, phi-u, phi-phi. */
typedef struct {
  int32_t nq;             /* must be 27: Quadrature(HEX,5), q = max(2,5,4,4,2*(3-1)) (src/parameters.jl:382-388) */
  const double* w;        /* [nq] */
  const double* geo_grad; /* [nq*8*3] */
  const double* u_val;    /* [nq*27] */
  const double* u_grad;   /* [nq*27*3] */
  const double* p_val;    /* [nq*4] */
  const double* phi_grad; /* [nq*64*3] reference gradients of the scalar Q3 basis */
} mhd_tables_h1h1_t;
typedef struct {
  const int32_t* cell_dofs[3]; /* u [ncells*81], p [ncells*4], phi [ncells*64]; signed 1-based, 0 = absent (u, p on solid cells) */
  int64_t nfree[3], nowned[3], ndir[3];
  const double* dir_values[3];
  int32_t field_order[3];      /* permutation of {0 = u, 1 = p, 2 = phi} */
} mhd_layout_h1h1_t;
int mhd_h1h1_operator_create(const mhd_mesh_t*, const mhd_tables_h1h1_t*, const mhd_layout_h1h1_t*, const mhd_params_t*,
                             mhd_operator_t** out);
/* enumeration of the touched entries of an H1-H1 cell (row << 8 | col, local numbering u|p|phi), = scatter-map order;
 * needs no device */
int mhd_h1h1_entry_order(uint16_t* order, int64_t* n);

/* ---- introspection for tests / benches ---- */
int mhd_operator_device_ptrs(mhd_operator_t*, void** rowptr_i64, void** colval_i32, void** nzval_f64);
int mhd_kernel_launch_count(int64_t* count); /* kernels launched by the library since mhd_init */
/* CUDA-event timing of the named kernels on the library stream ("jacobian", "residual", "spmv"): enable, run,
 * then read the accumulated device time and launch count (synchronises the stream). */
int mhd_profile_enable(int on);
int mhd_profile_get(const char* name, double* total_ms, int64_t* launches);
int mhd_profile_reset(void);
/* FP64 peak of this device, measured (no FP64 figure in MEASURED_PEAKS.json): register-resident
 * mma.sync.m8n8k4.f64 chains (kind 0) or DFMA chains (kind 1) on every SM, CUDA-event timed.  Out: TFLOP/s (FMA = 2). */
int mhd_fp64_peak(int32_t kind, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* MHDB200_H */
