// H1-H1 formulation on the device: operator creation, Jacobian and residual kernels.
//
//   jac_fluid_h1_h1 / res_fluid_h1_h1   src/weakforms.jl:440-466 / :415-438
//   jac_solid_h1_h1 / res_solid_h1_h1   src/weakforms.jl:474-478 / :468-472
//   spaces (u Q2, p P1disc, phi Q3 H1)   src/fespaces.jl:32-41, src/parameters.jl:528-532
//
// One CTA works on one cell at a time (persistent grid of 2 CTAs per SM): the per-cell arithmetic is the sequence of
// barrier-separated phases of h1h1_cell.h (geometry -> physical gradients -> point coefficients -> entries), every
// entry is a 27-point sum read from shared memory, and the values go to the CSR through the u16 scatter map of the
// symbolic phase (row-relative position, exclusive entries stored plainly, shared ones with atomicAdd) -- the same
// symbolic kernels as H1-HDiv, instantiated for the 149-dof layout (symbolic.cu).
// The sparse matrix, SpMV, dots, axpys and the FGMRES of krylov.cu / solver.cu work on the handle unchanged.
#include "common.h"
#include "h1h1_cell.h"

namespace mhd {

namespace {

constexpr int H1_NT = 256;

struct H1Args {
  const double* coords;
  const int32_t* cell_nodes;
  const int32_t* gids;
  const long long* rowstart;
  const double* dir;
  const double* tab;
};

struct DevStore {
  const uint16_t* map;
  double* nz;
  const long long* rowstart;  // shared memory
  __device__ __forceinline__ void operator()(int e, int li, int /*lj*/, double v) const {
    const uint16_t code = map[e];
    if (code == MAP_SKIP) return;
    double* p = nz + rowstart[li] + (code & 0x7FFF);
    if (code & MAP_EXCL) *p = v;
    else atomicAdd(p, v);
  }
};

struct DevAdd {
  const int32_t* gid;  // shared memory
  double* r;
  int64_t nrows;
  __device__ __forceinline__ void operator()(int li, double v) const {
    const int32_t g = gid[li];
    if (g >= 0 && g < nrows) atomicAdd(r + g, v);
  }
};

// RES: also assemble the residual (residual_and_jacobian!): the preparation phases are shared, the residual adds three
// short phases after the entries.  Enabled with MHD_H1H1_FUSED=1 until it has been measured on a B200 (next round).
template <int CONV, bool ZU, bool RES>
__global__ void __launch_bounds__(H1_NT, 2)
h1h1_jacobian_kernel(int64_t ncells, int64_t nrows, H1Args A, const double* __restrict__ x, const uint16_t* __restrict__ map,
                     double* __restrict__ nzval, double* __restrict__ r, h1::Params P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  h1::Shared& S = *reinterpret_cast<h1::Shared*>(smem_raw);
  const int tid = threadIdx.x;
  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    h1::phase_load(S, tid, H1_NT, A.coords, A.cell_nodes + cell * 8, A.gids + cell * h1::NLOC, A.rowstart + cell * h1::NLOC,
                   A.dir, x, A.tab);
    __syncthreads();
    h1::phase_geometry(S, tid, H1_NT, A.tab);
    __syncthreads();
    h1::phase_gradients(S, tid, H1_NT, A.tab);
    __syncthreads();
    if (CONV != 0 || RES) {
      h1::phase_point_values(S, tid, H1_NT);
      __syncthreads();
    }
    h1::phase_jac_coefficients<CONV, ZU>(S, tid, H1_NT, P);
    __syncthreads();
    if (ZU) {
      h1::phase_jac_projection(S, tid, H1_NT);
      __syncthreads();
    }
    DevStore store{map + cell * h1::NENT_PAD, nzval, S.rowstart};
    h1::phase_jac_entries<CONV, ZU>(S, tid, H1_NT, P, store);
    __syncthreads();  // the next cell (or the residual phases) overwrite the shared data
    if (RES) {
      h1::phase_res_points<ZU>(S, tid, H1_NT);
      __syncthreads();
      h1::phase_res_coefficients<(CONV != 0 ? 1 : 0), ZU>(S, tid, H1_NT, P);
      __syncthreads();
      DevAdd add{S.gid, r, nrows};
      h1::phase_res_rows(S, tid, H1_NT, add);
      __syncthreads();
    }
  }
}

template <int CONV, bool ZU>
__global__ void __launch_bounds__(H1_NT, 2)
h1h1_residual_kernel(int64_t ncells, int64_t nrows, H1Args A, const double* __restrict__ x, double* __restrict__ r,
                     h1::Params P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  h1::Shared& S = *reinterpret_cast<h1::Shared*>(smem_raw);
  const int tid = threadIdx.x;
  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    h1::phase_load(S, tid, H1_NT, A.coords, A.cell_nodes + cell * 8, A.gids + cell * h1::NLOC, nullptr, A.dir, x, A.tab);
    __syncthreads();
    h1::phase_geometry(S, tid, H1_NT, A.tab);
    __syncthreads();
    h1::phase_gradients(S, tid, H1_NT, A.tab);
    __syncthreads();
    h1::phase_point_values(S, tid, H1_NT);
    __syncthreads();
    h1::phase_res_points<ZU>(S, tid, H1_NT);
    __syncthreads();
    h1::phase_res_coefficients<CONV, ZU>(S, tid, H1_NT, P);
    __syncthreads();
    DevAdd add{S.gid, r, nrows};
    h1::phase_res_rows(S, tid, H1_NT, add);
    __syncthreads();
  }
}

h1::Params make_params(const mhd_params_t& p) {
  h1::Params P;
  P.alpha = p.alpha;
  P.beta = p.beta;
  P.gamma = p.gamma;
  P.zeta_u = p.zeta_u;
  for (int i = 0; i < 3; i++) {
    P.B[i] = p.B[i];
    P.f[i] = p.f[i];
  }
  return P;
}

H1Args make_args(const mhd_operator* op) {
  return H1Args{op->d_coords, op->d_cell_nodes, op->d_gids, (const long long*)op->d_rowstart, op->d_dir, op->d_tables};
}

unsigned persistent_grid(int64_t ncells) {
  const int64_t g = (int64_t)device_sm_count() * 2;
  return (unsigned)(ncells < g ? ncells : g);
}

template <class K>
int opt_in_smem(K kernel) {
  MHD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(h1::Shared)));
  return 0;
}

}  // namespace

int h1h1_launch_jacobian(mhd_operator* op, const double* d_x, double* d_r) {
  MHD_CUDA(cudaMemsetAsync(op->d_nzval, 0, (size_t)op->nnz * sizeof(double), g_stream));
  if (d_r) MHD_CUDA(cudaMemsetAsync(d_r, 0, (size_t)op->nrows * sizeof(double), g_stream));
  const h1::Params P = make_params(op->prm);
  const unsigned grid = persistent_grid(op->ncells);
  const size_t smem = sizeof(h1::Shared);
  const int conv = op->prm.convection;
  const bool zu = op->prm.zeta_u != 0.0;
#define JKR(C, Z, R)                                                                                              \
  do {                                                                                                            \
    MHD_TRY(opt_in_smem(h1h1_jacobian_kernel<C, Z, R>));                                                          \
    h1h1_jacobian_kernel<C, Z, R><<<grid, H1_NT, smem, g_stream>>>(op->ncells, op->nrows, make_args(op), d_x, op->d_map, \
                                                                   op->d_nzval, d_r, P);                          \
  } while (0)
#define JK(C, Z) do { if (d_r) JKR(C, Z, true); else JKR(C, Z, false); } while (0)
  prof_begin(PROF_JAC);
  if (conv == 0) { if (zu) JK(0, true); else JK(0, false); }
  else if (conv == 1) { if (zu) JK(1, true); else JK(1, false); }
  else { if (zu) JK(2, true); else JK(2, false); }
#undef JK
#undef JKR
  prof_end(PROF_JAC);
  MHD_LAUNCH_CHECK();
  return 0;
}

int h1h1_launch_residual(mhd_operator* op, const double* d_x, double* d_r) {
  MHD_CUDA(cudaMemsetAsync(d_r, 0, (size_t)op->nrows * sizeof(double), g_stream));
  const h1::Params P = make_params(op->prm);
  const unsigned grid = persistent_grid(op->ncells);
  const size_t smem = sizeof(h1::Shared);
  const bool zu = op->prm.zeta_u != 0.0;
#define RK(C, Z)                                                                                              \
  do {                                                                                                        \
    MHD_TRY(opt_in_smem(h1h1_residual_kernel<C, Z>));                                                         \
    h1h1_residual_kernel<C, Z><<<grid, H1_NT, smem, g_stream>>>(op->ncells, op->nrows, make_args(op), d_x, d_r, P); \
  } while (0)
  prof_begin(PROF_RES);
  if (op->prm.convection == 0) { if (zu) RK(0, true); else RK(0, false); }
  else { if (zu) RK(1, true); else RK(1, false); }  // picard and newton share the residual
#undef RK
  prof_end(PROF_RES);
  MHD_LAUNCH_CHECK();
  return 0;
}

}  // namespace mhd

using namespace mhd;

extern "C" {

int mhd_h1h1_entry_order(uint16_t* order, int64_t* n) {
  MHD_CHECK(n != nullptr, MHD_E_INVALID, "mhd_h1h1_entry_order: null argument");
  *n = h1::NENT;
  if (order)
    for (int e = 0; e < h1::NENT; e++) {
      int li, lj;
      h1::entry_rowcol(e, &li, &lj);
      order[e] = (uint16_t)(li << 8 | lj);
    }
  return MHD_OK;
}

int mhd_h1h1_operator_create(const mhd_mesh_t* mesh, const mhd_tables_h1h1_t* tab, const mhd_layout_h1h1_t* lay,
                             const mhd_params_t* prm, mhd_operator_t** out) {
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CHECK(mesh && tab && lay && prm && out, MHD_E_INVALID, "mhd_h1h1_operator_create: null argument");
  MHD_CHECK(tab->nq == h1::NQ, MHD_E_INVALID, "mhd_h1h1_operator_create: nq=%d, only the 27-point rule (q=5) is supported",
            tab->nq);
  MHD_CHECK(tab->w && tab->geo_grad && tab->u_val && tab->u_grad && tab->p_val && tab->phi_grad, MHD_E_INVALID,
            "mhd_h1h1_operator_create: null table");
  MHD_CHECK(mesh->ncells > 0 && mesh->nnodes > 0, MHD_E_INVALID, "mhd_h1h1_operator_create: empty mesh");
  MHD_CHECK(mesh->index_base == 0 || mesh->index_base == 1, MHD_E_INVALID, "index_base must be 0 or 1");
  MHD_CHECK(prm->convection >= 0 && prm->convection <= 2, MHD_E_INVALID, "invalid convection mode %d", prm->convection);
  MHD_CUDA(cudaSetDevice(g_device));
  const int slot[3] = {MHD_FIELD_U, MHD_FIELD_P, MHD_FIELD_PHI};  // ABI field index -> operator field slot
  int seen[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) {
    const int f = lay->field_order[i];
    MHD_CHECK(f >= 0 && f < 3 && !seen[f], MHD_E_INVALID, "field_order is not a permutation of 0..2");
    seen[f] = 1;
  }
  mhd_operator* op = new mhd_operator();
  op->formulation = FORM_H1H1;
  op->ncells = mesh->ncells;
  op->nnodes = mesh->nnodes;
  op->prm = *prm;
  for (int i = 0; i < 3; i++) {
    const int s = slot[i];
    op->nfree[s] = lay->nfree[i];
    op->nowned[s] = lay->nowned[i] > 0 || lay->nfree[i] == 0 ? lay->nowned[i] : lay->nfree[i];
    op->ndir[s] = lay->ndir[i];
    if (op->nowned[s] > op->nfree[s] || op->nfree[s] < 0 || op->ndir[s] < 0) {
      set_error("inconsistent nfree/nowned/ndir for field %d", i);
      delete op;
      return MHD_E_INVALID;
    }
  }
  for (int i = 0; i < 3; i++) op->field_order[i] = slot[lay->field_order[i]];
  op->field_order[3] = MHD_FIELD_J;  // empty slot
  int64_t own = 0, gh = 0, dir = 0;
  for (int i = 0; i < 4; i++) {
    const int f = op->field_order[i];
    op->own_off[f] = own;
    own += op->nowned[f];
  }
  for (int i = 0; i < 4; i++) {
    const int f = op->field_order[i];
    op->ghost_off[f] = own + gh;
    gh += op->nfree[f] - op->nowned[f];
  }
  for (int f = 0; f < 4; f++) {
    op->dir_off[f] = dir;
    dir += op->ndir[f];
  }
  op->nrows = own;
  op->ncols = own + gh;
  op->ndir_total = dir;
  if (op->ncols >= (int64_t)INT32_MAX) {
    set_error("local vector length %lld exceeds int32 column indices", (long long)op->ncols);
    delete op;
    return MHD_E_CAPACITY;
  }
  // ---- signed per-field ids -> one int32 table (>= 0 local free id, < 0: -(index into the Dirichlet values)-1)
  const int nd[3] = {h1::NU, h1::NP, h1::NF};
  const int lo[3] = {h1::OFF_U, h1::OFF_P, h1::OFF_F};
  std::vector<int32_t> gids((size_t)op->ncells * h1::NLOC);
  for (int i = 0; i < 3; i++) {
    const int f = slot[i];
    const int32_t* cd = lay->cell_dofs[i];
    if (!cd) {
      set_error("cell_dofs[%d] is null", i);
      delete op;
      return MHD_E_INVALID;
    }
    for (int64_t c = 0; c < op->ncells; c++)
      for (int k = 0; k < nd[i]; k++) {
        const int32_t id = cd[c * nd[i] + k];
        int32_t g;
        if (id > 0) {
          if (id > op->nfree[f]) {
            set_error("cell %lld field %d: dof id %d > nfree %lld", (long long)c, i, id, (long long)op->nfree[f]);
            delete op;
            return MHD_E_INVALID;
          }
          g = id <= op->nowned[f] ? (int32_t)(op->own_off[f] + id - 1) : (int32_t)(op->ghost_off[f] + (id - 1 - op->nowned[f]));
        } else if (id < 0) {
          if (-id > op->ndir[f]) {
            set_error("cell %lld field %d: Dirichlet id %d beyond ndir %lld", (long long)c, i, id, (long long)op->ndir[f]);
            delete op;
            return MHD_E_INVALID;
          }
          g = -(int32_t)(op->dir_off[f] + (-id - 1)) - 1;
        } else {
          // id 0: absent dof (u, p on solid cells) = Dirichlet dof with value 0 (one extra zero ends the value array)
          if (!(mesh->cell_solid && mesh->cell_solid[c] && i < 2)) {
            set_error("cell %lld field %d: dof id 0 is only valid for u/p on solid cells", (long long)c, i);
            delete op;
            return MHD_E_INVALID;
          }
          g = -(int32_t)dir - 1;
        }
        gids[(size_t)c * h1::NLOC + lo[i] + k] = g;
      }
  }
  std::vector<int32_t> cn((size_t)op->ncells * 8);
  for (size_t i = 0; i < cn.size(); i++) {
    const int32_t v = mesh->cell_nodes[i] - mesh->index_base;
    if (v < 0 || v >= mesh->nnodes) {
      set_error("cell_nodes[%zu]=%d out of range", i, mesh->cell_nodes[i]);
      delete op;
      return MHD_E_INVALID;
    }
    cn[i] = v;
  }
  std::vector<double> dirv((size_t)dir + 1, 0.0);
  for (int i = 0; i < 3; i++) {
    const int f = slot[i];
    if (op->ndir[f] > 0 && lay->dir_values[i]) memcpy(&dirv[op->dir_off[f]], lay->dir_values[i], op->ndir[f] * sizeof(double));
  }
  // ---- packed tables (h1h1_cell.h T_*): gradients direction-major
  std::vector<double> T(h1::T_TOTAL);
  h1::pack_tables(tab->w, tab->geo_grad, tab->u_val, tab->u_grad, tab->p_val, tab->phi_grad, T.data());
  int rc = 0;
#define CR(x) if (!rc) rc = (x)
  CR(dev_alloc(&op->d_coords, op->nnodes * 3));
  CR(dev_alloc(&op->d_cell_nodes, op->ncells * 8));
  CR(dev_alloc(&op->d_gids, op->ncells * h1::NLOC));
  CR(dev_alloc(&op->d_dir, dir + 1));
  CR(dev_alloc(&op->d_tables, (int64_t)h1::T_TOTAL));
  CR(dev_alloc(&op->d_x, op->ncols));
  CR(dev_alloc(&op->d_y, op->ncols));
  CR(h2d(op->d_coords, mesh->coords, op->nnodes * 3));
  CR(h2d(op->d_cell_nodes, cn.data(), op->ncells * 8));
  CR(h2d(op->d_gids, gids.data(), op->ncells * h1::NLOC));
  CR(h2d(op->d_dir, dirv.data(), dir + 1));
  CR(h2d(op->d_tables, T.data(), (int64_t)h1::T_TOTAL));
  CR(ensure_red(op, 4096 + 65 * 1024));
  if (!rc && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "sync", __FILE__, __LINE__);
#undef CR
  if (rc) {
    mhd_operator_destroy(op);
    return rc;
  }
  *out = op;
  return MHD_OK;
}

}  // extern "C"
