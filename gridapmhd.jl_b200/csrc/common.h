// Internal definitions shared by the translation units of libmhdb200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mhdb200.h"

namespace mhd {

// ---- local (per cell) layout: u(81: a + 27 c) | p(4) | j(36) | phi(8)
constexpr int NQ = 27;
constexpr int NU = 81, NP = 4, NJ = 36, NF = 8;
constexpr int OFF_U = 0, OFF_P = 81, OFF_J = 85, OFF_F = 121;
constexpr int NLOC = 129;

// Enumeration of the touched cell entries = the order in which the Jacobian kernel consumes the scatter map
// (assembly.cu).  Entries are addressed in the PERMUTED local numbering of a cell: inside each field the local dofs are
// sorted by global id (u: the 27 nodes sorted, slot index c*27+s; j: the 36 dofs sorted), so that a sweep over
// consecutive slots of a row walks consecutive nnz of the CSR row.  The range of every warp job is padded to a multiple
// of 32 entries (pad codes = MAP_SKIP) so that a job's codes are fetched with one address and immediate offsets.
//   uu : warp w = 0..7 (row tile mt = w/2 of 8 node slots, column half np = w%2 of 16|11 node slots), component c:
//        [nrow][ncol] with columns (node slot, component d) component-fastest
//   jj : job s = 0..4: rows 8 s .. (8 | 4 of them) x 36 columns
//   uj : job (s = 0..4: 8|4 j slots, h = 0..1: 16|11 node slots): [c][node][j] then the ju part [j][(node, d)]
//   up : [81][4] ; pu : [4][(node, d)] ; j-phi : [36][8] ; phi-j : [8][36]
__host__ __device__ constexpr int pad32(int n) { return (n + 31) / 32 * 32; }
__host__ __device__ constexpr int uu_nrow(int mt) { return mt < 3 ? 8 : 3; }
__host__ __device__ constexpr int uu_ncol(int np) { return np ? 33 : 48; }
__host__ __device__ constexpr int uj_nm(int s) { return s < 4 ? 8 : 4; }
__host__ __device__ constexpr int uj_na(int h) { return h ? 11 : 16; }
constexpr int SEC_UU = 0;
__host__ __device__ constexpr int uu_base(int w, int c) {
  return SEC_UU + 3 * (pad32(8 * 48) + pad32(8 * 33)) * (w >> 1) + (w & 1) * 3 * pad32(uu_nrow(w >> 1) * 48) +
         c * pad32(uu_nrow(w >> 1) * uu_ncol(w & 1));
}
constexpr int SEC_JJ = uu_base(6, 0) + 3 * (pad32(3 * 48) + pad32(3 * 33));
__host__ __device__ constexpr int jj_base(int s) { return SEC_JJ + pad32(8 * NJ) * s; }
constexpr int SEC_UJ = jj_base(4) + pad32(4 * NJ);
__host__ __device__ constexpr int uj_part(int s, int h) { return pad32(3 * uj_na(h) * uj_nm(s)); }
__host__ __device__ constexpr int uj_base(int s, int h) {
  return SEC_UJ + 2 * (uj_part(0, 0) + uj_part(0, 1)) * (s < 4 ? s : 4) + h * 2 * uj_part(s, 0);
}
constexpr int SEC_UP = uj_base(4, 1) + 2 * uj_part(4, 1);
constexpr int SEC_PU = SEC_UP + pad32(NU * NP);
constexpr int SEC_JF = SEC_PU + pad32(NP * NU);
constexpr int SEC_FJ = SEC_JF + pad32(NJ * NF);
constexpr int NENT = SEC_FJ + pad32(NF * NJ);   // 15584 map codes per cell, 14913 of them real entries
constexpr int NENT_PAD = NENT;                  // per-cell stride of the u16 scatter map (multiple of 32)
constexpr uint16_t ORDER_PAD = 0xFFFFu;         // entry_order() value of a padding code
constexpr int PERM_STRIDE = 64;  // bytes per cell: [0,27) node slot -> reference node, [27,63) j slot -> reference dof | 0x80 if sign < 0

constexpr uint16_t MAP_SKIP = 0xFFFFu;  // entry dropped (Dirichlet row/col or non-owned row)
constexpr uint16_t MAP_EXCL = 0x8000u;  // nnz receives exactly one contribution: plain store, no atomic
constexpr int MAX_ROW_NNZ = 0x7FFF;

// packed reference tables (doubles)
constexpr int T_W = 0;
constexpr int T_GG = T_W + NQ;              // [27][8][3]
constexpr int T_NU = T_GG + NQ * 8 * 3;     // [27][27]
constexpr int T_DNU = T_NU + NQ * 27;       // [27][27][3]
constexpr int T_PP = T_DNU + NQ * 27 * 3;   // [27][4]
constexpr int T_PSI = T_PP + NQ * 4;        // [27][36][3]
constexpr int T_DPSI = T_PSI + NQ * 36 * 3; // [27][36]
constexpr int T_CHI = T_DPSI + NQ * 36;     // [27][8]
constexpr int T_GGT = T_CHI + NQ * 8;       // [8][3][27]  geo_grad[q][v][k] at (v*3+k)*27+q (coalesced across q)
constexpr int T_TOTAL = T_GGT + NQ * 24;
// Second copy of the basis tables in the PANEL layout of the assembly kernels (reference order, rows k-major, padded
// leading dimensions): one bulk copy per cell brings it into shared memory, where it is transformed in place.
constexpr int PT_G = 0;                     // [27 q][3 k][28]  d N_a / d xi_k   (column 27 = 0)
constexpr int PT_N = PT_G + 81 * 28;        // [27 q][28]       N_a
constexpr int PT_PSI = PT_N + 27 * 28;      // [27 q][3 k][36]  psi_m component k
constexpr int PT_DIV = PT_PSI + 81 * 36;    // [27 q][36]       div psi_m
constexpr int PT_PP = PT_DIV + 27 * 36;     // [27 q][4]
constexpr int PT_CHI = PT_PP + 27 * 4;      // [27 q][8]
constexpr int PT_TOTAL = PT_CHI + 27 * 8;   // 7236 doubles = 57888 B (multiple of 16)

struct Comm;  // postprocess.cu
int hunt_error_norms(mhd_operator* op, const double* d_x, const mhd_tables_t* t, const mhd_hunt_post_t* p, double* out6);
// comm.cu

struct Halo {
  int nneigh = 0;
  std::vector<int> ranks;
  std::vector<int64_t> send_ptr, recv_ptr;
  int32_t* d_send_idx = nullptr;
  int32_t* d_recv_idx = nullptr;
  double* d_send_buf = nullptr;
  double* d_recv_buf = nullptr;
  int64_t nsend = 0, nrecv = 0;
  // fused SpMV + halo over NVLink peer memory (see krylov.cu: spmv_fused_halo)
  bool fused = false;
  void* ipc_mem = nullptr;              // [flags 2 x 64 u32 | inbox parity 0 | inbox parity 1], exported with CUDA IPC
  std::vector<void*> peer_mem;          // the neighbours' ipc_mem, opened with cudaIpcOpenMemHandle
  struct HaloDev* d_dev = nullptr;      // device copy of the plan
  int32_t* d_ghost_src = nullptr;       // [nsend] destination ghost slots on the neighbours (send_dst)
  long long* d_row_bits = nullptr;      // [nrows+1] position of the first ghost column of every row (== end of the row if it has none)
  int32_t* d_if_rows = nullptr;         // [n_if_rows] interface rows (rows with at least one ghost column), ascending
  int64_t n_if_rows = 0;
  int* d_err = nullptr;
  const double* inbox[2] = {nullptr, nullptr};  // this rank's inboxes and arrival counters (inside ipc_mem): kernel parameters
  const unsigned* flags = nullptr;
  unsigned epoch = 0;
};

constexpr int HALO_MAX_NEIGH = 64;
constexpr int HALO_FLAG_BYTES = 2 * HALO_MAX_NEIGH * 4;
constexpr int HALO_CHUNK = 2048;        // ghost values pushed per CTA

struct HaloDev {
  int nneigh, npush;
  int64_t nrecv;
  double* peer_inbox[2][HALO_MAX_NEIGH];     // where my values for neighbour k land (parity p), already offset
  unsigned* peer_flag[2][HALO_MAX_NEIGH];    // my arrival counter inside neighbour k's memory
  unsigned expected[HALO_MAX_NEIGH];         // CTAs neighbour k uses to push to me
  int64_t send_begin[HALO_MAX_NEIGH + 1];
  const int32_t* send_idx;
  const int32_t* send_dst;                   // [nsend] ghost slot (ghost id - nrows) of each sent value on its neighbour
  const int* push_neigh;                     // [npush] neighbour of each push CTA
  const int64_t* push_begin;                 // [npush] first send-list entry of each push CTA
  const unsigned* my_flags;                  // [2][HALO_MAX_NEIGH]
  const double* my_inbox[2];
  const long long* rowptr_tagged;            // ghost_lo[row]: first ghost entry of the row (ghost columns are sorted last)
  int* err;
};

}  // namespace mhd

namespace mhd {
constexpr int FORM_HDIV = 0;  // H1-HDiv (u,p,j,phi), 129 local dofs: assembly.cu
constexpr int FORM_H1H1 = 1;  // H1-H1 (u,p,phi), 149 local dofs: h1h1.cu (field slots: 0 = u, 1 = p, 3 = phi; slot 2 empty)
}  // namespace mhd

struct mhd_operator {
  int formulation = mhd::FORM_HDIV;
  int jac_version = 5;             // 5: tensor-core panel products (assembly.cu, any tables);
                                   // 7: fully sum-factorised kernel (hdiv_v7.cu), chosen when the tables have the tensor structure
  unsigned char* d_tab7 = nullptr; // v7: h7::Tab7
  std::vector<unsigned char> h_small7;  // v7: h7::Small7 (uploaded to constant memory before a launch when it is not the current one)
  uint32_t* d_shared_mask = nullptr;  // v7: bit s set <=> the 32-byte sector s of nzval holds an nnz with != 1 contributions (cleared before an assembly)
  bool deterministic = false;      // v7: one launch per colour => run-to-run identical bits
  int32_t* d_color_cells = nullptr;   // cells sorted by colour
  int32_t* d_cell_order = nullptr;    // v7: traversal order of the cells (breadth-first over face neighbours), see v7_build_cell_order
  bool cell_order_tried = false;
  std::vector<int64_t> color_ptr;  // [ncolors + 1]
  std::vector<double> h_tables;    // host copy of the packed reference tables (T_* layout)
  int64_t ncells = 0, nnodes = 0;
  int64_t nfree[4] = {0, 0, 0, 0}, nowned[4] = {0, 0, 0, 0}, ndir[4] = {0, 0, 0, 0};
  int32_t field_order[4] = {0, 1, 2, 3};
  int64_t own_off[4], ghost_off[4], dir_off[4];
  int64_t nrows = 0;  // owned free dofs (matrix rows)
  int64_t ncols = 0;  // owned + ghost free dofs (vector length)
  int64_t nnz = 0;
  int64_t ndir_total = 0;
  mhd_params_t prm;
  bool has_symbolic = false;
  bool clear_pending = false;      // begin_clear() has been enqueued for the next launch_jacobian

  // device data
  double* d_coords = nullptr;      // [nnodes*3]
  int32_t* d_cell_nodes = nullptr; // [ncells*8] 0-based
  int32_t* d_gids = nullptr;       // [ncells*129] >=0 local free id (owned first, ghosts after); <0: -(dirichlet index+1)
  int8_t* d_jsign = nullptr;       // [ncells*36]
  int32_t* d_pgids = nullptr;      // [ncells*129] d_gids in the permuted local numbering (fields sorted by global id)
  uint8_t* d_perm = nullptr;       // [ncells*PERM_STRIDE] slot -> reference basis index (+ RT sign bit)
  int64_t* d_rowstart = nullptr;   // [ncells*129] first nnz of each local row in the permuted numbering (-1: dropped row)
  uint16_t* d_order = nullptr;     // [NENT] (row slot << 8 | col slot) of each map entry, kernel consumption order
  uint8_t* d_cell_solid = nullptr; // [ncells] or null
  double* d_cell_sigma = nullptr;  // [ncells] or null
  double* d_dir = nullptr;         // [ndir_total]
  double* d_tables = nullptr;      // packed reference tables (see assembly.cu)
  double* d_ptab = nullptr;        // the same tables in panel layout (PT_*)
  int64_t* d_rowptr = nullptr;     // [nrows+1]
  int32_t* d_colval = nullptr;     // [nnz]
  double* d_nzval = nullptr;       // [nnz]
  uint16_t* d_map = nullptr;       // [ncells*NENT_PAD]
  int64_t nentries = 0, nexclusive = 0;
  // scratch
  double* d_x = nullptr;           // [ncols] staging for host-pointer calls
  double* d_y = nullptr;           // [ncols]
  double* d_red = nullptr;         // reduction scratch
  int64_t red_cap = 0;
  mhd::Halo halo;
};

namespace mhd {

extern cudaStream_t g_stream;
extern int g_device;
extern int64_t g_launches;
extern int g_nranks, g_rank;

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define MHD_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) return mhd::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define MHD_CHECK(cond, code, ...) \
  do {                             \
    if (!(cond)) {                 \
      mhd::set_error(__VA_ARGS__); \
      return (code);               \
    }                              \
  } while (0)

#define MHD_TRY(call)        \
  do {                       \
    int _rc = (call);        \
    if (_rc != 0) return _rc; \
  } while (0)

#define MHD_LAUNCH_CHECK()                    \
  do {                                        \
    mhd::g_launches++;                        \
    MHD_CUDA(cudaPeekAtLastError());          \
  } while (0)

bool is_device_ptr(const void* p);
int device_sm_count();  // multiprocessors of g_device (queried once per mhd_init; abi.cu)

// event timing of the dominant kernels (bench.py roofline): prof_begin/prof_end bracket a launch
enum { PROF_JAC = 0, PROF_RES = 1, PROF_SPMV = 2, PROF_PATCH_SETUP = 3, PROF_PATCH_APPLY = 4, PROF_N = 5 };
extern bool g_prof_on;
void prof_begin(int which);
void prof_end(int which);

template <class T>
int dev_alloc(T** p, int64_t n) {
  *p = nullptr;
  if (n <= 0) n = 1;
  MHD_CUDA(cudaMalloc((void**)p, (size_t)n * sizeof(T)));
  return 0;
}
template <class T>
int h2d(T* d, const T* h, int64_t n) {
  if (n > 0) MHD_CUDA(cudaMemcpyAsync(d, h, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, g_stream));
  return 0;
}
template <class T>
int d2h(T* h, const T* d, int64_t n) {
  if (n > 0) MHD_CUDA(cudaMemcpyAsync(h, d, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, g_stream));
  return 0;
}

// symbolic.cu
int symbolic_build(mhd_operator* op);
int build_permutation(mhd_operator* op);  // d_perm, d_pgids (needs d_gids, d_jsign)
// assembly.cu
void entry_order(std::vector<uint16_t>& ord);  // [NENT] (row slot << 8 | col slot), mirrors the kernel's sweeps
// assembly.cu
int pack_tables(mhd_operator* op, const mhd_tables_t* t);
int launch_jacobian(mhd_operator* op, const double* d_x, double* d_r /* nullable: fused residual */);
int begin_clear(mhd_operator* op, double* d_r /* nullable */, bool matrix = true, bool side = true);
int end_clear();  // optional: start clearing before the state is copied in
void assembly_finalize();
int launch_residual(mhd_operator* op, const double* d_x, double* d_r);
// hdiv_v7.cu
void v7_entry_order(std::vector<uint16_t>& ord);
int v7_try_enable(mhd_operator* op);                                  // at operator creation: discovers the tensor structure of the tables
int v7_build_shared_mask(mhd_operator* op, const uint8_t* d_contrib); // end of the symbolic phase
int v7_build_cell_order(mhd_operator* op);                            // lazily, before the first launch
int v7_zero_shared(mhd_operator* op, cudaStream_t stream, double* d_r /* nullable: residual cleared by the same launch */);             // clears the sectors of nzval that hold shared nnz
int v7_launch(mhd_operator* op, const double* d_x, double* d_r, int mode /* 0: Jacobian, 1: residual + Jacobian, 2: residual */);
// h1h1.cu
int h1h1_launch_jacobian(mhd_operator* op, const double* d_x, double* d_r /* nullable: fused residual */);
int h1h1_launch_residual(mhd_operator* op, const double* d_x, double* d_r);
// patch.cu: vertex-patch block-Jacobi smoother of the (u,j) block
struct PatchData;
int patch_create(PatchData** out, int64_t n_uj, int64_t npatch, const int64_t* ptr, const int32_t* dofs);
void patch_destroy(PatchData* P);
int patch_setup(PatchData* P, mhd_operator* op);  // gather the sub-blocks of the current Jacobian and invert them
int patch_apply(PatchData* P, const double* d_r, double* d_z, double omega, bool accumulate);
int64_t patch_bytes(const PatchData* P);
// krylov.cu
int launch_spmv(mhd_operator* op, const double* d_x, double* d_y);
// y[0..nr) = (A x)[0..nr) including the ghost exchange (fused peer-memory kernel when connected, NCCL otherwise)
int spmv_with_halo(mhd_operator* op, int64_t nr, double* d_x, double* d_y);
int spmv_row_range(mhd_operator* op, int64_t r0, int64_t r1, const double* d_x, double* d_y);
int launch_dot(mhd_operator* op, int64_t n, const double* d_x, const double* d_y, double* d_out);
int launch_axpy(int64_t n, double a, const double* d_x, double* d_y);
int launch_multi_dot(mhd_operator* op, int64_t n, int k, const double* d_V, int64_t ldv, const double* d_w, double* d_h);
int launch_multi_axpy(int64_t n, int k, const double* d_V, int64_t ldv, const double* d_h, double sign, double* d_w);
int ensure_red(mhd_operator* op, int64_t ndoubles);
// fused Gram-Schmidt building blocks of FGMRES (one launch each; k + 1 <= 18)
bool gs_fused_ok(int k);
int launch_gs_dots(mhd_operator* op, int64_t n, int k, bool with_norm, const double* d_V, int64_t ldv, const double* d_w, double* d_out,
                   const int* d_run_if = nullptr);
int launch_gs_update(int64_t n, int k, const double* d_V, int64_t ldv, const double* d_h, const double* d_w, const double* d_scale,
                     const int* d_dead, double* d_out, const int* d_reorth = nullptr, int pass = 0, double* d_plain_out = nullptr);
// postprocess.cu
int hunt_error_norms(mhd_operator* op, const double* d_x, const mhd_tables_t* t, const mhd_hunt_post_t* p, double* out6);
// comm.cu
int halo_exchange(mhd_operator* op, double* d_x);
// at a host synchronisation point: MHD_E_COMM if a fused SpMV + halo kernel gave up waiting for a neighbour (the products computed
// since then used stale ghost values)
int halo_check(mhd_operator* op);
int allreduce_sum(double* d_buf, int n);

}  // namespace mhd
