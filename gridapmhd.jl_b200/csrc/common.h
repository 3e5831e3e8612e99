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

// canonical enumeration of the touched cell entries (row-major inside each block section)
constexpr int SEC_UU = 0;                       // [81][81]
constexpr int SEC_UP = SEC_UU + NU * NU;        // [81][4]
constexpr int SEC_PU = SEC_UP + NU * NP;        // [4][81]
constexpr int SEC_UJ = SEC_PU + NP * NU;        // [81][36]
constexpr int SEC_JU = SEC_UJ + NU * NJ;        // [36][81]
constexpr int SEC_JJ = SEC_JU + NJ * NU;        // [36][36]
constexpr int SEC_JF = SEC_JJ + NJ * NJ;        // [36][8]
constexpr int SEC_FJ = SEC_JF + NJ * NF;        // [8][36]
constexpr int NENT = SEC_FJ + NF * NJ;          // 14913
constexpr int NENT_PAD = (NENT + 7) / 8 * 8;    // per-cell stride of the u16 scatter map (16 B aligned)

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
constexpr int T_TOTAL = T_CHI + NQ * 8;

struct Comm;  // comm.cu

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
  long long* d_row_bits = nullptr;      // [nrows+1] tagged copy of rowptr (bit 62: row has ghost columns)
  int* d_err = nullptr;
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
  const long long* rowptr_tagged;            // rowptr copy with bit 62 set on rows that end with ghost columns (unused by the vote kernel)
  int* err;
};

}  // namespace mhd

struct mhd_operator {
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

  // device data
  double* d_coords = nullptr;      // [nnodes*3]
  int32_t* d_cell_nodes = nullptr; // [ncells*8] 0-based
  int32_t* d_gids = nullptr;       // [ncells*129] >=0 local free id (owned first, ghosts after); <0: -(dirichlet index+1)
  int8_t* d_jsign = nullptr;       // [ncells*36]
  uint8_t* d_cell_solid = nullptr; // [ncells] or null
  double* d_cell_sigma = nullptr;  // [ncells] or null
  double* d_dir = nullptr;         // [ndir_total]
  double* d_tables = nullptr;      // packed reference tables (see assembly.cu)
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

// event timing of the dominant kernels (bench.py roofline): prof_begin/prof_end bracket a launch
enum { PROF_JAC = 0, PROF_RES = 1, PROF_SPMV = 2, PROF_N = 3 };
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
// assembly.cu
int pack_tables(mhd_operator* op, const mhd_tables_t* t);
int launch_jacobian(mhd_operator* op, const double* d_x, double* d_r /* nullable: fused residual */);
int launch_residual(mhd_operator* op, const double* d_x, double* d_r);
// krylov.cu
int launch_spmv(mhd_operator* op, const double* d_x, double* d_y);
// y[0..nr) = (A x)[0..nr) including the ghost exchange (fused peer-memory kernel when connected, NCCL otherwise)
int spmv_with_halo(mhd_operator* op, int64_t nr, double* d_x, double* d_y);
int launch_dot(mhd_operator* op, int64_t n, const double* d_x, const double* d_y, double* d_out);
int launch_axpy(int64_t n, double a, const double* d_x, double* d_y);
int launch_multi_dot(mhd_operator* op, int64_t n, int k, const double* d_V, int64_t ldv, const double* d_w, double* d_h);
int launch_multi_axpy(int64_t n, int k, const double* d_V, int64_t ldv, const double* d_h, double sign, double* d_w);
int ensure_red(mhd_operator* op, int64_t ndoubles);
// comm.cu
int halo_exchange(mhd_operator* op, double* d_x);
int allreduce_sum(double* d_buf, int n);

}  // namespace mhd
