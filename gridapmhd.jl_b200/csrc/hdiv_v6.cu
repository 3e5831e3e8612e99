// "v6" Jacobian kernel of the H1-HDiv formulation: the structure-exploiting cell code of hdiv_cell.h (sum-factorised uu block,
// K_ju from K_uj, symmetric jj) run by one CTA of 256 threads per cell on a persistent grid, values scattered through its
// own u16 map (enumeration h6::entry_rowcol, unpermuted local numbering).
//
// STATUS: opt-in (MHD_JAC_V6=1 when the operator is created + mhd_operator_set_tensor_structure).  The cell code is verified
// on the CPU emulation (tests/test_hdiv_v6_host.py); this wrapper and the kernel have NOT run on a GPU yet -- first thing to
// do next round: MHD_RUN_V6_TESTS=1 python -m pytest tests/test_hdiv_v6_gpu.py, then time it against assembly.cu.
#include <stdlib.h>

#include "common.h"
#include "hdiv_cell.h"

namespace mhd {

namespace {

constexpr int V6_NT = 256;

struct V6Args {
  const double* coords;
  const int32_t* cell_nodes;
  const int32_t* gids;
  const long long* rowstart;
  const int8_t* jsign;
  const uint8_t* cell_solid;
  const double* cell_sigma;
  const double* dir;
  const double* tab;
  const sf::Tables* sftab;
};

struct V6Store {
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

template <int CONV, bool ZU, bool ZJ>
__global__ void __launch_bounds__(V6_NT, 2)
hdiv_v6_jacobian_kernel(int64_t ncells, V6Args A, const double* __restrict__ x, const uint16_t* __restrict__ map,
                        double* __restrict__ nzval, h6::Params P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  h6::Shared& S = *reinterpret_cast<h6::Shared*>(smem_raw);
  sf::Tables& T = *reinterpret_cast<sf::Tables*>(smem_raw + sizeof(h6::Shared));
  const int tid = threadIdx.x;
  // the 1-D tables: once per CTA
  for (int i = tid; i < (int)(sizeof(sf::Tables) / 4); i += V6_NT)
    reinterpret_cast<uint32_t*>(&T)[i] = reinterpret_cast<const uint32_t*>(A.sftab)[i];
  __syncthreads();
  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    const bool solid = A.cell_solid != nullptr && A.cell_solid[cell] != 0;
    h6::phase_load(S, tid, V6_NT, A.coords, A.cell_nodes + cell * 8, A.gids + cell * h6::NLOC, A.rowstart + cell * h6::NLOC,
                   A.jsign + cell * h6::NJ, A.dir, x, A.tab, CONV != 0, solid, solid ? A.cell_sigma[cell] : 0.0, P.sigma);
    __syncthreads();
    h6::phase_geometry(S, tid, V6_NT, A.tab);
    __syncthreads();
    h6::phase_mapped_bases<CONV>(S, tid, V6_NT, A.tab);
    __syncthreads();
    h6::phase_coefficients<CONV, ZU>(S, tid, V6_NT, P, A.tab);
    __syncthreads();
    if (ZU) {
      h6::phase_projection(S, tid, V6_NT);
      __syncthreads();
    }
    sf::phase_stage1(S.W, T, tid, V6_NT);
    __syncthreads();
    sf::phase_stage2(S.W, T, tid, V6_NT);
    __syncthreads();
    V6Store store{map + cell * h6::NENT_PAD, nzval, S.rowstart};
    h6::phase_entries<CONV, ZU, ZJ>(S, T, tid, V6_NT, P, A.tab, store);
    __syncthreads();  // the next cell overwrites the shared data
  }
}

constexpr size_t V6_SMEM = sizeof(h6::Shared) + sizeof(sf::Tables);

template <class K>
int v6_opt_in(K kernel) {
  MHD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6_SMEM));
  return 0;
}

}  // namespace

void v6_entry_order(std::vector<uint16_t>& ord) {
  ord.resize(h6::NENT);
  for (int e = 0; e < h6::NENT; e++) {
    int li, lj;
    h6::entry_rowcol(e, &li, &lj);
    ord[e] = (uint16_t)(li << 8 | lj);
  }
}

int v6_launch_jacobian(mhd_operator* op, const double* d_x) {
  MHD_CHECK(op->jac_version == 6 && op->d_sftab != nullptr, MHD_E_STATE, "v6 Jacobian kernel is not enabled on this operator");
  MHD_CUDA(cudaMemsetAsync(op->d_nzval, 0, (size_t)op->nnz * sizeof(double), g_stream));
  h6::Params P;
  P.alpha = op->prm.alpha; P.beta = op->prm.beta; P.gamma = op->prm.gamma; P.sigma = op->prm.sigma;
  P.zeta_u = op->prm.zeta_u; P.zeta_j = op->prm.zeta_j;
  for (int i = 0; i < 3; i++) P.B[i] = op->prm.B[i];
  V6Args A{op->d_coords, op->d_cell_nodes, op->d_gids, (const long long*)op->d_rowstart, op->d_jsign, op->d_cell_solid,
           op->d_cell_sigma, op->d_dir, op->d_tables, (const sf::Tables*)op->d_sftab};
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g_device);
  const int64_t g64 = (int64_t)sms * 2;
  const unsigned grid = (unsigned)(op->ncells < g64 ? op->ncells : g64);
  const int conv = op->prm.convection;
  const bool zu = op->prm.zeta_u != 0.0, zj = op->prm.zeta_j != 0.0;
#define VK(C, U, J)                                                                                               \
  do {                                                                                                            \
    MHD_TRY(v6_opt_in(hdiv_v6_jacobian_kernel<C, U, J>));                                                          \
    hdiv_v6_jacobian_kernel<C, U, J><<<grid, V6_NT, V6_SMEM, g_stream>>>(op->ncells, A, d_x, op->d_map, op->d_nzval, P); \
  } while (0)
#define VKJ(C, U) do { if (zj) VK(C, U, true); else VK(C, U, false); } while (0)
#define VKU(C) do { if (zu) VKJ(C, true); else VKJ(C, false); } while (0)
  prof_begin(PROF_JAC);
  if (conv == 0) VKU(0);
  else if (conv == 1) VKU(1);
  else VKU(2);
#undef VKU
#undef VKJ
#undef VK
  prof_end(PROF_JAC);
  MHD_LAUNCH_CHECK();
  return 0;
}

}  // namespace mhd

using namespace mhd;

extern "C" int mhd_operator_set_tensor_structure(mhd_operator_t* op, const int8_t* node_ijk) {
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CHECK(op != nullptr && node_ijk != nullptr, MHD_E_INVALID, "mhd_operator_set_tensor_structure: null argument");
  MHD_CHECK(op->formulation == FORM_HDIV, MHD_E_INVALID, "mhd_operator_set_tensor_structure: H1-HDiv operators only");
  MHD_CHECK(!op->has_symbolic, MHD_E_STATE, "mhd_operator_set_tensor_structure: call before mhd_operator_symbolic");
  MHD_CHECK((int64_t)op->h_tables.size() >= (int64_t)h6::T_CHI, MHD_E_STATE, "operator holds no host copy of its tables");
  for (int i = 0; i < 81; i++) MHD_CHECK(node_ijk[i] >= 0 && node_ijk[i] <= 2, MHD_E_INVALID, "node_ijk entries must be 0, 1 or 2");
  sf::Tables T;
  const double dev = h6::derive_tensor_tables(op->h_tables.data() + h6::T_NU, op->h_tables.data() + h6::T_DNU, node_ijk, &T);
  MHD_CHECK(dev < 1e-11, MHD_E_INVALID,
            "the velocity tables are not tensor products of 1-D factors for this node map (deviation %.2e)", dev);
  MHD_CUDA(cudaSetDevice(g_device));
  cudaFree(op->d_sftab);
  op->d_sftab = nullptr;
  MHD_CUDA(cudaMalloc(&op->d_sftab, sizeof(sf::Tables)));
  MHD_CUDA(cudaMemcpyAsync(op->d_sftab, &T, sizeof(sf::Tables), cudaMemcpyHostToDevice, g_stream));
  MHD_CUDA(cudaStreamSynchronize(g_stream));
  const char* e = getenv("MHD_JAC_V6");
  op->jac_version = (e && atoi(e) != 0) ? 6 : 5;  // the structure is recorded either way; the kernel is opt-in until timed
  return MHD_OK;
}
