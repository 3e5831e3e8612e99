// Vertex-patch block-Jacobi (additive Schwarz) smoother of the (u,j) block, SURVEY 8 row f1.
//
// Stands behind `PatchBasedSmoothers.BlockJacobiSolver(space, ptopo; assembly = :star)` wrapped in
// `RichardsonSmoother(solver, niter, w)` as built by `gmg_block_jacobi_smoothers` (src/Solvers/gmg.jl:62-81): one patch per
// mesh vertex (`Geometry.PatchTopology(ReferenceFE{0}, model)`, supplied by the host as sorted dof lists), patch matrices
// = sub-blocks of the ASSEMBLED Jacobian (that is what `assembly = :star` block-Jacobi means), every patch solved exactly.
//
//   setup  (per Jacobian) : one CTA per patch gathers its sub-block from the CSR and inverts it in place
//                           (blocked Gauss-Jordan with partial pivoting, patch_cell.h); the explicit inverses stay in
//                           HBM (n_p^2 doubles per patch: 5.1 GB for the 12 675 patches of cfg2 -- 180 GB is what makes
//                           explicit inverses the right trade: the apply becomes one streaming pass)
//   apply                 : z (+)= omega * sum_p R_p^T inv(A_p) R_p r : one CTA per patch, a warp per row, coalesced row
//                           reads, warp-shuffle reduction, atomicAdd into z.  HBM-bound: 8 n_p^2 bytes per patch.
#include <stdlib.h>

#include "common.h"
#include "patch_cell.h"

namespace mhd {

struct PatchData {
  int64_t npatch = 0, n_uj = 0, total = 0;
  int maxn = 0;
  int64_t* d_ptr = nullptr;     // [npatch+1] into d_dofs
  int32_t* d_dofs = nullptr;    // sorted (u,j) row ids of each patch
  int64_t* d_matptr = nullptr;  // [npatch+1] into d_inv (sum of n_p^2)
  double* d_inv = nullptr;
  int* d_flag = nullptr;        // [0] number of singular patches
};

namespace {

#define DEV_PHASE(...)  \
  {                     \
    __VA_ARGS__;        \
    __syncthreads();    \
  }

// gather A[dofs, dofs] from the CSR into M (row-major n x n), then invert in place
template <int MINB>
__global__ void __launch_bounds__(patch::NMAX, MINB)
patch_gather_invert(const int64_t* __restrict__ pptr, const int32_t* __restrict__ pdofs, const int64_t* __restrict__ matptr,
                    const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colval, const double* __restrict__ nzval,
                    double* __restrict__ inv, int* __restrict__ flag) {
  __shared__ patch::Shared S;
  __shared__ int32_t dofs[patch::NMAX];
  const int p = blockIdx.x;
  const int n = (int)(pptr[p + 1] - pptr[p]);
  if (n == 0) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* M = inv + matptr[p];
  if (tid < n) dofs[tid] = pdofs[pptr[p] + tid];
  for (int i = tid; i < n * n; i += nt) M[i] = 0.0;
  __syncthreads();
  // a warp per row: every stored entry of the row whose column is in the patch (binary search in the sorted dof list)
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int i = warp; i < n; i += nw) {
    const int64_t r = dofs[i];
    for (int64_t e = rowptr[r] + lane; e < rowptr[r + 1]; e += 32) {
      const int32_t c = colval[e];
      int lo = 0, hi = n - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (dofs[mid] < c) lo = mid + 1; else hi = mid;
      }
      if (dofs[lo] == c) M[(int64_t)i * n + lo] = nzval[e];
    }
  }
  __syncthreads();
  MHD_PATCH_INVERT(DEV_PHASE, S, M, n);
  if (tid == 0 && S.singular) atomicAdd(flag, 1);
}

// z[dofs] += omega * inv * r[dofs]
__global__ void __launch_bounds__(256)
patch_apply_kernel(const int64_t* __restrict__ pptr, const int32_t* __restrict__ pdofs, const int64_t* __restrict__ matptr,
                   const double* __restrict__ inv, const double* __restrict__ r, double* __restrict__ z, double omega) {
  __shared__ double rp[patch::NMAX];
  __shared__ int32_t dofs[patch::NMAX];
  const int p = blockIdx.x;
  const int n = (int)(pptr[p + 1] - pptr[p]);
  if (n == 0) return;
  const int tid = threadIdx.x;
  if (tid < n) {
    const int32_t d = pdofs[pptr[p] + tid];
    dofs[tid] = d;
    rp[tid] = r[d];
  }
  __syncthreads();
  const double* M = inv + matptr[p];
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < n; i += 8) {
    const double* row = M + (int64_t)i * n;
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s += row[j] * rp[j];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) atomicAdd(z + dofs[i], omega * s);
  }
}

}  // namespace

int patch_create(PatchData** out, int64_t n_uj, int64_t npatch, const int64_t* ptr, const int32_t* dofs) {
  MHD_CHECK(out && ptr && npatch > 0 && ptr[0] == 0, MHD_E_INVALID, "patches: bad arguments");
  std::vector<int64_t> matptr(npatch + 1, 0);
  int maxn = 0;
  for (int64_t p = 0; p < npatch; p++) {
    const int64_t n = ptr[p + 1] - ptr[p];
    MHD_CHECK(n >= 0, MHD_E_INVALID, "patches: patch_ptr is not monotone at patch %lld", (long long)p);
    MHD_CHECK(n <= patch::NMAX, MHD_E_CAPACITY, "patches: patch %lld has %lld dofs, the limit is %d", (long long)p,
              (long long)n, patch::NMAX);
    MHD_CHECK(n == 0 || dofs != nullptr, MHD_E_INVALID, "patches: null dof list");
    for (int64_t i = ptr[p]; i < ptr[p + 1]; i++) {
      MHD_CHECK(dofs[i] >= 0 && dofs[i] < n_uj, MHD_E_INVALID,
                "patches: dof %d of patch %lld is outside the (u,j) block [0,%lld)", dofs[i], (long long)p, (long long)n_uj);
      MHD_CHECK(i == ptr[p] || dofs[i] > dofs[i - 1], MHD_E_INVALID, "patches: dofs of patch %lld are not sorted/unique",
                (long long)p);
    }
    matptr[p + 1] = matptr[p] + n * n;
    if (n > maxn) maxn = (int)n;
  }
  PatchData* P = new PatchData();
  P->npatch = npatch;
  P->n_uj = n_uj;
  P->maxn = maxn;
  P->total = matptr[npatch];
  int rc = 0;
#define CR(x) if (!rc) rc = (x)
  CR(dev_alloc(&P->d_ptr, npatch + 1));
  CR(dev_alloc(&P->d_dofs, ptr[npatch]));
  CR(dev_alloc(&P->d_matptr, npatch + 1));
  CR(dev_alloc(&P->d_inv, P->total));
  CR(dev_alloc(&P->d_flag, 1));
  CR(h2d(P->d_ptr, ptr, npatch + 1));
  CR(h2d(P->d_dofs, dofs, ptr[npatch]));
  CR(h2d(P->d_matptr, matptr.data(), npatch + 1));
  if (!rc && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "sync", __FILE__, __LINE__);
#undef CR
  if (rc) {
    patch_destroy(P);
    return rc;
  }
  *out = P;
  return 0;
}

void patch_destroy(PatchData* P) {
  if (!P) return;
  cudaFree(P->d_ptr);
  cudaFree(P->d_dofs);
  cudaFree(P->d_matptr);
  cudaFree(P->d_inv);
  cudaFree(P->d_flag);
  delete P;
}

int patch_setup(PatchData* P, mhd_operator* op) {
  MHD_CUDA(cudaMemsetAsync(P->d_flag, 0, sizeof(int), g_stream));
  prof_begin(PROF_PATCH_SETUP);
  static int minb = 0;  // resident CTAs per SM the kernel is compiled for (MHD_PATCH_CTAS: A/B builds, default 4)
  if (!minb) {
    const char* e = getenv("MHD_PATCH_CTAS");
    minb = e ? atoi(e) : 4;
  }
  if (minb >= 4)
    patch_gather_invert<4><<<(unsigned)P->npatch, patch::NMAX, 0, g_stream>>>(P->d_ptr, P->d_dofs, P->d_matptr, op->d_rowptr,
                                                                              op->d_colval, op->d_nzval, P->d_inv, P->d_flag);
  else if (minb == 3)
    patch_gather_invert<3><<<(unsigned)P->npatch, patch::NMAX, 0, g_stream>>>(P->d_ptr, P->d_dofs, P->d_matptr, op->d_rowptr,
                                                                              op->d_colval, op->d_nzval, P->d_inv, P->d_flag);
  else
    patch_gather_invert<2><<<(unsigned)P->npatch, patch::NMAX, 0, g_stream>>>(P->d_ptr, P->d_dofs, P->d_matptr, op->d_rowptr,
                                                                              op->d_colval, op->d_nzval, P->d_inv, P->d_flag);
  prof_end(PROF_PATCH_SETUP);
  MHD_LAUNCH_CHECK();
  int nsing = 0;
  MHD_TRY(d2h(&nsing, P->d_flag, 1));
  MHD_CUDA(cudaStreamSynchronize(g_stream));
  MHD_CHECK(nsing == 0, MHD_E_INVALID, "patch smoother: %d singular patch matrices", nsing);
  return 0;
}

int patch_apply(PatchData* P, const double* d_r, double* d_z, double omega, bool accumulate) {
  if (!accumulate) MHD_CUDA(cudaMemsetAsync(d_z, 0, (size_t)P->n_uj * sizeof(double), g_stream));
  prof_begin(PROF_PATCH_APPLY);
  patch_apply_kernel<<<(unsigned)P->npatch, 256, 0, g_stream>>>(P->d_ptr, P->d_dofs, P->d_matptr, P->d_inv, d_r, d_z, omega);
  prof_end(PROF_PATCH_APPLY);
  MHD_LAUNCH_CHECK();
  return 0;
}

int64_t patch_bytes(const PatchData* P) { return P ? P->total * 8 : 0; }

}  // namespace mhd
