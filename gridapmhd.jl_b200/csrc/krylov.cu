// Krylov building blocks on the device: CSR SpMV, dot, axpy and the fused Gram-Schmidt kernels of FGMRES.
// Stand behind mul!(y,A,x) / dot / axpy-style broadcasts that GridapSolvers' FGMRES issues on PVector /
// PSparseMatrix (configured at src/Solvers/badia2024.jl:36-40); all HBM-bandwidth bound.
#include <stdlib.h>

#include "common.h"

namespace mhd {

static int g_sms = 0;
static int sms() {
  if (!g_sms) cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, g_device);
  return g_sms ? g_sms : 148;
}

int ensure_red(mhd_operator* op, int64_t n) {
  if (op->red_cap >= n) return 0;
  cudaStreamSynchronize(g_stream);
  cudaFree(op->d_red);
  op->d_red = nullptr;
  op->red_cap = 0;
  MHD_TRY(dev_alloc(&op->d_red, n));
  op->red_cap = n;
  return 0;
}

// ---------------------------------------------------------------- SpMV: one warp per row, rows ~200 nnz.
// Each lane streams (col,val) pairs with stride 32: fully coalesced 128 B (cols) + 256 B (vals) per step,
// x gathered through the read-only path (x is tiny next to A and stays L2 resident).
constexpr int SPMV_WARPS = 8;

template <int U>
__global__ void __launch_bounds__(SPMV_WARPS * 32)
spmv_warp_row(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
              const double* __restrict__ nzval, const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * SPMV_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * SPMV_WARPS;
  for (int64_t row = warp0; row < nrows; row += nwarps) {
    const int64_t lo = rowptr[row], hi = rowptr[row + 1];
    double s[U];
#pragma unroll
    for (int u = 0; u < U; u++) s[u] = 0.0;
    int64_t p = lo + lane;
    // U x 32 entries in flight per warp: all (col,val) loads are issued before the dependent x gathers
    for (; p + 32 * (U - 1) < hi; p += 32 * U) {
      int32_t c[U];
      double v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        c[u] = __ldg(colval + p + 32 * u);
        v[u] = __ldg(nzval + p + 32 * u);
      }
#pragma unroll
      for (int u = 0; u < U; u++) s[u] = fma(v[u], __ldg(x + c[u]), s[u]);
    }
    for (; p < hi; p += 32) s[0] = fma(__ldg(nzval + p), __ldg(x + __ldg(colval + p)), s[0]);
    double t = s[0];
#pragma unroll
    for (int u = 1; u < U; u++) t += s[u];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) y[row] = t;
  }
}

int launch_spmv(mhd_operator* op, const double* d_x, double* d_y) {
  if (op->nrows == 0) return 0;
  static int unroll = 0, ctas_per_sm = 0;
  if (!unroll) {
    const char* e = getenv("MHD_SPMV_UNROLL");
    unroll = e ? atoi(e) : 2;      // tuned on B200/cfg2: rows of ~200 nnz, 2 x 32 entries in flight per warp
    const char* g = getenv("MHD_SPMV_CTAS_PER_SM");
    ctas_per_sm = g ? atoi(g) : 32; // grid = SMs x 32 x 4 CTAs: fine-grained grid-stride evens out row lengths
  }
  int64_t blocks = (op->nrows + SPMV_WARPS - 1) / SPMV_WARPS;
  const int64_t cap = (int64_t)sms() * ctas_per_sm * 4;
  if (blocks > cap) blocks = cap;
  prof_begin(PROF_SPMV);
  if (unroll == 1)
    spmv_warp_row<1><<<(unsigned)blocks, SPMV_WARPS * 32, 0, g_stream>>>(op->nrows, op->d_rowptr, op->d_colval, op->d_nzval, d_x, d_y);
  else if (unroll == 2)
    spmv_warp_row<2><<<(unsigned)blocks, SPMV_WARPS * 32, 0, g_stream>>>(op->nrows, op->d_rowptr, op->d_colval, op->d_nzval, d_x, d_y);
  else if (unroll == 8)
    spmv_warp_row<8><<<(unsigned)blocks, SPMV_WARPS * 32, 0, g_stream>>>(op->nrows, op->d_rowptr, op->d_colval, op->d_nzval, d_x, d_y);
  else
    spmv_warp_row<4><<<(unsigned)blocks, SPMV_WARPS * 32, 0, g_stream>>>(op->nrows, op->d_rowptr, op->d_colval, op->d_nzval, d_x, d_y);
  prof_end(PROF_SPMV);
  MHD_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- reductions (deterministic two-stage)
constexpr int RED_T = 256;
constexpr int RED_MAXB = 1024;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// partial[k*gridDim.x + block] = sum over the block's grid-stride slice of w[i] * V[k*ldv + i]
template <int KB>
__global__ void __launch_bounds__(RED_T)
multi_dot_partial(int64_t n, int k0, int kcount, const double* __restrict__ V, int64_t ldv,
                  const double* __restrict__ w, double* __restrict__ partial, int pstride) {
  __shared__ double sh[RED_T / 32];
  double acc[KB];
#pragma unroll
  for (int j = 0; j < KB; j++) acc[j] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * RED_T + threadIdx.x; i < n; i += (int64_t)gridDim.x * RED_T) {
    const double wi = w[i];
#pragma unroll
    for (int j = 0; j < KB; j++)
      if (j < kcount) acc[j] = fma(wi, V[(int64_t)(k0 + j) * ldv + i], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < KB; j++) {
    if (j < kcount) {
      const double r = block_sum(acc[j], sh);
      if (threadIdx.x == 0) partial[(int64_t)(k0 + j) * pstride + blockIdx.x] = r;
    }
  }
}

__global__ void __launch_bounds__(RED_T)
reduce_partials(int nblocks, int pstride, const double* __restrict__ partial, double* __restrict__ out) {
  __shared__ double sh[RED_T / 32];
  const int k = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += RED_T) s += partial[(int64_t)k * pstride + i];
  const double r = block_sum(s, sh);
  if (threadIdx.x == 0) out[k] = r;
}

static int red_blocks(int64_t n) {
  int64_t b = (n + RED_T * 4 - 1) / (RED_T * 4);
  const int64_t cap = sms() * 4 < RED_MAXB ? sms() * 4 : RED_MAXB;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int launch_multi_dot(mhd_operator* op, int64_t n, int k, const double* d_V, int64_t ldv, const double* d_w, double* d_h) {
  const int nb = red_blocks(n);
  MHD_TRY(ensure_red(op, 4096 + (int64_t)k * RED_MAXB));
  double* partial = op->d_red + 4096;
  for (int k0 = 0; k0 < k; k0 += 8) {
    const int kc = k - k0 < 8 ? k - k0 : 8;
    multi_dot_partial<8><<<nb, RED_T, 0, g_stream>>>(n, k0, kc, d_V, ldv, d_w, partial, RED_MAXB);
    MHD_LAUNCH_CHECK();
  }
  reduce_partials<<<k, RED_T, 0, g_stream>>>(nb, RED_MAXB, partial, d_h);
  MHD_LAUNCH_CHECK();
  return 0;
}

int launch_dot(mhd_operator* op, int64_t n, const double* d_x, const double* d_y, double* d_out) {
  return launch_multi_dot(op, n, 1, d_x, n, d_y, d_out);
}

// w += sign * sum_j h[j] V_j   (h read from device memory: no host round trip inside the Krylov loop)
template <int KB>
__global__ void __launch_bounds__(256)
multi_axpy_kernel(int64_t n, int k0, int kcount, const double* __restrict__ V, int64_t ldv, const double* __restrict__ h,
                  double sign, double* __restrict__ w) {
  double hh[KB];
#pragma unroll
  for (int j = 0; j < KB; j++) hh[j] = j < kcount ? sign * h[k0 + j] : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    double s = w[i];
#pragma unroll
    for (int j = 0; j < KB; j++)
      if (j < kcount) s = fma(hh[j], V[(int64_t)(k0 + j) * ldv + i], s);
    w[i] = s;
  }
}

int launch_multi_axpy(int64_t n, int k, const double* d_V, int64_t ldv, const double* d_h, double sign, double* d_w) {
  if (n == 0) return 0;
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sms() * 8;
  if (b > cap) b = cap;
  for (int k0 = 0; k0 < k; k0 += 8) {
    const int kc = k - k0 < 8 ? k - k0 : 8;
    multi_axpy_kernel<8><<<(unsigned)b, 256, 0, g_stream>>>(n, k0, kc, d_V, ldv, d_h, sign, d_w);
    MHD_LAUNCH_CHECK();
  }
  return 0;
}

__global__ void __launch_bounds__(256) axpy_kernel(int64_t n, double a, const double* __restrict__ x, double* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) y[i] = fma(a, x[i], y[i]);
}

int launch_axpy(int64_t n, double a, const double* d_x, double* d_y) {
  if (n == 0) return 0;
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sms() * 8;
  if (b > cap) b = cap;
  axpy_kernel<<<(unsigned)b, 256, 0, g_stream>>>(n, a, d_x, d_y);
  MHD_LAUNCH_CHECK();
  return 0;
}

}  // namespace mhd
