// Krylov building blocks on the device: CSR SpMV, dot, axpy and the fused Gram-Schmidt kernels of FGMRES.
// Stand behind mul!(y,A,x) / dot / axpy-style broadcasts that GridapSolvers' FGMRES issues on PVector /
// PSparseMatrix (configured at src/Solvers/badia2024.jl:36-40); all HBM-bandwidth bound.
#include <stdlib.h>

#include "common.h"

namespace mhd {

// Grid sizes are multiples of the SM count of the device actually in use -- queried, never assumed.
int device_sm_count() {
  static int dev = -1, sms = 0;
  if (dev != g_device || sms <= 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, g_device) != cudaSuccess || v <= 0) return 1;
    sms = v;
    dev = g_device;
  }
  return sms;
}
static int sms() {
  return device_sm_count();
}

int ensure_red(mhd_operator* op, int64_t n) {
  if (op->red_cap >= n) return 0;
  cudaStreamSynchronize(g_stream);
  cudaFree(op->d_red);
  op->d_red = nullptr;
  op->red_cap = 0;
  MHD_TRY(dev_alloc(&op->d_red, n));
  MHD_CUDA(cudaMemsetAsync(op->d_red, 0, (size_t)(n < 4096 ? n : 4096) * sizeof(double), g_stream));  // tickets of the fused reductions
  op->red_cap = n;
  return 0;
}

// ---------------------------------------------------------------- SpMV: one warp per row, rows ~200 nnz.
// Each lane streams (col,val) pairs with stride 32: fully coalesced 128 B (cols) + 256 B (vals) per step,
// x gathered through the read-only path (x is tiny next to A and stays L2 resident).
constexpr int SPMV_WARPS = 8;

// sum over the entries [lo, hi) of one row, this lane's share: U x 32 entries in flight per warp, all (col,val) loads issued
// before the dependent x gathers (unpredicated main loop + tail: predicated loads made ptxas serialise col -> x -> fma)
template <int U>
__device__ __forceinline__ double warp_row_partial(const int32_t* __restrict__ colval, const double* __restrict__ nzval,
                                                   const double* __restrict__ x, int64_t lo, int64_t hi, int lane) {
  double s[U];
#pragma unroll
  for (int u = 0; u < U; u++) s[u] = 0.0;
  int64_t p = lo + lane;
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
  return t;
}

template <int U>
__global__ void __launch_bounds__(SPMV_WARPS * 32)
spmv_warp_row(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
              const double* __restrict__ nzval, const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * SPMV_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * SPMV_WARPS;
  for (int64_t row = warp0; row < nrows; row += nwarps) {
    double t = warp_row_partial<U>(colval, nzval, x, rowptr[row], rowptr[row + 1], lane);
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

// ---------------------------------------------------------------- fused SpMV + ghost exchange over NVLink peer memory
// Replaces pack -> ncclSend/ncclRecv -> unpack -> SpMV by two launches per product:
//   * product kernel: the first `npush` CTAs store this rank's interface values straight into the neighbours' inboxes (peer-mapped
//     memory, st.global over NVLink), fence, and bump an arrival counter inside the neighbour's memory; all other CTAs run the
//     warp-per-row product on the LOCAL columns.  Columns are sorted and ghost ids come last, so that is a prefix of every row
//     and the whole exchange is overlapped with it;
//   * tail kernel (interface rows only): waits for the neighbours' counters and adds the ghost entries of those rows.
//   Inboxes are double-buffered by the parity of the product count and the counters are monotone (target = expected
//   CTAs x use count), so nothing is ever reset and a fast neighbour cannot overwrite data still being read.
// Memory model of the exchange (PTX ISA "memory consistency model"): the pusher's value stores, fence.sc.sys
// (__threadfence_system) and relaxed system-scope atomic on the counter form a release pattern; thread 0 of a tail CTA polls the
// counter with ld.relaxed.sys and, once it has seen the target, performs ONE ld.acquire.sys -- an acquire pattern that synchronises
// with the release; the CTA barrier behind it orders every other thread of the CTA after that acquire (causality order).  The
// ghost values themselves are read with ld.relaxed.sys (coherent, served by L2): the non-coherent path (__ldg / ld.global.nc) is
// only legal for data no one writes during the kernel, which is true for x and the matrix but not for the inbox.
__device__ __forceinline__ unsigned ld_relaxed_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// The exchange lives in three out-of-line functions so that the row loop of the kernel below compiles to the same code as the
// single-GPU product (inlined, their live values cost the loop its load batching: ptxas serialised col -> x -> fma, +40..100 %).
__device__ __noinline__ void halo_push_cta(const HaloDev* __restrict__ H, const double* __restrict__ x, int parity, bool dbg) {
  const long long tp0 = dbg ? clock64() : 0;
  const int k = H->push_neigh[blockIdx.x];
  const int64_t b0 = H->push_begin[blockIdx.x];
  const int64_t kend = H->send_begin[k + 1];
  const int64_t b1 = b0 + HALO_CHUNK < kend ? b0 + HALO_CHUNK : kend;
  double* dst = H->peer_inbox[parity][k];
  // each value goes straight to its final ghost slot of the neighbour (send_dst = ghost id - nrows over there)
  for (int64_t i = b0 + threadIdx.x; i < b1; i += blockDim.x) dst[H->send_dst[i]] = x[H->send_idx[i]];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd_system(H->peer_flag[parity][k], 1u);
  if (dbg && threadIdx.x == 0) atomicMax(H->err + 5, (int)(clock64() - tp0));
}

// thread 0 of a tail CTA: wait for every neighbour's counter, then acquire
__device__ __noinline__ void halo_wait(const HaloDev* __restrict__ H, const unsigned* fl, int nn, unsigned round, bool dbg) {
  const long long t0 = dbg ? clock64() : 0;
  for (int k = 0; k < nn; k++) {
    const unsigned target = H->expected[k] * round;
    unsigned spins = 0;
    while (ld_relaxed_sys_u32(fl + k) < target) {
      if (++spins > (1u << 28)) {  // never hang the GPU: flag the error (the host turns it into MHD_E_COMM) and go on
        atomicExch(H->err, 1);
        break;
      }
    }
    (void)ld_acquire_sys_u32(fl + k);  // synchronises with the pusher's fence + atomic: its values are visible from here on
  }
  if (dbg) {
    atomicMax(H->err + 1, (int)(clock64() - t0));
    atomicAdd(H->err + 2, 1);
  }
}

// Product kernel: the first `npush` CTAs push; all others are the single-GPU product restricted to the LOCAL columns of each row
// (columns are sorted and ghost ids come last, their start ghost_lo[row] is known from the symbolic phase).  No wait code in here:
// with it inlined ptxas either serialised the loads of the row loop (32 registers) or lost occupancy (40 registers, +20 %).
__global__ void __launch_bounds__(SPMV_WARPS * 32)
spmv_fused_halo(int64_t nr, const int64_t* __restrict__ rowptr, const long long* __restrict__ ghost_lo, const int32_t* __restrict__ colval,
                const double* __restrict__ nzval, const double* __restrict__ x, double* __restrict__ y,
                const HaloDev* __restrict__ H, int parity, int npush, int dbg) {
  if ((int)blockIdx.x < npush) {
    halo_push_cta(H, x, parity, dbg != 0);  // one chunk of the send list of one neighbour
    return;
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)(blockIdx.x - npush) * SPMV_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)(gridDim.x - npush) * SPMV_WARPS;
  for (int64_t row = warp0; row < nr; row += nwarps) {
    double t = warp_row_partial<2>(colval, nzval, x, rowptr[row], ghost_lo[row], lane);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) y[row] = t;
  }
}

// Tail kernel, interface rows only (a few per cent of the rows, launched right behind the product kernel): wait for the
// neighbours' counters -- normally long there, their push CTAs ran at the START of the neighbours' product kernels -- and add
// the ghost entries [ghost_lo, row end) through the coherent path (ld.relaxed.sys; values written by peer GPUs).
__global__ void __launch_bounds__(SPMV_WARPS * 32)
spmv_ghost_tail(int64_t nif, int64_t nr, const int32_t* __restrict__ if_rows, int64_t nrows, const int64_t* __restrict__ rowptr,
                const long long* __restrict__ ghost_lo, const int32_t* __restrict__ colval, const double* __restrict__ nzval,
                double* __restrict__ y, const HaloDev* __restrict__ H, const double* inbox, const unsigned* flags, int nn,
                unsigned round, int dbg) {
  if (threadIdx.x == 0) halo_wait(H, flags, nn, round, dbg != 0);
  __syncthreads();  // CTA-scope hand-over of thread 0's system-scope acquire (causality order)
  const int lane = threadIdx.x & 31;
  const double* xg = inbox - nrows;  // ghost column c lives at inbox[c - nrows]
  for (int64_t i = (int64_t)blockIdx.x * SPMV_WARPS + (threadIdx.x >> 5); i < nif; i += (int64_t)gridDim.x * SPMV_WARPS) {
    const int64_t row = if_rows[i];
    if (row >= nr) continue;  // block-row products (nr < nrows)
    double s = 0.0;
    for (int64_t p = ghost_lo[row] + lane, hi = rowptr[row + 1]; p < hi; p += 32)
      s = fma(__ldg(nzval + p), ld_relaxed_sys_f64(xg + __ldg(colval + p)), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] += s;
  }
}

// SpMV over the first nr rows only (block row of the block-ordered matrix), local x (ghosts already in place)
__global__ void __launch_bounds__(256)
spmv_rows_kernel(int64_t nr, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
                 const double* __restrict__ nzval, const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t row = warp0; row < nr; row += nwarps) {
    const int64_t lo = rowptr[row], hi = rowptr[row + 1];
    double s = 0.0;
    for (int64_t p = lo + lane; p < hi; p += 32) s = fma(nzval[p], __ldg(x + colval[p]), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
  }
}

// y[r0..r1) = (A x)[r0..r1): the rows of one diagonal block (single GPU; the block solvers of the H1-H1 preconditioner)
int spmv_row_range(mhd_operator* op, int64_t r0, int64_t r1, const double* d_x, double* d_y) {
  if (r1 <= r0) return 0;
  MHD_CHECK(g_nranks == 1 && r0 >= 0 && r1 <= op->nrows, MHD_E_INVALID, "spmv_row_range: bad range / more than one rank");
  int64_t blocks = (r1 - r0 + SPMV_WARPS - 1) / SPMV_WARPS;
  const int64_t cap = (int64_t)sms() * 32 * 4;
  if (blocks > cap) blocks = cap;
  spmv_warp_row<2><<<(unsigned)blocks, SPMV_WARPS * 32, 0, g_stream>>>(r1 - r0, op->d_rowptr + r0, op->d_colval, op->d_nzval, d_x,
                                                                       d_y + r0);
  MHD_LAUNCH_CHECK();
  return 0;
}

int spmv_with_halo(mhd_operator* op, int64_t nr, double* d_x, double* d_y) {
  if (nr <= 0) return 0;
  Halo& h = op->halo;
  if (g_nranks > 1 && h.fused && h.nneigh > 0) {
    int npush = 0;  // CTAs that push interface values (the same enumeration as in mhd_operator_halo_ipc_connect)
    for (int k = 0; k < h.nneigh; k++) npush += (int)((h.send_ptr[k + 1] - h.send_ptr[k] + HALO_CHUNK - 1) / HALO_CHUNK);
    int64_t blocks = (nr + SPMV_WARPS - 1) / SPMV_WARPS;
    const int64_t cap = (int64_t)sms() * 32 * 4;
    if (blocks > cap) blocks = cap;
    const int parity = (int)(h.epoch & 1u);
    const unsigned round = h.epoch / 2u + 1u;
    h.epoch++;
    const int dbg = getenv("MHD_HALO_DEBUG") ? 1 : 0;
    prof_begin(PROF_SPMV);
    spmv_fused_halo<<<(unsigned)(blocks + npush), SPMV_WARPS * 32, 0, g_stream>>>(nr, op->d_rowptr, h.d_row_bits, op->d_colval, op->d_nzval,
                                                                                  d_x, d_y, h.d_dev, parity, npush, dbg);
    if (h.n_if_rows > 0) {
      int64_t tb = (h.n_if_rows + SPMV_WARPS - 1) / SPMV_WARPS;
      if (tb > (int64_t)sms() * 8) tb = (int64_t)sms() * 8;
      spmv_ghost_tail<<<(unsigned)tb, SPMV_WARPS * 32, 0, g_stream>>>(h.n_if_rows, nr, h.d_if_rows, op->nrows, op->d_rowptr, h.d_row_bits,
                                                                      op->d_colval, op->d_nzval, d_y, h.d_dev, h.inbox[parity],
                                                                      h.flags + parity * HALO_MAX_NEIGH,
                                                                      getenv("MHD_FUSED_NOWAIT") ? 0 : h.nneigh, round, dbg);
    }
    prof_end(PROF_SPMV);
    MHD_LAUNCH_CHECK();
    return 0;
  }
  if (g_nranks > 1) MHD_TRY(halo_exchange(op, d_x));
  if (nr == op->nrows) return launch_spmv(op, d_x, d_y);
  int64_t blocks = (nr + 7) / 8;
  if (blocks > (int64_t)sms() * 32) blocks = (int64_t)sms() * 32;
  spmv_rows_kernel<<<(unsigned)blocks, 256, 0, g_stream>>>(nr, op->d_rowptr, op->d_colval, op->d_nzval, d_x, d_y);
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

// One-pass Gram-Schmidt reductions: out[j] = <V_j, w> for j < k and, when with_norm, out[k] = <w, w> -- ONE launch: every block
// writes its partial sums, the block that draws the last ticket adds them up in block order (deterministic).  k + with_norm <= KB.
template <int KB>
__global__ void __launch_bounds__(RED_T)
gs_dots_kernel(int64_t n, int k, int with_norm, const double* __restrict__ V, int64_t ldv, const double* __restrict__ w,
               double* __restrict__ partial, int pstride, unsigned* __restrict__ ticket, double* __restrict__ out,
               const int* __restrict__ run_if) {
  if (run_if != nullptr && *run_if == 0) return;  // conditional second Gram-Schmidt pass: every block takes the same exit
  __shared__ double sh[RED_T / 32];
  __shared__ bool last;
  double acc[KB];
#pragma unroll
  for (int j = 0; j < KB; j++) acc[j] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * RED_T + threadIdx.x; i < n; i += (int64_t)gridDim.x * RED_T) {
    const double wi = w[i];
#pragma unroll
    for (int j = 0; j < KB - 1; j++)
      if (j < k) acc[j] = fma(wi, V[(int64_t)j * ldv + i], acc[j]);
    acc[KB - 1] = fma(wi, wi, acc[KB - 1]);
  }
  const int nout = k + (with_norm ? 1 : 0);
#pragma unroll
  for (int j = 0; j < KB; j++) {
    const bool is_norm = j == KB - 1;
    if (j < k || (is_norm && with_norm)) {
      const double r = block_sum(acc[j], sh);
      if (threadIdx.x == 0) partial[(int64_t)(is_norm ? k : j) * pstride + blockIdx.x] = r;
    }
  }
  __threadfence();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  // one warp per output, all outputs at once (the serial loop over outputs with a block reduction each cost ~1.5 us per output:
  // 15 of the kernel's 27 us at k = 8); fixed summation order: lane-strided partial sums, then the shuffle tree
  for (int j = threadIdx.x >> 5; j < nout; j += RED_T / 32) {
    double s = 0.0;
    for (int i = threadIdx.x & 31; i < (int)gridDim.x; i += 32) s += __ldcg(partial + (int64_t)j * pstride + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) out[j] = s;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

constexpr int GS_KB = 18;  // m <= 16: h[0..j] and the norm in one pass

bool gs_fused_ok(int k) { return k + 1 <= GS_KB; }

int launch_gs_dots(mhd_operator* op, int64_t n, int k, bool with_norm, const double* d_V, int64_t ldv, const double* d_w, double* d_out,
                   const int* d_run_if) {
  MHD_CHECK(gs_fused_ok(k), MHD_E_INVALID, "launch_gs_dots: k = %d exceeds the fused kernel", k);
  const int nb = red_blocks(n);
  MHD_TRY(ensure_red(op, 4096 + (int64_t)GS_KB * RED_MAXB));
  double* partial = op->d_red + 4096;
  unsigned* ticket = reinterpret_cast<unsigned*>(op->d_red + 4000);  // zeroed by ensure_red, reset by the kernel
  gs_dots_kernel<GS_KB><<<nb, RED_T, 0, g_stream>>>(n, k, with_norm ? 1 : 0, d_V, ldv, d_w, partial, RED_MAXB, ticket, d_out, d_run_if);
  MHD_LAUNCH_CHECK();
  return 0;
}

// out = scale * (w - sum_j h[j] V_j): a Gram-Schmidt update fused with the normalisation of the next basis vector
// (scale_ptr == nullptr: out = w - V h, plain update; dead_ptr != nullptr && *dead_ptr: out = 0).
// Conditional re-orthogonalisation (reorth != nullptr): pass 1 -- *reorth ? plain update into `plain_out` : normalised update into
// `out`; pass 2 -- runs only if *reorth.  (w and plain_out may alias: no __restrict__ on them)
template <int KB>
__global__ void __launch_bounds__(256)
gs_update_kernel(int64_t n, int k, const double* __restrict__ V, int64_t ldv, const double* __restrict__ h, const double* w,
                 const double* __restrict__ scale_ptr, const int* __restrict__ dead_ptr, double* out, const int* __restrict__ reorth,
                 int pass, double* plain_out) {
  bool plain = false;
  if (reorth != nullptr) {
    const int r = *reorth;
    if (pass == 2 && r == 0) return;
    if (pass == 1 && r != 0) plain = true;
  }
  double hh[KB];
#pragma unroll
  for (int j = 0; j < KB; j++) hh[j] = j < k ? h[j] : 0.0;
  const bool dead = !plain && dead_ptr != nullptr && *dead_ptr != 0;
  const double sc = dead ? 0.0 : ((scale_ptr && !plain) ? *scale_ptr : 1.0);
  if (plain) out = plain_out;
  // two elements per thread and iteration: 2 (k + 1) independent loads in flight before the first fma
  for (int64_t i0 = (int64_t)blockIdx.x * 512 + threadIdx.x; i0 < n; i0 += (int64_t)gridDim.x * 512) {
    const int64_t i1 = i0 + 256;
    const bool two = i1 < n;
    double s0 = w[i0], s1 = two ? w[i1] : 0.0;
#pragma unroll
    for (int j = 0; j < KB; j++)
      if (j < k) {
        const double* Vj = V + (int64_t)j * ldv;
        const double a = Vj[i0], b = two ? Vj[i1] : 0.0;
        s0 = fma(-hh[j], a, s0);
        s1 = fma(-hh[j], b, s1);
      }
    out[i0] = dead ? 0.0 : sc * s0;
    if (two) out[i1] = dead ? 0.0 : sc * s1;
  }
}

int launch_gs_update(int64_t n, int k, const double* d_V, int64_t ldv, const double* d_h, const double* d_w, const double* d_scale,
                     const int* d_dead, double* d_out, const int* d_reorth, int pass, double* d_plain_out) {
  if (n == 0) return 0;
  MHD_CHECK(k <= GS_KB, MHD_E_INVALID, "launch_gs_update: k = %d exceeds the fused kernel", k);
  int64_t b = (n + 511) / 512;  // one pass: every thread owns two elements
  const int64_t cap = (int64_t)sms() * 32;
  if (b > cap) b = cap;
  gs_update_kernel<GS_KB><<<(unsigned)b, 256, 0, g_stream>>>(n, k, d_V, ldv, d_h, d_w, d_scale, d_dead, d_out, d_reorth, pass, d_plain_out);
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
