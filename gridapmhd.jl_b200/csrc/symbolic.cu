// Symbolic phase on the GPU: CSR sparsity + the cell-entry -> nnz scatter map, once per mesh.
//
// Stands behind Gridap's symbolic_loop_matrix!/nz_allocation/create_from_nz (reached through
// SparseMatrixAssembler at src/main.jl:222,231): every (row,col) pair of the 8 touched blocks of every cell is
// inserted (explicit zeros kept), Dirichlet rows/cols dropped, columns sorted and unique within a row.
//
// Algorithm (all data-parallel, no global sort):
//   1. row -> (cell,local row) incidence lists (count, scan, fill)
//   2. one warp per row: gather the candidate columns of its incident cells into shared memory, bitonic sort,
//      unique -> row length (pass 1) and column values (pass 2)
//   0. (operator creation) per-cell permutation that sorts the local dofs of each field by global id
//   3. one CTA per cell: binary-search each touched entry (enumerated in the Jacobian kernel's consumption order,
//      common.h) in its row -> 16-bit row-relative position;
//      count contributions per nnz and flag the nnz that receive exactly one (plain store instead of atomic).
#include "common.h"
#include "h1h1_cell.h"
#include "hdiv7_cell.h"

namespace mhd {

// ---- which local columns a local row couples to, per formulation (the kernels below are templates over this)
// H1-HDiv: the 8 touched blocks of jac_fluid_h1_hdiv (src/weakforms.jl:311): u-row: u,p,j | p-row: u | j-row: u,j,phi | phi-row: j
struct LayoutHDiv {
  static constexpr int NLOC = mhd::NLOC;
  __device__ static __forceinline__ bool coupled(int li, int lj) {
    int fi = li < OFF_P ? 0 : (li < OFF_J ? 1 : (li < OFF_F ? 2 : 3));
    int fj = lj < OFF_P ? 0 : (lj < OFF_J ? 1 : (lj < OFF_F ? 2 : 3));
    // 4 bits per row field: coupled column fields (bit fj): u:0b0111 p:0b0001 j:0b1101 phi:0b0100
    const unsigned m = 0x4D17u;
    return (m >> (fi * 4 + fj)) & 1u;
  }
};
// H1-H1: the 6 touched blocks of jac_fluid_h1_h1 (src/weakforms.jl:465): u-row: u,p,phi | p-row: u | phi-row: u,phi
struct LayoutH1H1 {
  static constexpr int NLOC = h1::NLOC;
  __device__ static __forceinline__ bool coupled(int li, int lj) {
    int fi = li < h1::OFF_P ? 0 : (li < h1::OFF_F ? 1 : 2);
    int fj = lj < h1::OFF_P ? 0 : (lj < h1::OFF_F ? 1 : 2);
    // 4 bits per row field (bit fj): u:0b111 p:0b001 phi:0b101
    const unsigned m = 0x517u;
    return (m >> (fi * 4 + fj)) & 1u;
  }
};

// ---------------------------------------------------------------- 0. per-cell permutation (once per operator)
// Inside each field the local dofs of a cell are sorted by global id (Dirichlet / absent dofs last, ties in reference
// order).  u: the 27 nodes are sorted by the id of their first component; the three components keep the node order.
__global__ void __launch_bounds__(64)
cell_permutation(int64_t ncells, const int32_t* __restrict__ gids, const int8_t* __restrict__ jsign,
                 uint8_t* __restrict__ perm, int32_t* __restrict__ pgids) {
  __shared__ int32_t key[64];
  __shared__ uint8_t pm[64];
  const int64_t cell = blockIdx.x;
  const int t = threadIdx.x;
  const int32_t* g = gids + cell * NLOC;
  // t < 27: node t (key = id of component 0); 27 <= t < 63: j dof t - 27
  int32_t k = INT32_MAX;
  if (t < 27) k = g[t];
  else if (t < 63) k = g[OFF_J + t - 27];
  if (k < 0) k = INT32_MAX;
  key[t] = k;
  __syncthreads();
  if (t < 63) {
    const int lo = t < 27 ? 0 : 27, hi = t < 27 ? 27 : 63;
    int rank = 0;
    for (int o = lo; o < hi; o++) rank += (key[o] < k) || (key[o] == k && o < t);
    uint8_t v = (uint8_t)(t - lo);
    if (t >= 27 && jsign[cell * NJ + t - 27] < 0) v |= 0x80;
    pm[lo + rank] = v;
  }
  if (t == 63) pm[63] = 0;
  __syncthreads();
  perm[cell * PERM_STRIDE + t] = pm[t];
  int32_t* pg = pgids + cell * NLOC;
  for (int i = t; i < NLOC; i += 64) {
    int src = i;
    if (i < NU) src = (i / 27) * 27 + pm[i % 27];
    else if (i >= OFF_J && i < OFF_F) src = OFF_J + (pm[27 + i - OFF_J] & 0x7F);
    pg[i] = g[src];
  }
}

int build_permutation(mhd_operator* op) {
  cudaFree(op->d_perm); op->d_perm = nullptr;
  cudaFree(op->d_pgids); op->d_pgids = nullptr;
  MHD_TRY(dev_alloc(&op->d_perm, op->ncells * PERM_STRIDE));
  MHD_TRY(dev_alloc(&op->d_pgids, op->ncells * NLOC));
  cell_permutation<<<(unsigned)op->ncells, 64, 0, g_stream>>>(op->ncells, op->d_gids, op->d_jsign, op->d_perm, op->d_pgids);
  MHD_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- exclusive scan (int32 counts -> int64 offsets)
constexpr int SCAN_B = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_B * SCAN_ITEMS;

template <class TIn>
__global__ void scan_tile_sums(const TIn* in, int64_t n, int64_t* sums) {
  __shared__ int64_t sh[SCAN_B];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int64_t s = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int64_t idx = base + (int64_t)i * SCAN_B + threadIdx.x;
    if (idx < n) s += (int64_t)in[idx];
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = SCAN_B / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = sh[0];
}

// out[i] = tile_offset + exclusive prefix inside the tile; writes out[n] = total when last tile
template <class TIn>
__global__ void scan_apply(const TIn* in, int64_t n, const int64_t* tile_off, int64_t* out) {
  __shared__ int64_t sh[SCAN_B];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int64_t v[SCAN_ITEMS];
  int64_t s = 0;
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int64_t idx = base + i;
    v[i] = idx < n ? (int64_t)in[idx] : 0;
    s += v[i];
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  // Hillis-Steele inclusive scan over the 256 thread sums
  for (int o = 1; o < SCAN_B; o <<= 1) {
    int64_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  int64_t run = tile_off[blockIdx.x] + sh[threadIdx.x] - s;
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int64_t idx = base + i;
    if (idx < n) out[idx] = run;
    run += v[i];
    if (idx == n - 1) out[n] = run;
  }
}

// exclusive scan of n counts into out[0..n] (out[n] = total). Recursive over tile sums.
template <class TIn>
static int exclusive_scan(const TIn* d_in, int64_t n, int64_t* d_out) {
  if (n == 0) {
    MHD_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int64_t), g_stream));
    return 0;
  }
  int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  int64_t *d_sums = nullptr, *d_offs = nullptr;
  MHD_TRY(dev_alloc(&d_sums, ntiles));
  MHD_TRY(dev_alloc(&d_offs, ntiles + 1));
  int rc = 0;
  if (ntiles == 1) {  // base case: a single tile starts at offset 0
    if (cudaMemsetAsync(d_offs, 0, 2 * sizeof(int64_t), g_stream) != cudaSuccess)
      rc = cuda_fail(cudaGetLastError(), "memset", __FILE__, __LINE__);
  } else {
    scan_tile_sums<TIn><<<(unsigned)ntiles, SCAN_B, 0, g_stream>>>(d_in, n, d_sums);
    g_launches++;
    if (cudaPeekAtLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "scan_tile_sums", __FILE__, __LINE__);
    if (!rc) rc = exclusive_scan<int64_t>(d_sums, ntiles, d_offs);
  }
  if (!rc) {
    scan_apply<TIn><<<(unsigned)ntiles, SCAN_B, 0, g_stream>>>(d_in, n, d_offs, d_out);
    g_launches++;
    if (cudaPeekAtLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "scan_apply", __FILE__, __LINE__);
  }
  cudaStreamSynchronize(g_stream);
  cudaFree(d_sums);
  cudaFree(d_offs);
  return rc;
}

// ---------------------------------------------------------------- 1. row incidence
__global__ void count_incidence(const int32_t* gids, int64_t nent, int64_t nrows, int32_t* cnt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nent) return;
  int32_t g = gids[i];
  if (g >= 0 && g < nrows) atomicAdd(&cnt[g], 1);
}

__global__ void fill_incidence(const int32_t* gids, int64_t nent, int64_t nrows, const int64_t* inc_ptr,
                               int32_t* cursor, int64_t* inc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nent) return;
  int32_t g = gids[i];
  if (g >= 0 && g < nrows) {
    int32_t k = atomicAdd(&cursor[g], 1);
    inc[inc_ptr[g] + k] = i;  // i = cell*129 + li
  }
}

// ---------------------------------------------------------------- 2. per-row sort/unique
constexpr int ROW_CAP = 4096;       // candidate columns per row held in shared memory
constexpr int ROW_WARPS = 2;        // warps (= rows) per CTA

template <class L, bool FILL>
__global__ void __launch_bounds__(ROW_WARPS * 32)
row_columns(const int32_t* __restrict__ gids, const int64_t* __restrict__ inc_ptr, const int64_t* __restrict__ inc,
            int64_t nrows, int32_t* __restrict__ row_len, const int64_t* __restrict__ rowptr,
            int32_t* __restrict__ colval, int* __restrict__ overflow) {
  __shared__ int32_t buf[ROW_WARPS][ROW_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * ROW_WARPS + warp;
  if (row >= nrows) return;
  int32_t* b = buf[warp];
  int n = 0;
  for (int64_t p = inc_ptr[row]; p < inc_ptr[row + 1]; p++) {
    int64_t ci = inc[p];
    int64_t cell = ci / L::NLOC;
    int li = (int)(ci % L::NLOC);
    const int32_t* g = gids + cell * L::NLOC;
    for (int base = 0; base < L::NLOC; base += 32) {
      int lj = base + lane;
      int32_t c = -1;
      if (lj < L::NLOC && L::coupled(li, lj)) c = g[lj];
      unsigned m = __ballot_sync(0xffffffffu, c >= 0);
      if (c >= 0) {
        int pos = n + __popc(m & ((1u << lane) - 1));
        if (pos < ROW_CAP) b[pos] = c;
      }
      n += __popc(m);
    }
  }
  if (n > ROW_CAP) {
    if (lane == 0) atomicExch(overflow, 1);
    return;
  }
  // pad to a power of two and bitonic-sort inside the warp
  int np2 = 32;
  while (np2 < n) np2 <<= 1;
  for (int i = n + lane; i < np2; i += 32) b[i] = INT32_MAX;
  __syncwarp();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < np2; i += 32) {
        int ixj = i ^ j;
        if (ixj > i) {
          int32_t a = b[i], c = b[ixj];
          bool up = (i & k) == 0;
          if ((a > c) == up) { b[i] = c; b[ixj] = a; }
        }
      }
      __syncwarp();
    }
  }
  // unique
  int64_t out0 = FILL ? rowptr[row] : 0;
  int count = 0;
  for (int base = 0; base < n; base += 32) {
    int i = base + lane;
    bool keep = i < n && (i == 0 || b[i] != b[i - 1]);
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (FILL && keep) colval[out0 + count + __popc(m & ((1u << lane) - 1))] = b[i];
    count += __popc(m);
  }
  if (!FILL && lane == 0) row_len[row] = count;
}

// ---------------------------------------------------------------- 3. scatter map
template <class L>
__global__ void __launch_bounds__(256)
build_map(const int32_t* __restrict__ gids /* permuted */, const uint16_t* __restrict__ order, const int nent_cell,
          const int nent_pad, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
          int64_t nrows, uint16_t* __restrict__ map, uint8_t* __restrict__ contrib) {
  __shared__ int32_t g[L::NLOC];
  int64_t cell = blockIdx.x;
  for (int i = threadIdx.x; i < L::NLOC; i += blockDim.x) g[i] = gids[cell * L::NLOC + i];
  __syncthreads();
  uint16_t* m = map + cell * nent_pad;
  for (int e = threadIdx.x; e < nent_pad; e += blockDim.x) {
    uint16_t code = MAP_SKIP;
    if (e < nent_cell) {
      const int li = order[e] >> 8, lj = order[e] & 0xFF;
      int32_t r = -1, c = -1;
      if (order[e] != ORDER_PAD) { r = g[li]; c = g[lj]; }
      if (r >= 0 && r < nrows && c >= 0) {
        int64_t lo = rowptr[r], hi = rowptr[r + 1] - 1, base = lo;
        while (lo < hi) {
          int64_t mid = (lo + hi) >> 1;
          if (colval[mid] < c) lo = mid + 1; else hi = mid;
        }
        code = (uint16_t)(lo - base);
        // saturating contribution counter (1 byte per nnz)
        uint32_t* word = (uint32_t*)(contrib + (lo & ~(int64_t)3));
        unsigned sh = (unsigned)(lo & 3) * 8;
        uint32_t old = *word;
        if (((old >> sh) & 0xFF) < 2) {
          // CAS loop: increment the byte, saturate at 2
          uint32_t assumed;
          do {
            assumed = old;
            uint32_t bval = (assumed >> sh) & 0xFF;
            if (bval >= 2) break;
            old = atomicCAS(word, assumed, assumed + (1u << sh));
          } while (old != assumed);
        }
      }
    }
    m[e] = code;
  }
}

template <class L>
__global__ void __launch_bounds__(256)
flag_exclusive(const int32_t* __restrict__ gids /* permuted */, const uint16_t* __restrict__ order, const int nent_cell,
               const int nent_pad, const int64_t* __restrict__ rowptr, const uint8_t* __restrict__ contrib,
               uint16_t* __restrict__ map, unsigned long long* __restrict__ stats) {
  __shared__ int32_t g[L::NLOC];
  int64_t cell = blockIdx.x;
  for (int i = threadIdx.x; i < L::NLOC; i += blockDim.x) g[i] = gids[cell * L::NLOC + i];
  __syncthreads();
  uint16_t* m = map + cell * nent_pad;
  unsigned nent = 0, nex = 0;
  for (int e = threadIdx.x; e < nent_cell; e += blockDim.x) {
    uint16_t code = m[e];
    if (code == MAP_SKIP) continue;
    const int li = order[e] >> 8;
    int64_t idx = rowptr[g[li]] + code;
    nent++;
    if (contrib[idx] == 1) {
      m[e] = code | MAP_EXCL;
      nex++;
    }
  }
  // block reduce the two counters
  for (int o = 16; o > 0; o >>= 1) {
    nent += __shfl_down_sync(0xffffffffu, nent, o);
    nex += __shfl_down_sync(0xffffffffu, nex, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&stats[0], (unsigned long long)nent);
    atomicAdd(&stats[1], (unsigned long long)nex);
  }
}

__global__ void cell_row_starts(int64_t n, int64_t nrows, const int32_t* __restrict__ pgids,
                                const int64_t* __restrict__ rowptr, int64_t* __restrict__ rowstart) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t g = pgids[i];
  rowstart[i] = (g >= 0 && g < nrows) ? rowptr[g] : -1;
}

__global__ void max_row_len(const int32_t* row_len, int64_t nrows, int* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int v = i < nrows ? row_len[i] : 0;
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_down_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

// map_gids: the cell -> local id table in the numbering the entry enumeration `order` refers to
template <class L>
static int symbolic_build_impl(mhd_operator* op, const int32_t* map_gids, const std::vector<uint16_t>& order, int nent_cell,
                               int nent_pad) {
  const int64_t nent = op->ncells * L::NLOC;
  const int64_t nrows = op->nrows;
  int32_t *d_cnt = nullptr, *d_cursor = nullptr, *d_rowlen = nullptr;
  int64_t *d_incptr = nullptr, *d_inc = nullptr;
  int* d_flags = nullptr;  // [0] overflow, [1] max row length
  uint8_t* d_contrib = nullptr;
  unsigned long long* d_stats = nullptr;
  int rc = 0;
  auto cleanup = [&]() {
    cudaStreamSynchronize(g_stream);
    cudaFree(d_cnt); cudaFree(d_cursor); cudaFree(d_rowlen); cudaFree(d_incptr); cudaFree(d_inc);
    cudaFree(d_flags); cudaFree(d_contrib); cudaFree(d_stats);
  };
#define SY(x) do { if (!rc) rc = (x); } while (0)
#define SYC(x) do { if (!rc) { cudaError_t _e = (x); if (_e != cudaSuccess) rc = cuda_fail(_e, #x, __FILE__, __LINE__); } } while (0)
#define SYL() do { if (!rc) { g_launches++; cudaError_t _e = cudaPeekAtLastError(); if (_e != cudaSuccess) rc = cuda_fail(_e, "launch", __FILE__, __LINE__); } } while (0)
  SY(dev_alloc(&d_cnt, nrows));
  SY(dev_alloc(&d_cursor, nrows));
  SY(dev_alloc(&d_rowlen, nrows));
  SY(dev_alloc(&d_incptr, nrows + 1));
  SY(dev_alloc(&d_flags, 2));
  SY(dev_alloc(&d_stats, 2));
  SYC(cudaMemsetAsync(d_cnt, 0, nrows * sizeof(int32_t), g_stream));
  SYC(cudaMemsetAsync(d_cursor, 0, nrows * sizeof(int32_t), g_stream));
  SYC(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), g_stream));
  SYC(cudaMemsetAsync(d_stats, 0, 2 * sizeof(unsigned long long), g_stream));
  const unsigned gb = (unsigned)((nent + 255) / 256);
  if (!rc) { count_incidence<<<gb, 256, 0, g_stream>>>(op->d_gids, nent, nrows, d_cnt); SYL(); }
  SY(exclusive_scan<int32_t>(d_cnt, nrows, d_incptr));
  int64_t ninc = 0;
  SY(d2h(&ninc, d_incptr + nrows, 1));
  SYC(cudaStreamSynchronize(g_stream));
  SY(dev_alloc(&d_inc, ninc));
  if (!rc) { fill_incidence<<<gb, 256, 0, g_stream>>>(op->d_gids, nent, nrows, d_incptr, d_cursor, d_inc); SYL(); }
  const unsigned rb = (unsigned)((nrows + ROW_WARPS - 1) / ROW_WARPS);
  if (!rc) {
    row_columns<L, false><<<rb, ROW_WARPS * 32, 0, g_stream>>>(op->d_gids, d_incptr, d_inc, nrows, d_rowlen, nullptr, nullptr, d_flags);
    SYL();
  }
  if (!rc) { max_row_len<<<(unsigned)((nrows + 255) / 256), 256, 0, g_stream>>>(d_rowlen, nrows, d_flags + 1); SYL(); }
  int flags[2] = {0, 0};
  SY(d2h(flags, d_flags, 2));
  SYC(cudaStreamSynchronize(g_stream));
  if (!rc && flags[0]) {
    set_error("symbolic: a row couples to more than %d candidate columns (vertex valence too high)", ROW_CAP);
    rc = MHD_E_CAPACITY;
  }
  if (!rc && flags[1] > MAX_ROW_NNZ) {
    set_error("symbolic: row length %d exceeds the 15-bit scatter-map limit %d", flags[1], MAX_ROW_NNZ);
    rc = MHD_E_CAPACITY;
  }
  cudaFree(op->d_rowptr); op->d_rowptr = nullptr;
  SY(dev_alloc(&op->d_rowptr, nrows + 1));
  SY(exclusive_scan<int32_t>(d_rowlen, nrows, op->d_rowptr));
  int64_t nnz = 0;
  SY(d2h(&nnz, op->d_rowptr + nrows, 1));
  SYC(cudaStreamSynchronize(g_stream));
  cudaFree(op->d_colval); op->d_colval = nullptr;
  cudaFree(op->d_nzval); op->d_nzval = nullptr;
  cudaFree(op->d_map); op->d_map = nullptr;
  SY(dev_alloc(&op->d_colval, nnz));
  SY(dev_alloc(&op->d_nzval, nnz));
  SY(dev_alloc(&op->d_map, op->ncells * nent_pad));
  if (!rc && !op->d_order) {
    SY(dev_alloc(&op->d_order, nent_cell));
    SY(h2d(op->d_order, order.data(), nent_cell));
    SYC(cudaStreamSynchronize(g_stream));
  }
  SY(dev_alloc(&d_contrib, (nnz + 3) / 4 * 4 + 4));
  SYC(cudaMemsetAsync(op->d_nzval, 0, (size_t)(nnz > 0 ? nnz : 1) * sizeof(double), g_stream));
  SYC(cudaMemsetAsync(d_contrib, 0, (size_t)((nnz + 3) / 4 * 4 + 4), g_stream));
  if (!rc) {
    row_columns<L, true><<<rb, ROW_WARPS * 32, 0, g_stream>>>(op->d_gids, d_incptr, d_inc, nrows, nullptr, op->d_rowptr, op->d_colval, d_flags);
    SYL();
  }
  cudaFree(op->d_rowstart); op->d_rowstart = nullptr;
  SY(dev_alloc(&op->d_rowstart, nent));
  if (!rc) { cell_row_starts<<<gb, 256, 0, g_stream>>>(nent, nrows, map_gids, op->d_rowptr, op->d_rowstart); SYL(); }
  if (!rc) { build_map<L><<<(unsigned)op->ncells, 256, 0, g_stream>>>(map_gids, op->d_order, nent_cell, nent_pad, op->d_rowptr, op->d_colval, nrows, op->d_map, d_contrib); SYL(); }
  if (!rc) { flag_exclusive<L><<<(unsigned)op->ncells, 256, 0, g_stream>>>(map_gids, op->d_order, nent_cell, nent_pad, op->d_rowptr, d_contrib, op->d_map, d_stats); SYL(); }
  unsigned long long stats[2] = {0, 0};
  SY(d2h(stats, d_stats, 2));
  if (!rc && op->formulation == FORM_HDIV && op->jac_version == 7) {
    op->nnz = nnz;
    SY(v7_build_shared_mask(op, d_contrib));
  }
  SYC(cudaStreamSynchronize(g_stream));
  cleanup();
  if (rc) return rc;
  op->nnz = nnz;
  op->nentries = (int64_t)stats[0];
  op->nexclusive = (int64_t)stats[1];
  op->has_symbolic = true;
  return 0;
#undef SY
#undef SYC
#undef SYL
}

int symbolic_build(mhd_operator* op) {
  if (op->formulation == FORM_H1H1) {
    std::vector<uint16_t> ord(h1::NENT);
    for (int e = 0; e < h1::NENT; e++) {
      int li, lj;
      h1::entry_rowcol(e, &li, &lj);
      ord[e] = (uint16_t)(li << 8 | lj);
    }
    return symbolic_build_impl<LayoutH1H1>(op, op->d_gids, ord, h1::NENT, h1::NENT_PAD);
  }
  std::vector<uint16_t> ord;
  if (op->jac_version == 7) {  // sum-factorised kernel: its own enumeration in the permuted local numbering
    v7_entry_order(ord);
    return symbolic_build_impl<LayoutHDiv>(op, op->d_pgids, ord, h7::NENT, h7::NENT);
  }
  entry_order(ord);
  return symbolic_build_impl<LayoutHDiv>(op, op->d_pgids, ord, NENT, NENT_PAD);
}

}  // namespace mhd
