// "v7" kernel of the H1-HDiv formulation (jac_fluid_h1_hdiv / jac_solid_h1_hdiv, res_fluid_h1_hdiv / res_solid_h1_hdiv,
// src/weakforms.jl:255-338): the fully sum-factorised cell code of hdiv7_cell.h run by one CTA of 256 threads per cell on a
// persistent grid (2 CTAs per SM).  ~0.13 M FMA per cell instead of 0.43 M, operands of a few hundred bytes instead of 27 x 27
// panels.  MODE 0 = jacobian!, 1 = residual_and_jacobian! (the residual phases ride in the barrier intervals of the Jacobian
// phases), 2 = residual! alone.
//
// * The 1-D factors are DISCOVERED from the plain tables of mhd_tables_t when the operator is created (hdiv7_tables.h); an
//   operator whose tables lack the tensor structure keeps the generic tensor-core kernel of assembly.cu.
// * Values are staged in shared memory in destination order (permuted local numbering: every field sorted by global id),
//   chunk by chunk in two buffers; while chunk n+1 is computed, chunk n is swept out 32 consecutive map entries per
//   instruction.  The u16 map codes of a chunk arrive in shared memory by cp.async one barrier interval before they are needed,
//   the ids / row starts of the next cell likewise, its state values one interval early in a register.
// * nnz that receive exactly one contribution (67 % on Hunt meshes) are stored plainly; the others are accumulated with
//   RED.ADD.F64 after `zero_shared_kernel` cleared just the 32-byte sectors that hold them (a bit mask from the symbolic phase)
//   -- no 1.2 GB memset.
// * Cells are walked breadth-first over face neighbours (v7_build_cell_order) so that cells sharing sectors meet in L2:
//   DRAM traffic 1.12x the algorithmic bytes.  Deterministic mode: one launch per colour of a cell colouring.
// What bounds it (ncu, profiles/r2_*): the shared-memory data pipe (64 %) and instruction issue (46 %), not HBM -- DESIGN.md 4.2b, 7.
#include <stdlib.h>

#include "common.h"
#include "hdiv7_cell.h"

namespace mhd {

namespace {

#ifndef MHD_V7_NT
#define MHD_V7_NT 256
#endif
constexpr int V7_NT = MHD_V7_NT;   // threads per CTA (one cell per CTA); a multiple of 32
constexpr int V7_NW = V7_NT / 32;  // warps

struct V7Args {
  const double* coords;
  const int32_t* cell_nodes;
  const int32_t* pgids;
  const long long* rowstart;
  const uint8_t* perm;
  const uint8_t* cell_solid;
  const double* cell_sigma;
  const double* dir;
  const h7::Tab7* tab;
  const int32_t* cell_list;  // nullable: cells of one colour (deterministic mode)
  unsigned long long* clk;   // nullable: phase clocks
};

// ids of the NEXT cell, fetched with cp.async while the current cell is being integrated (phase_load then only gathers x)
struct NextIds {
  uint8_t perm[64];   // 16-byte copies
  int32_t nodes[8];   // 16-byte copies
  long long rowstart[h7::NLOC];
  int32_t gid[h7::NLOC];
  int32_t pad_[1];
};
static_assert(sizeof(NextIds) % 16 == 0, "NextIds must keep 16-byte alignment");

constexpr int V7_CODES = h7::CH_UJ_PAD;  // u16 codes of the largest chunk
constexpr size_t V7_SMEM_CELL = (sizeof(h7::Cell7) + 15) / 16 * 16;
constexpr size_t V7_SMEM_SMALL = (sizeof(h7::SmallDyn) + 15) / 16 * 16;
constexpr size_t V7_OFF_NEXT = V7_SMEM_CELL + V7_SMEM_SMALL;
constexpr size_t V7_OFF_CODES = V7_OFF_NEXT + sizeof(NextIds);
constexpr size_t V7_SMEM = V7_OFF_CODES + V7_CODES * sizeof(uint16_t);
static_assert(2 * (V7_SMEM + 64 + 1024) <= 233472, "two CTAs of the v7 kernel must fit one SM");
static_assert(V7_NT % 32 == 0 && V7_NT >= 256 && V7_NT <= 512, "LOAD_ITEMS, the 243-item chunks and fetch_next assume 256..512 threads");

// The cell-independent 1-D tables exist twice: the ones looked up with thread-dependent indices in shared memory (SmallDyn C),
// everything in constant memory (Small7 K): a table operand whose index is known at compile time (the unrolled inner loops of
// the hot phases) becomes an immediate constant-bank operand of its DFMA -- no shared-memory wavefront, which is what bounds
// the kernel.  (Measured: routing the thread-dependent lookups through the constant cache as well -- LDC with divergent
// addresses -- is slower than shared memory.)  One copy per module: v7_launch_jacobian re-uploads it when an operator with
// different tables comes along.
__constant__ h7::Small7 c_small7;
std::vector<unsigned char> g_small7_loaded;  // host copy of what c_small7 holds (compared by content: 5.6 KB)
int g_small7_device = -1;

extern __shared__ __align__(16) unsigned char v7_smem[];
__device__ __forceinline__ h7::Cell7& sm_cell() { return *reinterpret_cast<h7::Cell7*>(v7_smem); }
__device__ __forceinline__ h7::SmallDyn& sm_small() { return *reinterpret_cast<h7::SmallDyn*>(v7_smem + V7_SMEM_CELL); }
__device__ __forceinline__ NextIds& sm_next() { return *reinterpret_cast<NextIds*>(v7_smem + V7_OFF_NEXT); }
__device__ __forceinline__ uint16_t* sm_codes() { return reinterpret_cast<uint16_t*>(v7_smem + V7_OFF_CODES); }
template <int B>
__device__ __forceinline__ double* sm_buf() { return B == 0 ? sm_cell().r1 : sm_cell().r3; }

// Each heavy phase is its own (non-inlined) function: one register allocation per phase, nothing hoisted across phases --
// inlined into one body ptxas wants 164 registers and spills at the 128 the two resident CTAs allow.  The functions re-derive
// their shared-memory references from the extern array (a reference PARAMETER would turn every access into a generic LD/ST).
#define V7_NI static __device__ __noinline__
template <int CONV, bool ZJ>
V7_NI void ni_fields(int tid, const h7::Params& P) { h7::phase_fields<CONV, ZJ>(sm_cell(), sm_small(), tid, V7_NT, P); }
template <bool ZJ>
V7_NI void ni_stage1_item(int item) { h7::stage1_item<ZJ>(sm_cell(), sm_small(), c_small7, item); }
// Which 32-item group ("warp item") a warp takes in each round.  The item classes of the stage contractions differ in cost by 4x
// (a Newton-base item sums up to four field groups, a u-p item has three outputs) and sit in index order, so with item = tid +
// 256 r the heavy classes pile up on warps 0-2 (max / mean = 1.28 in stage 2, 1.26 in stage 1: barrier stalls).  The tables
// deal the groups to the 8 warps by estimated cost, longest first; slot [w + 8 r] = group of warp w in round r.
__constant__ unsigned char c_stage1_group[24] = {5, 8, 11, 10, 12, 6, 7, 0, 17, 14, 13, 9, 4, 2, 3, 1, 23, 21, 15, 18, 22, 19, 20, 16};
__constant__ unsigned char c_stage2_group[32] = {10, 7, 8, 9, 16, 20, 18, 21, 19, 5, 15, 0, 14, 13, 12, 11,
                                                 27, 28, 17, 6, 4, 3, 2, 1, 30, 31, 26, 29, 25, 24, 23, 22};
template <int ROUNDS>
__device__ __forceinline__ int dealt_item(const unsigned char* group, int tid, int r) {
  if (V7_NT != 256) return tid + r * V7_NT;  // the tables are made for 8 warps
  return (int)group[(tid >> 5) + 8 * r] * 32 + (tid & 31);
}
template <bool ZJ>
__device__ __forceinline__ void ni_stage1(int tid) {  // 729 items on 256 threads: three calls, no loop (see stage1_item)
  constexpr int R = (h7::STAGE1_ITEMS + V7_NT - 1) / V7_NT;
  static_assert(V7_NT != 256 || R * 8 == 24, "c_stage1_group");
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int item = dealt_item<R>(c_stage1_group, tid, r);
    if (item < h7::STAGE1_ITEMS) ni_stage1_item<ZJ>(item);
  }
}
template <int CONV, bool ZJ>
V7_NI void ni_stage2_item(int item) { h7::stage2_item<CONV, ZJ>(sm_cell(), sm_small(), c_small7, item); }
template <int CONV, bool ZJ>
__device__ __forceinline__ void ni_stage2(int tid) {  // 1 017 items: four calls
  constexpr int R = (h7::STAGE2_ITEMS + V7_NT - 1) / V7_NT;
  static_assert(V7_NT != 256 || R * 8 == 32, "c_stage2_group");
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int item = dealt_item<R>(c_stage2_group, tid, r);
    if (item < h7::STAGE2_ITEMS) ni_stage2_item<CONV, ZJ>(item);
  }
}
// (the component / direction of a chunk is a run-time argument: ONE copy of the code -- the loop body of the kernel is several
// times the 32 KB instruction cache and every duplicated phase shows up as no_inst stalls)
template <int CONV, bool ZU>
V7_NI void ni_chunk_uu(int tid, int cc) {
  if (tid < 243) h7::chunk_uu_item<CONV, ZU>(sm_cell(), sm_small(), c_small7, tid, cc, (cc & 1) ? sm_cell().r3 : sm_cell().r1);
}
V7_NI void ni_chunk_uj_item(int item, const h7::Params& P, bool ju) {
  h7::chunk_uj_item(sm_cell(), c_small7, P, ju ? sm_cell().r1 : sm_cell().r3, item, ju);
}
__device__ __forceinline__ void ni_chunk_uj(int tid, const h7::Params& P, bool ju) {  // 324 items on 256 threads
#pragma unroll
  for (int r = 0; r < (324 + V7_NT - 1) / V7_NT; r++)
    if (tid + r * V7_NT < 324) ni_chunk_uj_item(tid + r * V7_NT, P, ju);
}
template <bool ZJ>
V7_NI void ni_chunk_rest(int tid) { h7::chunk_rest<ZJ>(sm_cell(), sm_small(), tid, V7_NT, sm_buf<1>()); }

#define V7_CLK(slot)                                                        \
  do {                                                                      \
    if (A.clk != nullptr && tid == 0) {                                     \
      const long long now_ = clock64();                                     \
      clk_acc[slot] += (unsigned long long)(now_ - (long long)clk_acc[7]);  \
      clk_acc[7] = (unsigned long long)now_;                                \
    }                                                                       \
  } while (0)

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
// ids of `cell` -> staging (one or two copies per thread), and its scatter map -> L2
__device__ __forceinline__ void fetch_next(NextIds& N, const V7Args& A, const uint16_t* __restrict__ map, int64_t cell, int tid, bool with_map = true) {
  for (int i = tid; i < h7::NLOC; i += V7_NT) {
    if (with_map) cp_async8(&N.rowstart[i], A.rowstart + cell * h7::NLOC + i);  // (residual! works before the symbolic phase)
    cp_async4(&N.gid[i], A.pgids + cell * h7::NLOC + i);
  }
  if (tid >= 192 && tid < 196) cp_async16(&N.perm[(tid - 192) * 16], A.perm + cell * PERM_STRIDE + (tid - 192) * 16);
  if (tid >= 224 && tid < 226) cp_async16(&N.nodes[(tid - 224) * 4], A.cell_nodes + cell * 8 + (tid - 224) * 4);
  const char* mp = reinterpret_cast<const char*>(map + cell * h7::NENT);
  if (with_map && tid * 128 < h7::NENT * 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(mp + tid * 128));
}

// shared-memory loads through 32-bit shared addresses (no generic-address arithmetic in the sweep loops)
// a map code, sign-extended: >= 0 RED at position code | -1 (MAP_SKIP) nothing | < -1 plain store at position code & 0x7FFF
__device__ __forceinline__ int lds_code(unsigned a) {
  int v;
  asm volatile("{\n\t.reg .s16 h;\n\tld.shared.s16 h, [%1];\n\tcvt.s32.s16 %0, h;\n\t}" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long lds_u64(unsigned a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// value v of an entry goes to rowbase + 8 * code: plain store (MAP_EXCL set, not MAP_SKIP) or reduction, predicated -- no branch
__device__ __forceinline__ void scatter_one(int code, unsigned long long rowbase, double v) {
  asm volatile(
      "{\n\t"
      ".reg .pred pst, prd;\n\t"
      ".reg .u32 c;\n\t"
      ".reg .u64 a;\n\t"
      "and.b32 c, %2, 0x7FFF;\n\t"
      "mad.wide.u32 a, c, 8, %0;\n\t"
      "setp.lt.s32 pst, %2, -1;\n\t"   // 0x8000 <= code < 0xFFFF as u16
      "setp.ge.s32 prd, %2, 0;\n\t"
      "@pst st.global.f64 [a], %1;\n\t"
      "@prd red.global.add.f64 [a], %1;\n\t"
      "}" ::"l"(rowbase), "d"(v), "r"(code)
      : "memory");
}

// One call per barrier interval: (1) sweep the chunk staged in the previous interval -- nseg 32-entry segments, ncol entries per
// row, rows from row0 (pads carry MAP_SKIP; ncol_recip = ceil(2^20 / ncol): exact quotient for i < 2^15) -- and (2) start the
// cp.async copy of the NEXT chunk's codes (consumed one interval later).  Warp w owns the segments w, w + 8, ...: it alone
// reads and refills them, so one code buffer serves all chunks (__syncwarp between the two halves is the only ordering needed).
// (bufsel: 0 = Cell7::r1, 1 = Cell7::r3 -- an index, not a pointer: a pointer PARAMETER makes every staged value a generic LD)
V7_NI void sweep_and_fetch(int bufsel, int nseg, int row0, unsigned ncol_recip, const uint16_t* __restrict__ next, int next_nseg, int tid) {
  const int w = tid >> 5, lane = tid & 31;
  const unsigned i0 = (unsigned)(w * 32 + lane), n = (unsigned)nseg * 32u;
  const unsigned sa_codes0 = smem_u32(sm_codes());
  unsigned sa_c = sa_codes0 + i0 * 2u;
  unsigned sa_b = smem_u32(bufsel ? sm_cell().r3 : sm_cell().r1) + i0 * 8u;
  const unsigned sa_r = smem_u32(sm_cell().rowaddr + row0);
  unsigned i = i0, q = i0 * ncol_recip;  // q >> 20 = row of entry i
  constexpr unsigned NT = (unsigned)V7_NT;
  const unsigned dq = NT * ncol_recip;
  // four entries per lane and iteration: all loads first, then the four predicated store / reduction pairs
  for (; i + 3u * NT < n; i += 4u * NT, q += 4u * dq, sa_c += 8u * NT, sa_b += 32u * NT) {
    const int c0 = lds_code(sa_c), c1 = lds_code(sa_c + 2u * NT), c2 = lds_code(sa_c + 4u * NT), c3 = lds_code(sa_c + 6u * NT);
    const double v0 = lds_f64(sa_b), v1 = lds_f64(sa_b + 8u * NT), v2 = lds_f64(sa_b + 16u * NT), v3 = lds_f64(sa_b + 24u * NT);
    const unsigned long long r0 = lds_u64(sa_r + ((q >> 20) << 3)), r1 = lds_u64(sa_r + (((q + dq) >> 20) << 3)),
                             r2 = lds_u64(sa_r + (((q + 2u * dq) >> 20) << 3)), r3 = lds_u64(sa_r + (((q + 3u * dq) >> 20) << 3));
    scatter_one(c0, r0, v0);
    scatter_one(c1, r1, v1);
    scatter_one(c2, r2, v2);
    scatter_one(c3, r3, v3);
  }
  for (; i < n; i += NT, q += dq, sa_c += 2u * NT, sa_b += 8u * NT)
    scatter_one(lds_code(sa_c), lds_u64(sa_r + ((q >> 20) << 3)), lds_f64(sa_b));
  __syncwarp();
  // lane l copies 16 bytes of segment w + NW (l / 4) (+ 8 NW per round): one instruction moves eight of the warp's segments
  int seg = w + V7_NW * (lane >> 2);
  const int off = seg * 64 + (lane & 3) * 16;
  unsigned dst = sa_codes0 + (unsigned)off;
  const char* src = reinterpret_cast<const char*>(next) + off;
#pragma unroll 1
  for (; seg < next_nseg; seg += 8 * V7_NW, dst += 512 * V7_NW, src += 512 * V7_NW)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__host__ __device__ constexpr unsigned recip20(int n) { return (unsigned)(((1u << 20) + n - 1) / n); }
static_assert((2943u * recip20(81)) >> 20 == 2943u / 81 && (2943u * recip20(36)) >> 20 == 2943u / 36, "reciprocal division");
// the last chunk holds five sections: jj | j-phi | phi-j | up | pu.  The row of entry i is cell-independent: rest_row(i)
__device__ __forceinline__ int rest_row(int i) {
  constexpr int R_JF = h7::R_JF, R_FJ = h7::R_FJ, R_UP = h7::R_UP, R_PU = h7::R_PU;
  constexpr int OFF_P = h7::OFF_P, OFF_J = h7::OFF_J, OFF_F = h7::OFF_F, NLOC = h7::NLOC;
  int row;
  if (i < R_JF) row = OFF_J + i / 36;
  else if (i < R_FJ) row = OFF_J + (i - R_JF) / 8;
  else if (i < R_UP) row = OFF_F + (i - R_FJ) / 36;
  else if (i < R_PU) row = (i - R_UP) / 4;
  else row = OFF_P + (i - R_PU) / 81;
  return row > NLOC - 1 ? NLOC - 1 : row;
}
// Warp w sweeps the segments w, w + NW, ...; a lane's rows (one byte each, at most 12) are computed ONCE per kernel and travel
// in three registers -- the if-chain with its divisions cost 12 instructions per entry (950 per cell).
constexpr int REST_SEGS = h7::CH_REST_PAD / 32, REST_ROUNDS = (REST_SEGS + V7_NW - 1) / V7_NW;
static_assert(REST_ROUNDS <= 12 && h7::NLOC <= 256, "packed row table of sweep_rest");
struct RestRows { unsigned w[3]; };
__device__ __forceinline__ RestRows rest_rows_of_lane(int tid) {
  RestRows R{{0u, 0u, 0u}};
  const int w = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int k = 0; k < REST_ROUNDS; k++) {
    const int seg = w + V7_NW * k;
    const unsigned row = seg < REST_SEGS ? (unsigned)rest_row(seg * 32 + lane) : 0u;
    R.w[k >> 2] |= row << (8 * (k & 3));
  }
  return R;
}
V7_NI void sweep_rest(int tid, unsigned rw0, unsigned rw1, unsigned rw2) {
  const int w = tid >> 5, lane = tid & 31;
  const unsigned sa_codes = smem_u32(sm_codes()), sa_buf = smem_u32(sm_cell().r3), sa_r = smem_u32(sm_cell().rowaddr);
#pragma unroll
  for (int k = 0; k < REST_ROUNDS; k++) {
    const int seg = w + V7_NW * k;
    if (seg < REST_SEGS) {
      const unsigned i = (unsigned)(seg * 32 + lane);
      const unsigned word = (k >> 2) == 0 ? rw0 : ((k >> 2) == 1 ? rw1 : rw2);
      const unsigned row = (word >> (8 * (k & 3))) & 0xFFu;
      scatter_one(lds_code(sa_codes + 2u * i), lds_u64(sa_r + 8u * row), lds_f64(sa_buf + 8u * i));
    }
  }
}

struct DevRAdd {
  double* r;
  int64_t nrows;
  __device__ __forceinline__ void operator()(int row, double v) const {
    const int32_t g = sm_cell().gid[row];
    if (g >= 0 && g < nrows) atomicAdd(r + g, v);
  }
};
template <int CONV, bool ZU, bool ZJ>
V7_NI void ni_res_fields(int tid, const h7::Params& P) { h7::res_fields<CONV, ZU, ZJ>(sm_cell(), sm_small(), tid, V7_NT, P); }

// MODE 0: jacobian! | 1: residual_and_jacobian! (the residual phases ride in the barrier intervals of the Jacobian phases) |
// 2: residual! (the same residual phases alone)
template <int CONV, bool ZU, bool ZJ, int MODE>
__global__ void __launch_bounds__(V7_NT, 2)
hdiv_v7_jacobian_kernel(int64_t ncells, int64_t nrows, V7Args A, const double* __restrict__ x, const uint16_t* __restrict__ map,
                        double* __restrict__ nzval, double* __restrict__ rout, const __grid_constant__ h7::Params P) {
  using namespace h7;
  constexpr bool RES = MODE != 0, JAC = MODE != 2;
  constexpr int WU = (CONV != 0 || RES) ? 1 : 0;  // velocity and its gradient at the points are needed
  constexpr int SEG_UU = CH_UU_PAD / 32, SEG_UJ = CH_UJ_PAD / 32, SEG_REST = CH_REST_PAD / 32;
  Cell7& S = sm_cell();
  SmallDyn& C = sm_small();
  NextIds& N = sm_next();
  const Small7& K = c_small7;
  small_from_tab(C, *A.tab, threadIdx.x, V7_NT);
  __shared__ unsigned long long clk_acc[8];  // [7] = time of the last stamp
  if (threadIdx.x == 0) {
    for (int i = 0; i < 7; i++) clk_acc[i] = 0;
    clk_acc[7] = (unsigned long long)clock64();
  }
  const RestRows rest_rows = rest_rows_of_lane(threadIdx.x);  // cell-independent: rows of this lane's entries in the last sweep
  double v_pre = 0.0;  // this thread's gathered value (state / vertex coordinate) of the cell about to start
  // cell ids of this CTA's current / next / next-but-one cell: with a cell list (traversal order, colours) the id is a global
  // load, issued two cells ahead so that neither the loop top nor fetch_next ever waits for it (it cost ~400 cycles per cell)
  const int64_t stride = (int64_t)gridDim.x;
  auto cell_at = [&](int64_t i) -> int64_t { return i < ncells ? (A.cell_list ? (int64_t)A.cell_list[i] : i) : 0; };
  int64_t cell_cur = cell_at(blockIdx.x), cell_n1 = cell_at(blockIdx.x + stride), cell_n2 = 0;
  if ((int64_t)blockIdx.x < ncells) {
    fetch_next(N, A, map, cell_cur, threadIdx.x, JAC);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < LOAD_ITEMS) v_pre = load_gather(threadIdx.x, A.coords, N.nodes, N.gid, A.dir, x, WU != 0, RES);
  }
  DevRAdd radd{rout, nrows};
  for (int64_t it = blockIdx.x; it < ncells; it += gridDim.x, cell_cur = cell_n1, cell_n1 = cell_n2) {
    // re-read the thread id inside the loop: keeps the index arithmetic of the phases from being hoisted out of the cell
    // loop (and spilled) -- it is a handful of integer instructions per phase
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int64_t cell = cell_cur;
    cell_n2 = cell_at(it + 2 * stride);  // consumed two iterations from now
    const uint16_t* m = map + cell * h7::NENT;
    const bool solid = A.cell_solid != nullptr && A.cell_solid[cell] != 0;
    __syncthreads();  // the previous cell's sweeps are done with S (the ids of this cell landed in N long ago)
    load_ids(S, tid, V7_NT, N.gid, JAC ? N.rowstart : nullptr, nzval, solid, solid ? A.cell_sigma[cell] : 0.0, P.sigma);
    if (tid < LOAD_ITEMS) load_scatter(S, C, tid, v_pre, N.perm, RES);
    __syncthreads();
    if (it + gridDim.x < ncells) {
      fetch_next(N, A, map, cell_n1, tid, JAC);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    V7_CLK(0);
    phase_geom_a<WU>(S, C, tid, V7_NT, &A.tab->gg[0][0]);
    if (RES) res_pv(S, C, tid, V7_NT);
    __syncthreads();
    phase_geom_b<WU>(S, C, tid, V7_NT, A.tab->w);
    __syncthreads();
    if (WU || ZU) {
      if (WU) phase_points(S, C, tid, V7_NT);
      if (ZU) phase_Mp(S, C, tid, V7_NT);
      __syncthreads();
    }
    V7_CLK(1);
    if (JAC) ni_fields<CONV, ZJ>(tid, P);
    if (RES) res_divu(S, tid, V7_NT);
    if (ZU) phase_Minv(S, tid, V7_NT);
    __syncthreads();
    V7_CLK(2);
    if (JAC) ni_stage1<ZJ>(tid);
    if (RES) res_d(S, C, tid, V7_NT, radd);
    __syncthreads();
    if (JAC) ni_stage2<CONV, ZJ>(tid);
    if (RES) ni_res_fields<CONV, ZU, ZJ>(tid, P);
    __syncthreads();
    if (JAC) phase_D<ZU>(S, C, K, tid, V7_NT);
    if (RES) res_stage_a(S, C, tid, V7_NT, radd);
    __syncthreads();
    if (ZU || RES) {
      if (ZU && JAC) phase_E(S, tid, V7_NT, P.zeta_u);
      if (RES) res_stage_b(S, C, tid, V7_NT, radd);
      __syncthreads();
    }
    V7_CLK(3);
    if (!JAC) {
      res_stage_c(S, C, tid, V7_NT, radd);
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();  // the next cell's ids have landed
      if (it + gridDim.x < ncells && tid < LOAD_ITEMS) v_pre = load_gather(tid, A.coords, N.nodes, N.gid, A.dir, x, WU != 0, RES);
      continue;
    }
    // Every interval: fetch the codes of the chunk about to be computed (cp.async, behind the sweep of the previous chunk:
    // a warp refills only its own segments), sweep the chunk staged in the previous interval, compute the next one.
    sweep_and_fetch(0, 0, 0, 0, m + E_UU, SEG_UU, tid);
    ni_chunk_uu<CONV, ZU>(tid, 0);
    if (RES) res_stage_c(S, C, tid, V7_NT, radd);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    sweep_and_fetch(0, SEG_UU, 0, recip20(81), m + E_UU + CH_UU_PAD, SEG_UU, tid);
    ni_chunk_uu<CONV, ZU>(tid, 1);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    sweep_and_fetch(1, SEG_UU, 27, recip20(81), m + E_UU + 2 * CH_UU_PAD, SEG_UU, tid);
    ni_chunk_uu<CONV, ZU>(tid, 2);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    V7_CLK(4);
    sweep_and_fetch(0, SEG_UU, 54, recip20(81), m + E_UJ, SEG_UJ, tid);
    ni_chunk_uj(tid, P, false);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    sweep_and_fetch(1, SEG_UJ, 0, recip20(36), m + E_JU, SEG_UJ, tid);
    ni_chunk_uj(tid, P, true);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    V7_CLK(5);
    sweep_and_fetch(0, SEG_UJ, h7::OFF_J, recip20(81), m + E_REST, SEG_REST, tid);
    ni_chunk_rest<ZJ>(tid);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // the next cell's ids landed before the barrier above: gather its state now, behind the last sweep
    if (it + gridDim.x < ncells && tid < LOAD_ITEMS) v_pre = load_gather(tid, A.coords, N.nodes, N.gid, A.dir, x, WU != 0, RES);
    sweep_rest(tid, rest_rows.w[0], rest_rows.w[1], rest_rows.w[2]);
    V7_CLK(6);
  }
  if (A.clk != nullptr && threadIdx.x == 0)
    for (int i = 0; i < 7; i++) atomicAdd(A.clk + i, clk_acc[i]);
}

// nnz that receive more than one contribution are accumulated with RED and must start from zero.  Clearing them one by one
// costs more than the 1.2 GB memset it replaces (0.18 ms against 0.12 ms: 8-byte stores leave partial 32-byte sectors that
// the L2 has to fill from DRAM), so whole SECTORS are cleared: bit s of the mask <=> sector s (nnz 4 s .. 4 s + 3) holds a
// shared nnz.  Exclusive nnz inside a cleared sector are overwritten by their plain store later in the same assembly.
__global__ void __launch_bounds__(256)
zero_shared_kernel(int64_t nsectors, const uint32_t* __restrict__ mask, double* __restrict__ nz, int64_t nnz, double* __restrict__ r,
                   int64_t nr) {
  // the residual of a fused assembly is cleared by the same launch (r == nullptr: none)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += (int64_t)gridDim.x * blockDim.x) r[i] = 0.0;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nsectors; s += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t wd = __ldg(mask + (s >> 5));  // one word per warp: broadcast load
    if (!((wd >> (s & 31)) & 1u)) continue;
    const int64_t i = s * 4;
    if (i + 4 <= nnz) {
      double2* p = reinterpret_cast<double2*>(nz + i);
      p[0] = make_double2(0.0, 0.0);
      p[1] = make_double2(0.0, 0.0);
    } else {
      for (int64_t j = i; j < nnz; j++) nz[j] = 0.0;
    }
  }
}

__global__ void __launch_bounds__(256)
build_shared_mask(int64_t nsectors, int64_t nnz, const uint8_t* __restrict__ contrib, uint32_t* __restrict__ mask) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one sector per thread, one mask word per warp
  bool sh = false;
  if (s < nsectors)
    for (int64_t i = s * 4; i < s * 4 + 4 && i < nnz; i++) sh = sh || contrib[i] != 1;  // 0 contributions: keep it cleared too
  const uint32_t bits = __ballot_sync(0xffffffffu, sh);
  if ((threadIdx.x & 31) == 0 && s < (nsectors + 31) / 32 * 32) mask[s >> 5] = bits;
}

template <class K>
int v7_opt_in(K kernel) {  // (K is the same function-pointer type for every instantiation: no per-type caching here)
  MHD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V7_SMEM));
  return 0;
}

int sm_count7() { return device_sm_count(); }

}  // namespace

void v7_entry_order(std::vector<uint16_t>& ord) {
  ord.assign(h7::NENT, ORDER_PAD);
  for (int e = 0; e < h7::NENT; e++) {
    int li, lj;
    if (h7::entry_rowcol(e, &li, &lj)) ord[e] = (uint16_t)(li << 8 | lj);
  }
}

// Called at operator creation: discover the tensor structure; returns 0 also when there is none (jac_version stays 5)
int v7_try_enable(mhd_operator* op) {
  const char* e = getenv("MHD_JAC_V7");  // MHD_JAC_V7=0: keep the generic kernel (A/B runs, tests of assembly.cu)
  const bool want = e ? (atoi(e) != 0) : true;
  if (!want || op->formulation != FORM_HDIV) return 0;
  const double* h = op->h_tables.data();
  h7::Tab7* T = new h7::Tab7;
  const bool ok = h7::build_tab7(h + T_W, h + T_GG, h + T_NU, h + T_DNU, h + T_PP, h + T_PSI, h + T_DPSI, h + T_CHI, T);
  int rc = 0;
  if (ok) {
    rc = dev_alloc((unsigned char**)&op->d_tab7, (int64_t)sizeof(h7::Tab7));
    if (!rc && cudaMemcpyAsync(op->d_tab7, T, sizeof(h7::Tab7), cudaMemcpyHostToDevice, g_stream) != cudaSuccess)
      rc = cuda_fail(cudaGetLastError(), "copy of the v7 tables", __FILE__, __LINE__);
    if (!rc && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "sync", __FILE__, __LINE__);
    if (!rc) {
      h7::Small7 sm;
      h7::small_from_tab(sm, *T, 0, 1);
      op->h_small7.assign((const unsigned char*)&sm, (const unsigned char*)&sm + sizeof(sm));
      op->jac_version = 7;
    }
  }
  delete T;
  return rc;
}

// symbolic phase: bit mask of the 32-byte sectors of nzval that hold an nnz with != 1 contributions
int v7_build_shared_mask(mhd_operator* op, const uint8_t* d_contrib) {
  const int64_t nsectors = (op->nnz + 3) / 4, nwords = (nsectors + 31) / 32;
  cudaFree(op->d_shared_mask);
  op->d_shared_mask = nullptr;
  MHD_TRY(dev_alloc(&op->d_shared_mask, nwords));
  const int64_t nthreads = nwords * 32;
  build_shared_mask<<<(unsigned)((nthreads + 255) / 256), 256, 0, g_stream>>>(nsectors, op->nnz, d_contrib, op->d_shared_mask);
  MHD_LAUNCH_CHECK();
  return 0;
}

// Traversal order of the cells.  Two cells that share a dof write into the same rows, and where their column runs meet they
// share 32-byte sectors: if the second one comes while the sector is still in L2 the two partial writes merge there, otherwise
// the sector goes to DRAM half written and comes back for a read-modify-write.  In mesh order the z-neighbours of a Hunt cell
// are 4 096 cells = 14 waves of the persistent grid = 0.6 GB of writes apart.  A breadth-first order over FACE neighbours (cells
// sharing a j dof) keeps every neighbour within about two level sets (a few hundred cells, inside the 126 MB L2) on any mesh.
// MHD_V7_ORDER=0 keeps the mesh order (A/B).
int v7_build_cell_order(mhd_operator* op) {
  if (op->cell_order_tried) return 0;
  op->cell_order_tried = true;
  const char* e = getenv("MHD_V7_ORDER");
  if (e && atoi(e) == 0) return 0;
  const int64_t nc = op->ncells;
  if (nc < 2) return 0;
  std::vector<int32_t> g((size_t)nc * NLOC);
  MHD_TRY(d2h(g.data(), op->d_gids, nc * NLOC));
  MHD_CUDA(cudaStreamSynchronize(g_stream));
  // the (at most two) cells of every face dof
  std::vector<int32_t> owner((size_t)op->ncols * 2, -1);
  for (int64_t c = 0; c < nc; c++)
    for (int k = 0; k < 36; k++) {
      const int32_t id = g[(size_t)c * NLOC + OFF_J + k];
      if (id < 0) continue;
      int32_t* o = &owner[(size_t)id * 2];
      if (o[0] < 0) o[0] = (int32_t)c;
      else if (o[0] != (int32_t)c && o[1] < 0) o[1] = (int32_t)c;
    }
  std::vector<int32_t> order;
  order.reserve((size_t)nc);
  std::vector<uint8_t> seen((size_t)nc, 0);
  for (int64_t seed = 0; seed < nc; seed++) {  // one breadth-first sweep per connected component
    if (seen[seed]) continue;
    seen[seed] = 1;
    size_t head = order.size();
    order.push_back((int32_t)seed);
    while (head < order.size()) {
      const int32_t c = order[head++];
      for (int k = 0; k < 36; k++) {
        const int32_t id = g[(size_t)c * NLOC + OFF_J + k];
        if (id < 0) continue;
        for (int s = 0; s < 2; s++) {
          const int32_t n = owner[(size_t)id * 2 + s];
          if (n >= 0 && !seen[n]) {
            seen[n] = 1;
            order.push_back(n);
          }
        }
      }
    }
  }
  MHD_CHECK((int64_t)order.size() == nc, MHD_E_STATE, "cell order: %lld of %lld cells", (long long)order.size(), (long long)nc);
  MHD_TRY(dev_alloc(&op->d_cell_order, nc));
  MHD_TRY(h2d(op->d_cell_order, order.data(), nc));
  MHD_CUDA(cudaStreamSynchronize(g_stream));
  return 0;
}

int v7_zero_shared(mhd_operator* op, cudaStream_t stream, double* d_r) {
  const int64_t nsectors = (op->nnz + 3) / 4;
  const int64_t want = (nsectors + 255) / 256;
  const int64_t cap = (int64_t)sm_count7() * 64;
  zero_shared_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(nsectors, op->d_shared_mask, op->d_nzval, op->nnz, d_r,
                                                                               d_r ? op->nrows : 0);
  MHD_LAUNCH_CHECK();
  return 0;
}

// mode 0: Jacobian (d_r unused) | 1: residual + Jacobian | 2: residual only
int v7_launch(mhd_operator* op, const double* d_x, double* d_r, int mode) {
  MHD_CHECK(op->jac_version == 7 && op->d_tab7 != nullptr && (mode == 2 || op->d_shared_mask != nullptr), MHD_E_STATE,
            "v7 kernel is not enabled on this operator");
  if (mode == 0) d_r = nullptr;
  // clearing (shared sectors of nzval, the residual): when the caller passes host vectors it was started on a side stream BEFORE
  // the copy of the state was enqueued and overlaps it (abi.cu); otherwise it runs here, in line
  if (!op->clear_pending) MHD_TRY(begin_clear(op, d_r, mode != 2, false));
  op->clear_pending = false;
  MHD_TRY(end_clear());
  h7::Params P;
  P.alpha = op->prm.alpha; P.beta = op->prm.beta; P.gamma = op->prm.gamma; P.sigma = op->prm.sigma;
  P.zeta_u = op->prm.zeta_u; P.zeta_j = op->prm.zeta_j;
  for (int i = 0; i < 3; i++) { P.B[i] = op->prm.B[i]; P.f[i] = op->prm.f[i]; P.g[i] = op->prm.g[i]; }
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("MHD_JAC_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  static unsigned long long* d_clk = nullptr;
  if ((dbg & 16) && !d_clk) MHD_CUDA(cudaMalloc((void**)&d_clk, 8 * sizeof(unsigned long long)));
  if (dbg & 16) MHD_CUDA(cudaMemsetAsync(d_clk, 0, 8 * sizeof(unsigned long long), g_stream));
  if (g_small7_device != g_device || g_small7_loaded != op->h_small7) {
    g_small7_loaded = op->h_small7;  // the copy below reads this buffer asynchronously: it must outlive the call
    MHD_CUDA(cudaMemcpyToSymbolAsync(c_small7, g_small7_loaded.data(), sizeof(h7::Small7), 0, cudaMemcpyHostToDevice, g_stream));
    MHD_CUDA(cudaStreamSynchronize(g_stream));
    g_small7_device = g_device;
  }
  MHD_TRY(v7_build_cell_order(op));
  V7Args A{op->d_coords, op->d_cell_nodes, op->d_pgids, (const long long*)op->d_rowstart, op->d_perm, op->d_cell_solid,
           op->d_cell_sigma, op->d_dir, (const h7::Tab7*)op->d_tab7, op->d_cell_order, (dbg & 16) ? d_clk : nullptr};
  const int64_t g64 = (int64_t)sm_count7() * 2;
  const int conv = op->prm.convection;
  const bool zu = op->prm.zeta_u != 0.0, zj = op->prm.zeta_j != 0.0;
  int64_t ncells = op->ncells;
  unsigned grid = (unsigned)(ncells < g64 ? ncells : g64);
#define VKM(C, U, J, M)                                                                                                    \
  do {                                                                                                                     \
    MHD_TRY(v7_opt_in(hdiv_v7_jacobian_kernel<C, U, J, M>));                                                                \
    hdiv_v7_jacobian_kernel<C, U, J, M><<<grid, V7_NT, V7_SMEM, g_stream>>>(ncells, op->nrows, A, d_x, op->d_map,           \
                                                                            op->d_nzval, d_r, P);                          \
  } while (0)
#define VK(C, U, J) do { if (mode == 0) VKM(C, U, J, 0); else if (mode == 1) VKM(C, U, J, 1); else VKM(C, U, J, 2); } while (0)
#define VKJ(C, U) do { if (zj) VK(C, U, true); else VK(C, U, false); } while (0)
#define VKU(C) do { if (zu) VKJ(C, true); else VKJ(C, false); } while (0)
#define VKC() do { if (conv == 0) VKU(0); else if (conv == 1) VKU(1); else VKU(2); } while (0)
  prof_begin(mode == 2 ? PROF_RES : PROF_JAC);
  if (op->deterministic && op->d_color_cells != nullptr) {
    // one launch per colour: cells of a colour share no dof, colours run in stream order => a fixed summation order
    for (size_t c = 0; c + 1 < op->color_ptr.size(); c++) {
      ncells = op->color_ptr[c + 1] - op->color_ptr[c];
      if (ncells == 0) continue;
      A.cell_list = op->d_color_cells + op->color_ptr[c];
      grid = (unsigned)(ncells < g64 ? ncells : g64);
      VKC();
      MHD_LAUNCH_CHECK();
    }
  } else {
    VKC();
    MHD_LAUNCH_CHECK();
  }
  prof_end(mode == 2 ? PROF_RES : PROF_JAC);
#undef VKC
#undef VKU
#undef VKJ
#undef VK
#undef VKM
  if (dbg & 16) {
    unsigned long long h[8];
    MHD_CUDA(cudaMemcpyAsync(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
    MHD_CUDA(cudaStreamSynchronize(g_stream));
    const double per = 1.0 / (double)op->ncells;
    fprintf(stderr, "[mhd v7 phase clocks / cell] load %.0f  geometry+points %.0f  fields %.0f  stages+D %.0f  uu %.0f  uj/ju %.0f  rest+tail %.0f\n",
            h[0] * per, h[1] * per, h[2] * per, h[3] * per, h[4] * per, h[5] * per, h[6] * per);
  }
  return 0;
}

}  // namespace mhd
