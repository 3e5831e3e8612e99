// FP64 sm_100a assembly kernels: cell-wise integration of the Jacobian and residual of the inductionless MHD
// H1-HDiv weak form, scattered into CSR / the residual vector through the precomputed map.
//
// Integrands: jac_fluid_h1_hdiv (src/weakforms.jl:283-312), res_fluid_h1_hdiv (src/weakforms.jl:255-281),
// solid cells jac/res_solid_h1_hdiv (:314-338), conv (weakforms.jl:670), local projection (weakforms.jl:672-681).
// Notation of SURVEY.md Appendix A.
//
// Design (persistent CTAs, 2 per SM, 8 warps, one cell at a time):
//   prep   : geometry Jacobians at the 27 Gauss points, physical gradients of the Q2 basis, Piola-mapped RT basis, all
//            pre-scaled by sqrt(w_q |det J_q|) so that every block is a plain product sum_k A[k][m] B[k][n] of two
//            shared-memory panels.  The reference tables arrive in panel layout with ONE bulk copy (cp.async.bulk +
//            mbarrier) per cell, issued before the id / state loads, and are transformed IN PLACE AND PERMUTED: a warp
//            owns the rows of a quadrature point, every lane gathers the reference column of its slot, the warp
//            synchronises and writes the slot columns.  Slots follow the cell's permutation (each field sorted by
//            global id, symbolic.cu), so consecutive columns of a product are consecutive nnz of a CSR row and the
//            operand fragments are plain strided loads.
//   residual: point values as (panel rows) x (state) products, a per-point coefficient stage, then the row products
//            panel^T x coefficients -- all on the tensor cores; in the fused kernel the row products are jobs of the
//            main phase.
//   main   : one barrier-free phase.  Every warp runs tensor-core jobs (FP64 mma.sync m8n8k4): its tile of the uu block
//            (all 9 component pairs) and then jobs drawn from a shared counter (jj strips, uj/ju, up/pu, j-phi/phi-j).
//            A job's accumulators are transposed through a warp-private staging tile into destination order
//            (row-major, velocity components fastest) and swept out with 32 consecutive entries per store / RED
//            instruction; the 16-bit map codes of a job are fetched before its MMAs.
#include <stdlib.h>

#include "common.h"

namespace mhd {

int pack_tables(mhd_operator* op, const mhd_tables_t* t) {
  MHD_CHECK(t->w && t->geo_grad && t->u_val && t->u_grad && t->p_val && t->j_val && t->j_div && t->phi_val,
            MHD_E_INVALID, "mhd_tables_t has a null table");
  std::vector<double> h(T_TOTAL);
  memcpy(&h[T_W], t->w, NQ * sizeof(double));
  memcpy(&h[T_GG], t->geo_grad, NQ * 24 * sizeof(double));
  memcpy(&h[T_NU], t->u_val, NQ * 27 * sizeof(double));
  memcpy(&h[T_DNU], t->u_grad, NQ * 81 * sizeof(double));
  memcpy(&h[T_PP], t->p_val, NQ * 4 * sizeof(double));
  memcpy(&h[T_PSI], t->j_val, NQ * 108 * sizeof(double));
  memcpy(&h[T_DPSI], t->j_div, NQ * 36 * sizeof(double));
  memcpy(&h[T_CHI], t->phi_val, NQ * 8 * sizeof(double));
  std::vector<double> pt(PT_TOTAL, 0.0);
  for (int q = 0; q < NQ; q++) {
    for (int v = 0; v < 8; v++)
      for (int k = 0; k < 3; k++) h[T_GGT + (v * 3 + k) * 27 + q] = t->geo_grad[(q * 8 + v) * 3 + k];
    for (int a = 0; a < 27; a++) {
      pt[PT_N + q * 28 + a] = t->u_val[q * 27 + a];
      for (int k = 0; k < 3; k++) pt[PT_G + (q * 3 + k) * 28 + a] = t->u_grad[(q * 27 + a) * 3 + k];
    }
    for (int m = 0; m < 36; m++) {
      pt[PT_DIV + q * 36 + m] = t->j_div[q * 36 + m];
      for (int k = 0; k < 3; k++) pt[PT_PSI + (q * 3 + k) * 36 + m] = t->j_val[(q * 36 + m) * 3 + k];
    }
    for (int k = 0; k < 4; k++) pt[PT_PP + q * 4 + k] = t->p_val[q * 4 + k];
    for (int l = 0; l < 8; l++) pt[PT_CHI + q * 8 + l] = t->phi_val[q * 8 + l];
  }
  MHD_TRY(dev_alloc(&op->d_ptab, PT_TOTAL));
  MHD_TRY(h2d(op->d_ptab, pt.data(), PT_TOTAL));
  MHD_TRY(dev_alloc(&op->d_tables, T_TOTAL));
  MHD_TRY(h2d(op->d_tables, h.data(), T_TOTAL));
  MHD_CUDA(cudaStreamSynchronize(g_stream));
  op->h_tables = h;  // kept for the structure discovery of hdiv_v7.cu (v7_try_enable)
  return 0;
}

// The order in which the Jacobian kernel consumes the scatter map (layout in common.h); slots are indices of the
// permuted local numbering.  symbolic.cu resolves every (row slot, col slot) to its nnz.
void entry_order(std::vector<uint16_t>& ord) {
  ord.assign(NENT, ORDER_PAD);
  auto put = [&](int e, int ri, int ci) { ord[e] = (uint16_t)((ri << 8) | ci); };
  for (int w = 0; w < 8; w++) {
    const int mt = w >> 1, np = w & 1, nrow = uu_nrow(mt), ncol = uu_ncol(np);
    for (int c = 0; c < 3; c++)
      for (int r = 0; r < nrow; r++)
        for (int col = 0; col < ncol; col++)
          put(uu_base(w, c) + r * ncol + col, c * 27 + 8 * mt + r, (col % 3) * 27 + 16 * np + col / 3);
  }
  for (int m = 0; m < NJ; m++)
    for (int n = 0; n < NJ; n++) put(jj_base(m / 8) + (m % 8) * NJ + n, OFF_J + m, OFF_J + n);
  for (int s = 0; s < 5; s++)
    for (int h = 0; h < 2; h++) {
      const int nm = uj_nm(s), na = uj_na(h), b0 = uj_base(s, h);
      for (int c = 0; c < 3; c++)
        for (int al = 0; al < na; al++)
          for (int ml = 0; ml < nm; ml++) put(b0 + (c * na + al) * nm + ml, c * 27 + 16 * h + al, OFF_J + 8 * s + ml);
      const int b1 = b0 + uj_part(s, h);
      for (int ml = 0; ml < nm; ml++)
        for (int col = 0; col < 3 * na; col++) put(b1 + ml * 3 * na + col, OFF_J + 8 * s + ml, (col % 3) * 27 + 16 * h + col / 3);
    }
  for (int i = 0; i < NU; i++)
    for (int k = 0; k < NP; k++) put(SEC_UP + i * NP + k, i, OFF_P + k);
  for (int k = 0; k < NP; k++)
    for (int col = 0; col < NU; col++) put(SEC_PU + k * NU + col, OFF_P + k, (col % 3) * 27 + col / 3);
  for (int m = 0; m < NJ; m++)
    for (int l = 0; l < NF; l++) put(SEC_JF + m * NF + l, OFF_J + m, OFF_F + l);
  for (int l = 0; l < NF; l++)
    for (int m = 0; m < NJ; m++) put(SEC_FJ + l * NJ + m, OFF_F + l, OFF_J + m);
}

struct KParams {
  double alpha, beta, gamma, sigma, zeta_u, zeta_j;
  double B[3], f[3], g[3];
  const uint8_t* cell_solid;  // null: no solid sub-domain
  const double* cell_sigma;
  int dbg;  // MHD_JAC_DEBUG bit mask (timing experiments only): 2 no main phase, 4 preparation only once, 32 no L2 prefetch,
            // 16 phase clocks of thread 0 (summed over CTAs into clk_out, printed by the launcher)
  unsigned long long* clk_out;
};

constexpr int NT = 256;  // threads per CTA

// ---- shared-memory plan (doubles)
constexpr int LDN = 28;                       // padded leading dimension of 27-wide panels
// the first PT_TOTAL doubles mirror the panel-layout tables (PT_*): destination of the per-cell bulk copy
constexpr int S_G = PT_G;                     // [81][28]  sqrt(w) dN_a/dx_i, row = q*3+i
constexpr int S_N = PT_N;                     // [27][28]  sqrt(w) N_a
constexpr int S_PSI = PT_PSI;                 // [81][36]  sqrt(w) Piola(psi_m)_i, row = q*3+i
constexpr int S_DIV = PT_DIV;                 // [27][36]  sqrt(w) div psi_m   (contiguous after PSI)
constexpr int S_PP = PT_PP;                   // [27][4]
constexpr int S_CHI = PT_CHI;                 // [27][8]
constexpr int S_UG = PT_TOTAL;                // [27][28]  sqrt(w) (u_q . grad N_b)
constexpr int S_T = S_UG + 27 * LDN;           // [27][9]   (d_d u_c)(q), index d*3+c
constexpr int LDT = 10;                       // padded row of the velocity-gradient table
constexpr int SWN = 400;                      // warp-private staging tile (doubles)
constexpr int S_ST = S_T + 27 * LDT + 2;      // staging: 8 warps x SWN (the residual's scratch during the preparation)
constexpr int ST_SIZE = 8 * SWN;
constexpr int S_J = S_ST + ST_SIZE;           // [27][9] J sqrt(w |det|) / det
constexpr int S_INV = S_J + 243;              // [27][9] J^-1 sqrt(w |det|)
constexpr int S_DET = S_INV + 243;            // [27] sqrt(w |det J|) / det J
constexpr int S_SW = S_DET + 27;              // [27]
constexpr int S_X = S_SW + 27;                // [8][3]
constexpr int S_U = S_X + 24;                 // [129] local state
constexpr int S_SG = S_U + 130;               // [36] sign
constexpr int S_E = S_SG + 36;                // [4][81]  zeta_u M_p^-1 D, column = d*27+b
constexpr int S_D = S_E + 324;                // [81][4]  D[(c,a)][k] = sum_q w pi_k d_c N_a
constexpr int S_MI = S_D + 324;               // [16]
constexpr int S_UQ = S_MI + 16;               // [27][3] u at q (unweighted)
constexpr int S_SC = S_UQ + 82;               // [108] scale vector for the jj product
constexpr int S_GGT = S_SC + 108;             // [8][3][27] geometry-map gradients (persistent: loaded once per CTA)
constexpr int S_W = S_GGT + 648;              // [27] quadrature weights (persistent)
constexpr int S_END = S_W + 28;
constexpr int SMEM_BYTES = S_END * 8 + NLOC * 8 /*row starts*/ + 132 * 4 /*gids*/ + PERM_STRIDE + 16 /*job counter*/ +
                           128 /*phase clocks*/ + 16 /*mbarrier*/ + 132 * 4 /*next gids*/;

// ---------------------------------------------------------------------------------------------
// FP64 tensor-core path: mma.sync.m8n8k4 (DMMA).  Same peak rate as the FP64 FMA pipe on B200, but an 8x8x4 product
// costs one operand word per lane and operand fragments are shared by all tiles of a warp's output block.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// One warp computes the (8 MT) x (8 NTL) output block at (m0, n0) of  C = sum_k A[k][m] * sc[k] * B[k][n]
// (panels k-major in shared memory).  Fragment loads are bank-conflict free for leading dimensions = 4 or 12 mod 16
// (28, 36, 108, 4).  Rows/columns beyond M/N read neighbouring panel data and are discarded; k >= K contributes 0.
template <int MT, int NTL, bool SCALE>
__device__ __forceinline__ void warp_mma_acc(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                             int K, const double* __restrict__ sc, int scs, int m0, int n0,
                                             double (&acc)[MT][NTL][2]) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lk = lane & 3;
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NTL; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  const double* ap = A + m0 + lr;
  const double* bp = B + n0 + lr;
#pragma unroll 2
  for (int k0 = 0; k0 < K; k0 += 4) {
    const int kk = k0 + lk;
    const bool valid = kk < K;
    const int kc = valid ? kk : 0;
    double a[MT], b[NTL];
    const double s = SCALE ? sc[kc * scs] : 1.0;
#pragma unroll
    for (int i = 0; i < MT; i++) {
      const double v = ap[kc * lda + 8 * i];
      a[i] = valid ? v : 0.0;
    }
#pragma unroll
    for (int j = 0; j < NTL; j++) {
      const double v = bp[kc * ldb + 8 * j];
      b[j] = valid ? (SCALE ? v * s : v) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
      for (int j = 0; j < NTL; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// Same product with GATHERED operand columns: this lane's column of row tile i is A[.][acol[i]], of column tile j
// B[.][bcol[j]] (the caller resolves slots of the permuted numbering to panel columns once per job).
// ALIAS0 >= 0: A and B are the same panel and row tile i of A is column tile ALIAS0 + i of B (products P^T P): the A
// fragments are taken from the B fragments instead of being loaded again.
template <int MT, int NTL, bool SCALE, int ALIAS0 = -1>
__device__ __forceinline__ void warp_mma_cols(const double* __restrict__ A, int lda, const int (&acol)[MT],
                                              const double* __restrict__ B, int ldb, const int (&bcol)[NTL], int K,
                                              const double* __restrict__ sc, int scs, double (&acc)[MT][NTL][2]) {
  const int lk = threadIdx.x & 3;
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NTL; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 2
  for (int k0 = 0; k0 < K; k0 += 4) {
    const int kk = k0 + lk;
    const bool valid = kk < K;
    const int kc = valid ? kk : 0;
    double a[MT], b[NTL], braw[NTL];
    const double s = SCALE ? sc[kc * scs] : 1.0;
#pragma unroll
    for (int j = 0; j < NTL; j++) {
      const double v = B[kc * ldb + bcol[j]];
      braw[j] = valid ? v : 0.0;
      b[j] = SCALE ? braw[j] * s : braw[j];
    }
#pragma unroll
    for (int i = 0; i < MT; i++) {
#ifdef MHD_NO_ALIAS
      if (false) {
#else
      if (ALIAS0 >= 0 && ALIAS0 + i < NTL) {
#endif
        a[i] = braw[ALIAS0 + i < NTL ? ALIAS0 + i : 0];
      } else {
        const double v = A[kc * lda + acol[i]];
        a[i] = valid ? v : 0.0;
      }
    }
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
      for (int j = 0; j < NTL; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// ---- bulk asynchronous copy (TMA engine, no tensor map) + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// Velocity gradient at the 27 Gauss points through the tensor cores:
//   Gq[q][d*3+c] = sum_b G'[(q,d)][b] u_c,b   (= sqrt(w) d_d u_c),   T[q][d*3+c] = Gq / sqrt(w)
// an 81 x 3 x 27 product: A = G' read row-major ((q,d) rows, b contiguous: conflict-free fragment loads since
// LDN = 12 mod 16), B = the local velocity dofs.  11 row tiles dealt to the 8 warps.  Computed ONCE per cell and
// shared by the residual and the Newton block of the fused kernel.
__device__ __forceinline__ void velocity_gradient_mma(const double* __restrict__ G, const double* __restrict__ U,
                                                      const double* __restrict__ sw, double* __restrict__ Gq,
                                                      double* __restrict__ T, int ldt, double tscale = 1.0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lr = lane >> 2, lk = lane & 3;
  for (int mt = warp; mt < 11; mt += NT / 32) {
    double c0 = 0.0, c1 = 0.0;
    const double* ap = G + (8 * mt + lr) * LDN;  // row (q,d) = 8 mt + lr (rows >= 81 read the next panel: discarded)
    const double* bp = U + lr * 27;              // column c = lr (columns >= 3 discarded)
#pragma unroll
    for (int k0 = 0; k0 < 28; k0 += 4) {
      const int kk = k0 + lk;
      const bool valid = kk < 27;
      const double a = valid ? ap[kk] : 0.0;
      const double b = (valid && lr < 3) ? bp[kk] : 0.0;
      dmma884(c0, c1, a, b);
    }
    const int m = 8 * mt + lr;  // accumulator row; columns 2 lk, 2 lk + 1
    if (m < 81) {
      const int q = m / 3, d = m - q * 3;
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const int c = 2 * lk + r;
        if (c < 3) {
          const double v = r ? c1 : c0;
          if (Gq) Gq[q * 9 + d * 3 + c] = v;
          if (T) T[q * ldt + d * 3 + c] = tscale * v / sw[q];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
struct CellCtx {
  double* sm;
  long long* row;  // [129] byte address of the first nnz of each local row, PERMUTED numbering (0: dropped)
  int32_t* gid;    // [129] permuted local numbering (state, residual, row starts)
  uint8_t* perm;   // [PERM_STRIDE] slot -> reference basis index
  int32_t* gidn;   // [129] dof ids of the CTA's next cell (fetched while the geometry is computed)
  bool have_next;  // gidn is valid for the cell being prepared
  uint64_t* mbar;  // completion of the table bulk copy
  uint32_t parity;
  long long* clk;  // [16] phase clock accumulators (MHD_JAC_DEBUG & 16), thread 0 only
  bool clk_on;
};

// phase timing (timing experiments only): thread 0 accumulates the cycles since the previous mark into slot i
#define MHD_MARK(cx, i)                                   \
  do {                                                    \
    if ((cx).clk_on && threadIdx.x == 0) {                \
      const long long _t = clock64();                     \
      (cx).clk[i] += _t - (cx).clk[15];                   \
      (cx).clk[15] = _t;                                  \
    }                                                     \
  } while (0)

__device__ __forceinline__ CellCtx make_ctx(double* smem) {
  CellCtx cx;
  cx.sm = smem;
  cx.row = (long long*)(smem + S_END);
  cx.gid = (int32_t*)(cx.row + NLOC);
  cx.perm = (uint8_t*)(cx.gid + 132);
  cx.clk = (long long*)(cx.perm + PERM_STRIDE + 16);
  cx.mbar = (uint64_t*)(cx.clk + 16);
  cx.gidn = (int32_t*)(cx.mbar + 2);
  cx.have_next = false;
  cx.parity = 0;
  cx.clk_on = false;
  return cx;
}

struct CellArgs {
  const double* tab;         // d_tables (T_*)
  const double* ptab;        // d_ptab (PT_*)
  const double* coords;
  const int32_t* cell_nodes;
  const int32_t* gids;       // permuted numbering (d_pgids): local dofs of each field sorted by global id
  const int64_t* rowstart;   // [ncells][129] first nnz of each local row (permuted numbering), -1: dropped
  const uint8_t* perm;
  const int8_t* jsign;
  const double* dirv;
};

// once per CTA: mbarrier, persistent geometry tables
__device__ __forceinline__ void prep_init(const CellCtx& cx, const CellArgs& A) {
  if (threadIdx.x == 0) mbar_init(cx.mbar, 1);
  for (int i = threadIdx.x; i < 648; i += NT) cx.sm[S_GGT + i] = A.tab[T_GGT + i];
  if (threadIdx.x < NQ) cx.sm[S_W + threadIdx.x] = A.tab[T_W + threadIdx.x];
  __syncthreads();
}

// loads + geometry + mapped bases.  NEED_STATE: 0 none, 1 u only, 2 all fields.  rowptr != null: also the nnz row starts
// (permuted numbering).  Must be entered right after a CTA barrier that retires every use of the previous panels.
template <int NEED_STATE>
__device__ __forceinline__ void cell_prep(CellCtx& cx, int64_t cell, int64_t next_cell /* -1: none */, const CellArgs& A,
                                          const double* __restrict__ x, const double* __restrict__ nz /* null: no row starts */) {
  double* sm = cx.sm;
  const int tid = threadIdx.x;
  if (tid == NT - 1) bulk_load(sm, A.ptab, PT_TOTAL * 8, cx.mbar);  // reference tables -> panel buffers (async)
  MHD_MARK(cx, 10);
  if (tid < NLOC) {
    const int i = tid;
    const int32_t g = cx.have_next ? cx.gidn[i] : A.gids[cell * NLOC + i];
    cx.gid[i] = g;
    if (NEED_STATE == 2 || (NEED_STATE == 1 && i < NU)) sm[S_U + i] = g >= 0 ? x[g] : A.dirv[-g - 1];
    if (nz) {
      const int64_t rs = A.rowstart[cell * NLOC + i];
      cx.row[i] = rs >= 0 ? (long long)(nz + rs) : 0;
    }
  } else if (tid < NLOC + 24) {
    const int t = tid - NLOC;
    sm[S_X + t] = A.coords[(int64_t)A.cell_nodes[cell * 8 + t / 3] * 3 + t % 3];
  } else if (tid < NLOC + 24 + PERM_STRIDE / 4) {
    const int t = tid - NLOC - 24;
    reinterpret_cast<uint32_t*>(cx.perm)[t] = reinterpret_cast<const uint32_t*>(A.perm + cell * PERM_STRIDE)[t];
  }
  MHD_MARK(cx, 11);
  __syncthreads();
  MHD_MARK(cx, 1);
  if (tid < NQ) {
    const int q = tid;
    double J[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) J[i][k] = 0.0;
#pragma unroll
    for (int v = 0; v < 8; v++) {
      double gk[3];
#pragma unroll
      for (int k = 0; k < 3; k++) gk[k] = sm[S_GGT + (v * 3 + k) * 27 + q];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double xv = sm[S_X + v * 3 + i];
#pragma unroll
        for (int k = 0; k < 3; k++) J[i][k] = __dadd_rn(__dmul_rn(xv, gk[k]), J[i][k]);  // rounded like the reference: hdiv7_cell.h
      }
    }
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double id = 1.0 / det;
    // inv[k][i] = d xi_k / d x_i  (inverse of J[i][k])
    double inv[3][3];
    inv[0][0] = c00 * id;
    inv[1][0] = c01 * id;
    inv[2][0] = c02 * id;
    inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    const double sw = sqrt(sm[S_W + q] * fabs(det));
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        sm[S_J + q * 9 + i * 3 + k] = J[i][k] * (sw * id);   // Piola factor sqrt(w|det|)/det folded in
        sm[S_INV + q * 9 + k * 3 + i] = inv[k][i] * sw;
      }
    sm[S_SW + q] = sw;
    sm[S_DET + q] = sw * id;  // sqrt(w |det|) / det
    MHD_MARK(cx, 9);
  } else if (NEED_STATE >= 1 && tid >= 32 && tid < 32 + 81) {
    // u at the quadrature points (unweighted) from the still untransformed N table, by warps without geometry work
    mbar_wait(cx.mbar, cx.parity);
    const int t = tid - 32, q = t / 3, i = t - q * 3;
    double s = 0.0;
#pragma unroll 9
    for (int a = 0; a < 27; a++) s = fma(sm[S_N + q * LDN + cx.perm[a]], sm[S_U + i * 27 + a], s);
    sm[S_UQ + t] = s;
  }
  // threads without geometry work: ids of the CTA's next cell -> shared memory, and its state / vertex lines -> L2
  cx.have_next = next_cell >= 0;
  if (next_cell >= 0) {
    if (tid >= 113 && tid < 113 + NLOC) {
      const int32_t gn = A.gids[next_cell * NLOC + tid - 113];
      cx.gidn[tid - 113] = gn;
      if (NEED_STATE >= 1 && gn >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + gn));
    } else if (tid >= 113 + NLOC && tid < 113 + NLOC + 8) {
      const int32_t nd = A.cell_nodes[next_cell * 8 + tid - 113 - NLOC];
      asm volatile("prefetch.global.L2 [%0];" ::"l"(A.coords + (int64_t)nd * 3));
    }
  }
  __syncthreads();
  MHD_MARK(cx, 2);
  mbar_wait(cx.mbar, cx.parity);  // (complete long ago) makes the bulk copy visible to this thread
  cx.parity ^= 1;
  // In-place transforms WITH the column permutation (slot -> reference basis function): one warp owns the rows of one
  // quadrature point; every lane gathers its (reference) column entries, the warp synchronises, then the transformed
  // values are written to the slot column.  Panels end up in the permuted numbering: fragment loads are plain strided.
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int ar = lane < 27 ? cx.perm[lane] : 27;                      // Q2 column of this lane (27 = zero pad column)
    const int m1 = lane + 32;                                           // second RT column (lanes 0..3)
    const int pm0 = cx.perm[27 + lane], pm1 = lane < 4 ? cx.perm[27 + m1] : 0;
    const double sg0 = (pm0 & 0x80) ? -1.0 : 1.0, sg1 = (pm1 & 0x80) ? -1.0 : 1.0;
    const int mr0 = pm0 & 0x7F, mr1 = pm1 & 0x7F;
    for (int q = warp; q < NQ; q += NT / 32) {
      const double d0 = sm[S_G + (q * 3 + 0) * LDN + ar], d1 = sm[S_G + (q * 3 + 1) * LDN + ar], d2 = sm[S_G + (q * 3 + 2) * LDN + ar];
      const double nn = sm[S_N + q * LDN + ar];
      const double p0 = sm[S_PSI + (q * 3 + 0) * NJ + mr0], p1 = sm[S_PSI + (q * 3 + 1) * NJ + mr0], p2 = sm[S_PSI + (q * 3 + 2) * NJ + mr0];
      const double dv0 = sm[S_DIV + q * NJ + mr0];
      double r0 = 0.0, r1 = 0.0, r2 = 0.0, dv1 = 0.0;
      if (lane < 4) {
        r0 = sm[S_PSI + (q * 3 + 0) * NJ + mr1]; r1 = sm[S_PSI + (q * 3 + 1) * NJ + mr1]; r2 = sm[S_PSI + (q * 3 + 2) * NJ + mr1];
        dv1 = sm[S_DIV + q * NJ + mr1];
      }
      const double* inv = sm + S_INV + q * 9;
      const double* J = sm + S_J + q * 9;
      const double sw = sm[S_SW + q], pdet = sm[S_DET + q];
      __syncwarp();
      if (lane < LDN) {
#pragma unroll
        for (int i = 0; i < 3; i++) sm[S_G + (q * 3 + i) * LDN + lane] = d0 * inv[0 * 3 + i] + d1 * inv[1 * 3 + i] + d2 * inv[2 * 3 + i];
        sm[S_N + q * LDN + lane] = nn * sw;
        if (lane == 27) sm[S_UG + q * LDN + lane] = 0.0;
      }
#pragma unroll
      for (int i = 0; i < 3; i++) sm[S_PSI + (q * 3 + i) * NJ + lane] = sg0 * (J[i * 3 + 0] * p0 + J[i * 3 + 1] * p1 + J[i * 3 + 2] * p2);
      sm[S_DIV + q * NJ + lane] = dv0 * sg0 * pdet;
      if (lane < 4) {
#pragma unroll
        for (int i = 0; i < 3; i++) sm[S_PSI + (q * 3 + i) * NJ + m1] = sg1 * (J[i * 3 + 0] * r0 + J[i * 3 + 1] * r1 + J[i * 3 + 2] * r2);
        sm[S_DIV + q * NJ + m1] = dv1 * sg1 * pdet;
      }
    }
  }
  for (int idx = tid; idx < NQ * 12; idx += NT) sm[S_PP + idx] *= sm[S_SW + (idx < 108 ? idx / 4 : (idx - 108) / 8)];
  __syncthreads();
  MHD_MARK(cx, 3);
}

// ---- residual scratch.  Point values (stage 1) live in the staging area (free until the main phase):
constexpr int RP_GQ = 0;            // [81][3]  sqrt(w) d_d u_c, row (q,d)
constexpr int RP_JQ = RP_GQ + 243;  // [81]     sqrt(w) j_i(q), row (q,i)
constexpr int RP_DJ = RP_JQ + 88;   // [27]     sqrt(w) div j   (tiles of 8 rows: 32 slots each)
constexpr int RP_FQ = RP_DJ + 32;   // [27]     sqrt(w) phi
constexpr int RP_PQ = RP_FQ + 32;   // [27]     sqrt(w) p
constexpr int RP_PR = RP_PQ + 32;   // [27]     sqrt(w) Pi_p(div u)   (zeta_u only)
constexpr int RP_RH = RP_PR + 32;   // [4]
constexpr int RES_GQ = RP_GQ;
// coefficient tables (stage 2) live where the geometry factors were (S_J, S_INV, S_DET: dead after the panel transforms)
// so that the row products (stage 3) can run as jobs of the main phase:
constexpr int RC_GU = S_J;            // [81][3]  coefficient of G'[(q,d)][a] in r_u[(c,a)], row (q,d), col c
constexpr int RC_FU = RC_GU + 243;    // [27][3]  coefficient of N'[q][a]        (contiguous after GU: one K = 108 product)
constexpr int RC_FJ = RC_FU + 81;     // [81]     coefficient of Psi'[(q,i)][m]
constexpr int RC_DJ = RC_FJ + 81;     // [27]     coefficient of Div'[q][m]     (contiguous after FJ)
constexpr int RC_DV = RC_DJ + 27;     // [27]     sqrt(w) div u
constexpr int RC_JD = RC_DV + 27;     // [27]     sqrt(w) div j
static_assert(RC_JD + 27 <= S_SW, "residual coefficient tables overflow the geometry scratch");
constexpr int NRESJOBS = 11;          // 4 r_u tiles, 5 r_j tiles, r_p, r_phi

// =============================================================================================
// Residual of one cell: res_fluid_h1_hdiv (src/weakforms.jl:255-281), res_solid_h1_hdiv (:314-325) on solid cells.
// Needs cell_prep<2> (all panels + the full local state), three stages:
//   1. point values at the 27 Gauss points as (panel rows) x (state) products on the tensor cores:
//      sqrt(w) grad u, j, div j, phi, p                                         [all warps, then a CTA barrier]
//   2. per-point coefficient tables (81 threads)                                  [then a CTA barrier]
//   3. rows: r = panel^T x coefficients, again on the tensor cores, as NRESJOBS independent warp jobs that only
//      read the panels and the tables -> in the fused kernel they are jobs of the barrier-free main phase.

// (8 rows of a row-major panel) x (up to 8 state columns): out[row0 + lr][2 lk + {0,1}]
template <class BF>
__device__ __forceinline__ void rows_mma(const double* __restrict__ Pn, int ld, int row0, int K, BF bval, double& c0, double& c1) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lk = lane & 3;
  const double* ap = Pn + (row0 + lr) * ld;
  double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;  // two accumulator chains (even / odd k-steps)
  for (int k0 = 0; k0 < K; k0 += 8) {
    {
      const int kk = k0 + lk;
      const bool valid = kk < K;
      const double a = valid ? ap[kk] : 0.0;
      const double b = valid ? bval(kk, lr) : 0.0;
      dmma884(e0, e1, a, b);
    }
    if (k0 + 4 < K) {
      const int kk = k0 + 4 + lk;
      const bool valid = kk < K;
      const double a = valid ? ap[kk] : 0.0;
      const double b = valid ? bval(kk, lr) : 0.0;
      dmma884(o0, o1, a, b);
    }
  }
  c0 = e0 + o0;
  c1 = e1 + o1;
}

// stages 1 and 2.  Enter after the cell_prep barrier; ends WITHOUT a barrier (the caller synchronises before stage 3).
// T != null: also the table alpha d_d u_c of the Newton block.
template <int CONV, bool ZU>
__device__ __forceinline__ void residual_points(const CellCtx& cx, int64_t cell, const KParams& P, double* __restrict__ T) {
  double* sm = cx.sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lk = lane & 3;
  const bool solid = P.cell_solid != nullptr && P.cell_solid[cell] != 0;
  const double sig_c = solid ? P.cell_sigma[cell] : P.sigma;
  const double* U = sm + S_U;
  double* Rp = sm + S_ST;
  // ---- stage 1: 11 tiles grad u | 11 tiles j | 4 tiles (div j, phi, p)
  for (int job = warp; job < 26; job += NT / 32) {
    double c0, c1;
    if (job < 11) {
      rows_mma(sm + S_G, LDN, 8 * job, 27, [&](int k, int n) { return n < 3 ? U[n * 27 + k] : 0.0; }, c0, c1);
      const int m = 8 * job + lr;
      if (m < 81) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const int c = 2 * lk + r;
          if (c < 3) {
            const double v = r ? c1 : c0;
            Rp[RP_GQ + m * 3 + c] = v;
            if (T) T[(m / 3) * LDT + (m % 3) * 3 + c] = P.alpha * v / sm[S_SW + m / 3];
          }
        }
      }
    } else if (job < 22) {
      const int t = job - 11;
      rows_mma(sm + S_PSI, NJ, 8 * t, NJ, [&](int k, int n) { return n == 0 ? U[OFF_J + k] : 0.0; }, c0, c1);
      if (lk == 0) Rp[RP_JQ + 8 * t + lr] = c0;
    } else {
      const int t = job - 22;
      rows_mma(sm + S_DIV, NJ, 8 * t, NJ, [&](int k, int n) { return n == 0 ? U[OFF_J + k] : 0.0; }, c0, c1);
      if (lk == 0) Rp[RP_DJ + 8 * t + lr] = c0;
      rows_mma(sm + S_CHI, NF, 8 * t, NF, [&](int k, int n) { return n == 0 ? U[OFF_F + k] : 0.0; }, c0, c1);
      if (lk == 0) Rp[RP_FQ + 8 * t + lr] = c0;
      rows_mma(sm + S_PP, NP, 8 * t, NP, [&](int k, int n) { return n == 0 ? U[OFF_P + k] : 0.0; }, c0, c1);
      if (lk == 0) Rp[RP_PQ + 8 * t + lr] = c0;
    }
  }
  __syncthreads();
  if (ZU) {
    // Pi_p(div u) = pi . M_p^-1 (pi, div u): 4x4 mass matrix, right-hand side, solve, evaluate
    if (tid < 16) {
      const int k = tid >> 2, l = tid & 3;
      double s = 0.0;
      for (int q = 0; q < NQ; q++) s = fma(sm[S_PP + q * 4 + k], sm[S_PP + q * 4 + l], s);
      sm[S_MI + tid] = s;
    }
    if (tid >= 32 && tid < 36) {
      const int k = tid - 32;
      double s = 0.0;
      for (int q = 0; q < NQ; q++)
        s = fma(sm[S_PP + q * 4 + k], Rp[RP_GQ + q * 9 + 0] + Rp[RP_GQ + q * 9 + 4] + Rp[RP_GQ + q * 9 + 8], s);
      Rp[RP_RH + k] = s;
    }
    __syncthreads();
    if (tid == 0) {
      double a[4][5];
      for (int i = 0; i < 4; i++) {
        for (int j = 0; j < 4; j++) a[i][j] = sm[S_MI + i * 4 + j];
        a[i][4] = Rp[RP_RH + i];
      }
      for (int p = 0; p < 4; p++) {
        const double ip = 1.0 / a[p][p];
        for (int j = 0; j < 5; j++) a[p][j] *= ip;
        for (int i = 0; i < 4; i++)
          if (i != p) {
            const double f = a[i][p];
            for (int j = 0; j < 5; j++) a[i][j] -= f * a[p][j];
          }
      }
      for (int i = 0; i < 4; i++) Rp[RP_RH + i] = a[i][4];
    }
    __syncthreads();
    if (tid < NQ) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) s = fma(sm[S_PP + tid * 4 + k], Rp[RP_RH + k], s);
      Rp[RP_PR + tid] = s;
    }
    __syncthreads();
  }
  // ---- stage 2: thread (q, c)
  if (tid < 81) {
    const int q = tid / 3, c = tid - q * 3;
    const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
    const double sw = sm[S_SW + q];
    const double* Gq = Rp + RP_GQ + q * 9;
    const double* jq = Rp + RP_JQ + q * 3;
    const double* uq = sm + S_UQ + q * 3;  // unweighted
    const double jxB = jq[c1] * P.B[c2] - jq[c2] * P.B[c1];
    const double uxB = sw * (uq[c1] * P.B[c2] - uq[c2] * P.B[c1]);
    double fu = -P.gamma * jxB - sw * P.f[c];
    if (CONV > 0) {
      // sqrt(w) (u . grad) u_c = sum_d u_d (sqrt(w) d_d u_c)
      double cv = 0.0;
#pragma unroll
      for (int d = 0; d < 3; d++) cv = fma(uq[d], Gq[d * 3 + c], cv);
      fu = fma(P.alpha, cv, fu);
    }
    sm[RC_FU + q * 3 + c] = fu;
    sm[RC_FJ + q * 3 + c] = jq[c] - sig_c * uxB - sw * P.g[c];
    double diag = -Rp[RP_PQ + q];
    if (ZU) diag = fma(P.zeta_u, Rp[RP_PR + q], diag);
#pragma unroll
    for (int d = 0; d < 3; d++) sm[RC_GU + (q * 3 + d) * 3 + c] = P.beta * Gq[d * 3 + c] + (d == c ? diag : 0.0);
    if (c == 0) {
      sm[RC_DV + q] = Gq[0] + Gq[4] + Gq[8];
      sm[RC_DJ + q] = P.zeta_j * Rp[RP_DJ + q] - sig_c * Rp[RP_FQ + q];
      sm[RC_JD + q] = Rp[RP_DJ + q];
    }
  }
}

// stage 3, one warp job: job 0..3 r_u node tiles | 4..8 r_j tiles | 9 r_p | 10 r_phi
__device__ __forceinline__ void residual_rows_job(const CellCtx& cx, int job, bool solid, int64_t nrows, double* __restrict__ r) {
  const double* sm = cx.sm;
  const int lane = threadIdx.x & 31, lr = lane >> 2, lk = lane & 3;
  double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;
  auto run = [&](const double* Pn, int ld, int col, int K, auto bval) {
    for (int k0 = 0; k0 < K; k0 += 8) {
      {
        const int kk = k0 + lk;
        const bool valid = kk < K;
        const double a = valid ? Pn[kk * ld + col] : 0.0;
        const double b = valid ? bval(kk, lr) : 0.0;
        dmma884(e0, e1, a, b);
      }
      if (k0 + 4 < K) {
        const int kk = k0 + 4 + lk;
        const bool valid = kk < K;
        const double a = valid ? Pn[kk * ld + col] : 0.0;
        const double b = valid ? bval(kk, lr) : 0.0;
        dmma884(o0, o1, a, b);
      }
    }
  };
  auto add = [&](int li, double v) {
    const int32_t g = cx.gid[li];
    if (g >= 0 && g < nrows) atomicAdd(&r[g], v);
  };
  if (job < 4) {
    if (solid) return;
    // r_u[(c,a)] = sum_{(q,d)} G'[(q,d)][a] GU[(q,d)][c] + sum_q N'[q][a] FU[q][c]   (G' and N' are contiguous: K = 108)
    const int a = 8 * job + lr;
    run(sm + S_G, LDN, a, 108, [&](int k, int n) { return n < 3 ? sm[RC_GU + k * 3 + n] : 0.0; });
    if (a < 27) {
      const int c = 2 * lk;
      if (c < 3) add(c * 27 + a, e0 + o0);
      if (c + 1 < 3) add((c + 1) * 27 + a, e1 + o1);
    }
  } else if (job < 9) {
    // r_j[m] = sum_{(q,i)} Psi'[(q,i)][m] FJ[(q,i)] + sum_q Div'[q][m] DJ[q]        (Psi' and Div' are contiguous)
    const int m = 8 * (job - 4) + lr;
    run(sm + S_PSI, NJ, m, 108, [&](int k, int n) { return n == 0 ? sm[RC_FJ + k] : 0.0; });
    if (m < NJ && lk == 0) add(OFF_J + m, e0 + o0);
  } else if (job == 9) {
    if (solid) return;
    // r_p[k] = -sum_q pi_k div u
    run(sm + S_PP, NP, lr & 3, NQ, [&](int k, int n) { return n == 0 ? -sm[RC_DV + k] : 0.0; });
    if (lr < NP && lk == 0) add(OFF_P + lr, e0 + o0);
  } else {
    // r_phi[l] = -/+ sum_q chi_l div j   (res_solid_h1_hdiv has the opposite sign)
    run(sm + S_CHI, NF, lr, NQ, [&](int k, int n) { return n == 0 ? sm[RC_JD + k] : 0.0; });
    if (lk == 0) add(OFF_F + lr, (solid ? 1.0 : -1.0) * (e0 + o0));
  }
}

// ---------------------------------------------------------------------------------------------
// Sweep of a staged tile: entry e = k*32 + lane of the job's map range <-> staged value Sw[(e / NCOL) * LD + e % NCOL],
// local row rowof(e / NCOL).  A warp instruction therefore covers 32 consecutive entries of a destination-ordered
// (row-major, sorted columns) range: full-sector stores / REDs wherever the nnz are contiguous.
// Map codes as signed 16-bit values: >= 0: RED.ADD at that row-relative position; < -1: bit 15 = exclusive nnz, plain
// store at position (code & 0x7FFF); -1 (MAP_SKIP): dropped.
template <int KMAX>
__device__ __forceinline__ void load_codes(int (&code)[KMAX], const uint16_t* __restrict__ cmap, int n) {
  const short* p = reinterpret_cast<const short*>(cmap) + (threadIdx.x & 31);
#pragma unroll
  for (int k = 0; k < KMAX; k++) code[k] = k * 32 < n ? (int)__ldg(p + k * 32) : -1;  // ranges are padded to 32
}

__device__ __forceinline__ void scatter_pred(double* p, double v, int code) {
  asm volatile(
      "{\n\t.reg .pred pr, ps;\n\t"
      "setp.ge.s32 pr, %2, 0;\n\t"
      "setp.lt.s32 ps, %2, -1;\n\t"
      "@ps st.global.f64 [%0], %1;\n\t"
      "@pr red.global.add.f64 [%0], %1;\n\t}"
      ::"l"(p), "d"(v), "r"(code)
      : "memory");
}

// row[] holds the byte address of the first nnz of each local row (0 for dropped rows: their codes are all SKIP)
template <int NCOL, int LD, int KMAX, class RowOf>
__device__ __forceinline__ void sweep(const int (&code)[KMAX], int n, const double* __restrict__ Sw,
                                      const long long* __restrict__ row, RowOf rowof) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < KMAX; k++) {
    if (k * 32 >= n) break;
    int r, col;
    if (NCOL >= 32) {
      r = (k * 32) / NCOL;
      col = (k * 32) % NCOL + lane;
      if (col >= NCOL) { col -= NCOL; r++; }
    } else {
      const int e = k * 32 + lane;
      r = e / NCOL;      // NCOL is 4 or 8
      col = e % NCOL;
    }
    const double v = Sw[r * LD + col];  // lanes of the padding (code SKIP) read at most 31 entries past the range: callers
                                        // size LD so that this stays inside the warp's own tile
    double* p = reinterpret_cast<double*>(row[rowof(r)]) + (code[k] & 0x7FFF);
    scatter_pred(p, v, code[k]);
  }
}

__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// =============================================================================================
// Jacobian kernel.  CONV: 0 none, 1 picard, 2 newton.  ZU: zeta_u != 0 (rank-4 update zeta_u D^T M_p^-1 D of the uu
// block, folded into the tensor-core products as one extra k-step).  RES: also assemble the residual at the same state
// (residual_and_jacobian!), sharing the cell preparation.
constexpr int NJOBS = 15;

template <int CONV, bool ZU, bool RES>
__global__ void __launch_bounds__(NT, 2)
jacobian_kernel(int64_t ncells, int64_t nrows, CellArgs A, const double* __restrict__ x,
                const uint16_t* __restrict__ map, double* __restrict__ nz,
                double* __restrict__ rvec, KParams P) {
  extern __shared__ __align__(16) double smem[];
  CellCtx cx = make_ctx(smem);
  cx.clk_on = (P.dbg & 16) != 0;
  if (cx.clk_on && threadIdx.x == 0) {
    for (int i = 0; i < 15; i++) cx.clk[i] = 0;
    cx.clk[15] = clock64();
  }
  double* sm = smem;
  int* jobctr = reinterpret_cast<int*>(cx.perm + PERM_STRIDE);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lk = lane & 3;
  double* Sw = sm + S_ST + warp * SWN;
  prep_init(cx, A);

  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    __syncthreads();  // every warp is done with the previous cell's panels
    MHD_MARK(cx, 0);
    {
      // pull the NEXT cell's scatter map (30 KB), dof ids and vertex ids into L2 while this cell is being processed
      const int64_t nxt = cell + gridDim.x;
      if (nxt < ncells && !(P.dbg & 32)) {
        const char* m0 = reinterpret_cast<const char*>(map + nxt * NENT_PAD);
        if (tid * 128 < NENT_PAD * 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(m0 + tid * 128));
        if (tid >= 16 && tid < 25) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(A.rowstart + nxt * NLOC) + (tid - 16) * 128));
        if (tid >= 8 && tid < 13) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(A.gids + nxt * NLOC) + (tid - 8) * 128));
        if (tid == 5) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.cell_nodes + nxt * 8));
        if (tid == 6) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.perm + nxt * PERM_STRIDE));
      }
    }
    const int64_t nxt_cell = cell + gridDim.x < ncells ? cell + gridDim.x : -1;
    if (!(P.dbg & 4) || cell == blockIdx.x)
    cell_prep<(RES ? 2 : (CONV > 0 ? 1 : 0))>(cx, cell, nxt_cell, A, x, nz);
    if (RES) residual_points<(CONV > 0 ? 1 : 0), ZU>(cx, cell, P, CONV == 2 ? sm + S_T : (double*)nullptr);  // also T = alpha d_d u_c
    else if (CONV == 2) velocity_gradient_mma(sm + S_G, sm + S_U, sm + S_SW, nullptr, sm + S_T, LDT, P.alpha);
    if (CONV > 0) {
      // UG[q][b] = sqrt(w) u_q . grad N_b
      for (int idx = tid; idx < NQ * 27; idx += NT) {
        const int q = idx / 27, b = idx - q * 27;
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) s = fma(sm[S_UQ + q * 3 + i], sm[S_G + (q * 3 + i) * LDN + b], s);
        sm[S_UG + q * LDN + b] = s;
      }
    }
    if (tid < 108) sm[S_SC + tid] = tid < 81 ? 1.0 : P.zeta_j;
    if (tid == 0) *jobctr = 0;
    if (ZU) {
      // D[(c,a)][k] (3 warps), pressure mass matrix (16 threads of another warp), then E = zeta_u M_p^-1 D
      if (warp < 3) {
        double acc[4][1][2];
        int ca[4];
#pragma unroll
        for (int i = 0; i < 4; i++) ca[i] = 8 * i + lr;
        const int cb[1] = {lr};
        warp_mma_cols<4, 1, false>(sm + S_G + warp * LDN, 3 * LDN, ca, sm + S_PP, 4, cb, NQ, nullptr, 0, acc);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int a = 8 * i + lr;
          if (a < 27 && lk < 2) st2(sm + S_D + (warp * 27 + a) * 4 + 2 * lk, acc[i][0][0], acc[i][0][1]);
        }
      } else if (warp == 3 && lane < 16) {
        const int k = lane >> 2, l = lane & 3;
        double s = 0.0;
        for (int q = 0; q < NQ; q++) s = fma(sm[S_PP + q * 4 + k], sm[S_PP + q * 4 + l], s);
        sm[S_MI + lane] = s;
      }
      __syncthreads();
      if (tid == 0) {
        // in-place inverse of the SPD 4x4 mass matrix (Gauss-Jordan, no pivoting)
        double a[4][8];
        for (int i = 0; i < 4; i++)
          for (int j = 0; j < 4; j++) {
            a[i][j] = sm[S_MI + i * 4 + j];
            a[i][4 + j] = i == j ? 1.0 : 0.0;
          }
        for (int p = 0; p < 4; p++) {
          const double ip = 1.0 / a[p][p];
          for (int j = 0; j < 8; j++) a[p][j] *= ip;
          for (int i = 0; i < 4; i++)
            if (i != p) {
              const double f = a[i][p];
              for (int j = 0; j < 8; j++) a[i][j] -= f * a[p][j];
            }
        }
        for (int i = 0; i < 4; i++)
          for (int j = 0; j < 4; j++) sm[S_MI + i * 4 + j] = a[i][4 + j];
      }
      __syncthreads();
      for (int idx = tid; idx < 324; idx += NT) {
        const int k = idx / 81, i = idx - k * 81;
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < 4; l++) s = fma(sm[S_MI + k * 4 + l], sm[S_D + i * 4 + l], s);
        sm[S_E + k * 81 + i] = P.zeta_u * s;
      }
    }
    __syncthreads();
    MHD_MARK(cx, 5);

    if (P.dbg & 2) continue;
    const uint16_t* cmap = map + cell * NENT_PAD;
    const long long* row = cx.row;
    // solid cells (jac_solid_h1_hdiv, weakforms.jl:327-338): own conductivity, +phi div j instead of -div j phi; their
    // u/p dofs are absent, so only the jj and j-phi jobs run
    const bool solid = P.cell_solid != nullptr && P.cell_solid[cell] != 0;
    const double sig_c = solid ? P.cell_sigma[cell] : P.sigma;
    const double fj_sign = solid ? 1.0 : -1.0;

    // ------------------------------------------------------------------ uu job (every warp; none on solid cells)
    // warp w owns the node-slot tile (rows 8 mt .. 8 mt + 7) x (columns 16 np .. 16 np + 15) for all 9 component pairs:
    //   K[(c,a),(d,b)] = delta_cd (beta S_ab + alpha C_ab) + alpha sum_q N_a N_b (d_d u_c)(q) + zeta_u (D^T M^-1 D)
    if (!solid) {
      const int mt = warp >> 1, np = warp & 1;
      const int nrow = uu_nrow(mt), ncol = uu_ncol(np);
      const int ldu = np ? 41 : 49;  // 8 rows + the padded tail of the 33-column sweep stay inside the SWN-double tile
      // operand columns of this lane: node slots resolved through the permutation (slots >= 27: the zero pad column)
      const int ca[1] = {8 * mt + lr};
      const int cb[2] = {16 * np + lr, 16 * np + 8 + lr};
      double base[1][2][2];
      // S = G'^T G': on the diagonal of the node-tile grid the A tile is one of the two B tiles
      if (mt == 2 * np) warp_mma_cols<1, 2, false, 0>(sm + S_G, LDN, ca, sm + S_G, LDN, cb, 81, nullptr, 0, base);
      else if (mt == 2 * np + 1) warp_mma_cols<1, 2, false, 1>(sm + S_G, LDN, ca, sm + S_G, LDN, cb, 81, nullptr, 0, base);
      else warp_mma_cols<1, 2, false>(sm + S_G, LDN, ca, sm + S_G, LDN, cb, 81, nullptr, 0, base);
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int r = 0; r < 2; r++) base[0][j][r] *= P.beta;
      if (CONV > 0) {
        double cc[1][2][2];
        warp_mma_cols<1, 2, false>(sm + S_N, LDN, ca, sm + S_UG, LDN, cb, NQ, nullptr, 0, cc);
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
          for (int r = 0; r < 2; r++) base[0][j][r] = fma(P.alpha, cc[0][j][r], base[0][j][r]);
      }
      double fa[7], fb[7][2];
      if (CONV == 2) {
#pragma unroll
        for (int ks = 0; ks < 7; ks++) {
          const int kk = 4 * ks + lk;
          const bool valid = kk < NQ;
          const int kc = valid ? kk : 0;
          const double va = sm[S_N + kc * LDN + ca[0]];
          const double vb0 = sm[S_N + kc * LDN + cb[0]], vb1 = sm[S_N + kc * LDN + cb[1]];
          fa[ks] = valid ? va : 0.0;
          fb[ks][0] = valid ? vb0 : 0.0;
          fb[ks][1] = valid ? vb1 : 0.0;
        }
      }
#pragma unroll 1
      for (int c = 0; c < 3; c++) {
        int code[12];
        const int nuu = nrow * ncol;
        load_codes(code, cmap + uu_base(warp, c), nuu);
#pragma unroll
        for (int d = 0; d < 3; d++) {
          double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
          if (CONV == 2) {
            const int dc = d * 3 + c;
#pragma unroll
            for (int ks = 0; ks < 7; ks++) {
              const int kk = 4 * ks + lk;
              const double at = fa[ks] * sm[S_T + (kk < NQ ? kk : 0) * LDT + dc];
              dmma884(acc[0][0], acc[0][1], at, fb[ks][0]);
              dmma884(acc[1][0], acc[1][1], at, fb[ks][1]);
            }
          }
          if (ZU) {
            const int am = 8 * mt + lr;
            const double da = am < 27 ? sm[S_D + (c * 27 + am) * 4 + lk] : 0.0;
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const int bn = 16 * np + 8 * j + lr;
              const double eb = bn < 27 ? sm[S_E + lk * 81 + d * 27 + bn] : 0.0;
              dmma884(acc[j][0], acc[j][1], da, eb);
            }
          }
          if (CONV == 2 || ZU || d == c) {
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
              for (int r = 0; r < 2; r++) {
                double v = acc[j][r];
                if (d == c) v += base[0][j][r];
                const int bl = 8 * j + 2 * lk + r;
                if (3 * bl < ncol) Sw[lr * ldu + 3 * bl + d] = v;  // (node slots >= 27 would run into the next row)
              }
          }
        }
        __syncwarp();
        if (CONV == 2 || ZU) {
          if (np == 0) sweep<48, 49, 12>(code, nuu, Sw, row, [&](int r) { return c * 27 + 8 * mt + r; });
          else sweep<33, 41, 12>(code, nuu, Sw, row, [&](int r) { return c * 27 + 8 * mt + r; });
        } else {
          // none / picard: only the diagonal component pairs carry values (the others stay at the memset zero)
#pragma unroll
          for (int k = 0; k < 12; k++) {
            const int e = k * 32 + lane;
            const int col = np == 0 ? e % 48 : e % 33;
            if (col % 3 != c) code[k] = -1;
          }
          if (np == 0) sweep<48, 49, 12>(code, nuu, Sw, row, [&](int r) { return c * 27 + 8 * mt + r; });
          else sweep<33, 41, 12>(code, nuu, Sw, row, [&](int r) { return c * 27 + 8 * mt + r; });
        }
        __syncwarp();
      }
    }

    MHD_MARK(cx, 6);
    // ------------------------------------------------------------------ pooled jobs, drawn from a shared counter
    // jobs: 0 up/pu | 1..3 jj | 4..7 uj(s,0) | 8..11 uj(s,1) | 12 j-phi | 13,14 uj(4,h) | then the residual row jobs.
    // The counter hands them out in the order of JOB_ORDER (decreasing cost).
    for (;;) {
      int job = 0;
      if (lane == 0) job = atomicAdd(jobctr, 1);
      job = __shfl_sync(0xffffffffu, job, 0);
      if (job >= NJOBS + (RES ? NRESJOBS : 0)) break;
      if (job < NJOBS) job = (0x0EDCBA9876540321ull >> (4 * job)) & 15;  // 1 2 3(jj) 0(up/pu) 4..11 (uj) 12 (j-phi) 13 14
      if (job >= NJOBS) {
        residual_rows_job(cx, job - NJOBS, solid, nrows, rvec);
      } else if (job >= 1 && job <= 3) {
        // ---- jj: sum_{q,i} Psi Psi + zeta_j sum_q Div Div (Psi and Div panels are contiguous).  Jobs 1, 2: two 8-row
        //      strips each (a <2,5> register block: 7 operand loads per 10 MMAs), job 3: the last strip (4 rows)
        const int s0 = 2 * (job - 1);
        const bool two = job < 3;
        int code0[9], code1[9];
        load_codes(code0, cmap + jj_base(s0), (s0 < 4 ? 8 : 4) * NJ);
        load_codes(code1, cmap + jj_base(s0 + 1), two ? 8 * NJ : 0);
        double acc[2][5][2];
        const int ca[2] = {8 * s0 + lr, two ? 8 * s0 + 8 + lr : 8 * s0 + lr};
        int cb[5];
#pragma unroll
        for (int j = 0; j < 5; j++) cb[j] = 8 * j + lr;
        // the A tiles are B tiles 0,1 | 2,3 | 4 of the same panel: their fragments are reused
        if (P.zeta_j != 0.0) {
          if (job == 1) warp_mma_cols<2, 5, true, 0>(sm + S_PSI, NJ, ca, sm + S_PSI, NJ, cb, 108, sm + S_SC, 1, acc);
          else if (job == 2) warp_mma_cols<2, 5, true, 2>(sm + S_PSI, NJ, ca, sm + S_PSI, NJ, cb, 108, sm + S_SC, 1, acc);
          else warp_mma_cols<2, 5, true, 4>(sm + S_PSI, NJ, ca, sm + S_PSI, NJ, cb, 108, sm + S_SC, 1, acc);
        } else {
          if (job == 1) warp_mma_cols<2, 5, false, 0>(sm + S_PSI, NJ, ca, sm + S_PSI, NJ, cb, 81, nullptr, 0, acc);
          else if (job == 2) warp_mma_cols<2, 5, false, 2>(sm + S_PSI, NJ, ca, sm + S_PSI, NJ, cb, 81, nullptr, 0, acc);
          else warp_mma_cols<2, 5, false, 4>(sm + S_PSI, NJ, ca, sm + S_PSI, NJ, cb, 81, nullptr, 0, acc);
        }
#pragma unroll
        for (int j = 0; j < 5; j++) st2(Sw + lr * 40 + 8 * j + 2 * lk, acc[0][j][0], acc[0][j][1]);
        __syncwarp();
        sweep<NJ, 40, 9>(code0, (s0 < 4 ? 8 : 4) * NJ, Sw, row, [&](int r) { return OFF_J + 8 * s0 + r; });
        __syncwarp();
        if (two) {
#pragma unroll
          for (int j = 0; j < 5; j++) st2(Sw + lr * 40 + 8 * j + 2 * lk, acc[1][j][0], acc[1][j][1]);
          __syncwarp();
          sweep<NJ, 40, 9>(code1, 8 * NJ, Sw, row, [&](int r) { return OFF_J + 8 * s0 + 8 + r; });
          __syncwarp();
        }
      } else if (job == 12) {
        // ---- j-phi: -sigma JF[m][l] ; phi-j: -/+ JF[m][l],  JF = sum_q Div[q][m] Chi[q][l]
        int cjf[9], cfj[9];
        load_codes(cjf, cmap + SEC_JF, NJ * NF);
        load_codes(cfj, cmap + SEC_FJ, NF * NJ);
        double acc[5][1][2];
        int ca[5];
#pragma unroll
        for (int i = 0; i < 5; i++) ca[i] = 8 * i + lr;
        const int cb[1] = {lr};
        warp_mma_cols<5, 1, false>(sm + S_DIV, NJ, ca, sm + S_CHI, 8, cb, NQ, nullptr, 0, acc);
#pragma unroll
        for (int i = 0; i < 5; i++) st2(Sw + (8 * i + lr) * 8 + 2 * lk, -sig_c * acc[i][0][0], -sig_c * acc[i][0][1]);
        __syncwarp();
        sweep<NF, NF, 9>(cjf, NJ * NF, Sw, row, [&](int r) { return OFF_J + r; });
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 5; i++) {
          const int m = 8 * i + lr;
          if (m < NJ) {
            Sw[(2 * lk) * 38 + m] = fj_sign * acc[i][0][0];
            Sw[(2 * lk + 1) * 38 + m] = fj_sign * acc[i][0][1];
          }
        }
        __syncwarp();
        sweep<NJ, 38, 9>(cfj, NF * NJ, Sw, row, [&](int r) { return OFF_F + r; });
        __syncwarp();
      } else if (solid) {
        continue;
      } else if (job == 0) {
        // ---- up: K_up[(c,a)][k] = -D ; pu: K_pu[k][(b,d)] = -D,  D[(c,a)][k] = sum_q G[(q,c)][a] Pp[q][k]
        int cup[11], cpu[11];
        load_codes(cup, cmap + SEC_UP, NU * NP);
        load_codes(cpu, cmap + SEC_PU, NP * NU);
        double acc[3][4][1][2];
        int ca[4];
#pragma unroll
        for (int i = 0; i < 4; i++) ca[i] = 8 * i + lr;
        const int cb[1] = {lr};
#pragma unroll
        for (int c = 0; c < 3; c++)
          warp_mma_cols<4, 1, false>(sm + S_G + c * LDN, 3 * LDN, ca, sm + S_PP, 4, cb, NQ, nullptr, 0, acc[c]);
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int a = 8 * i + lr;
            if (a < 27 && lk < 2) st2(Sw + (c * 27 + a) * 4 + 2 * lk, -acc[c][i][0][0], -acc[c][i][0][1]);
          }
        __syncwarp();
        sweep<NP, NP, 11>(cup, NU * NP, Sw, row, [&](int r) { return r; });
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int a = 8 * i + lr;
            if (a < 27 && lk < 2) {
              Sw[(2 * lk) * 82 + 3 * a + c] = -acc[c][i][0][0];
              Sw[(2 * lk + 1) * 82 + 3 * a + c] = -acc[c][i][0][1];
            }
          }
        __syncwarp();
        sweep<NU, 82, 11>(cpu, NP * NU, Sw, row, [&](int r) { return OFF_P + r; });
        __syncwarp();
      } else {
        // ---- uj / ju job (s, h): j slots 8 s .. (8 | 4 of them), node slots 16 h .. (16 | 11 of them).
        //   Q_i[a][m] = sum_q N'[q][a] Psi'[(q,i)][m] for i = 0..2 (tensor cores), then in registers
        //   R_c = (psi x B)_c-weighted = B_{c+2} Q_{c+1} - B_{c+1} Q_{c+2};  K_uj[(c,a)][m] = -gamma R_c[a][m],
        //   K_ju[m][(b,d)] = +sigma R_d[b][m]
        const int s_ = job < 8 ? job - 4 : (job < 12 ? job - 8 : 4);
        const int h = job < 8 ? 0 : (job < 12 ? 1 : job - 13);
        const int nm = uj_nm(s_), na = uj_na(h);
        const uint16_t* cbase = cmap + uj_base(s_, h);
        int cuj[12], cju[12];
        const int nuj = 3 * na * nm;
        load_codes(cuj, cbase, nuj);
        load_codes(cju, cbase + uj_part(s_, h), nuj);
        double q[2][3][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int c = 0; c < 3; c++) q[i][c][0] = q[i][c][1] = 0.0;
        {
          const int ca0 = 16 * h + lr, ca1 = 16 * h + 8 + lr;
          const double* ap = sm + S_N;
          const double* bp = sm + S_PSI + 8 * s_ + lr;
#pragma unroll 2
          for (int k0 = 0; k0 < 28; k0 += 4) {
            const int kk = k0 + lk;
            const bool valid = kk < NQ;
            const int kc = valid ? kk : 0;
            double av[2], bv[3];
#pragma unroll
            for (int i = 0; i < 2; i++) {
              const double v = ap[kc * LDN + (i ? ca1 : ca0)];
              av[i] = valid ? v : 0.0;
            }
#pragma unroll
            for (int c = 0; c < 3; c++) {
              const double v = bp[(kc * 3 + c) * NJ];
              bv[c] = valid ? v : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
              for (int c = 0; c < 3; c++) dmma884(q[i][c][0], q[i][c][1], av[i], bv[c]);
          }
        }
        double R[2][3][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
#pragma unroll
            for (int r = 0; r < 2; r++) R[i][c][r] = P.B[c2] * q[i][c1][r] - P.B[c1] * q[i][c2][r];
          }
        // uj part: Sw[(c * na + al) * 8 + ml]
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const int al = 8 * i + lr;
            if (al < na) st2(Sw + (c * na + al) * 8 + 2 * lk, -P.gamma * R[i][c][0], -P.gamma * R[i][c][1]);
          }
        __syncwarp();
        if (h == 0) {
          if (nm == 8) sweep<8, 8, 12>(cuj, nuj, Sw, row, [&](int r) { return (r >> 4) * 27 + (r & 15); });
          else sweep<4, 8, 12>(cuj, nuj, Sw, row, [&](int r) { return (r >> 4) * 27 + (r & 15); });
        } else {
          if (nm == 8) sweep<8, 8, 12>(cuj, nuj, Sw, row, [&](int r) { return (r / 11) * 27 + 16 + r % 11; });
          else sweep<4, 8, 12>(cuj, nuj, Sw, row, [&](int r) { return (r / 11) * 27 + 16 + r % 11; });
        }
        __syncwarp();
        // ju part: Sw[ml * ldb + 3 * al + d]; the padded tail of the 33-column sweep (8 rows + 24 entries) must stay in the tile
        const int ldb = h == 0 ? 50 : 41;
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const int al = 8 * i + lr;
            if (al < na) {
              Sw[(2 * lk) * ldb + 3 * al + c] = sig_c * R[i][c][0];
              Sw[(2 * lk + 1) * ldb + 3 * al + c] = sig_c * R[i][c][1];
            }
          }
        __syncwarp();
        if (h == 0) sweep<48, 50, 12>(cju, nuj, Sw, row, [&](int r) { return OFF_J + 8 * s_ + r; });
        else sweep<33, 41, 12>(cju, nuj, Sw, row, [&](int r) { return OFF_J + 8 * s_ + r; });
        __syncwarp();
      }
    }
    MHD_MARK(cx, 7);
  }
  if (cx.clk_on && threadIdx.x == 0 && P.clk_out)
    for (int i = 0; i < 12; i++) atomicAdd(P.clk_out + i, (unsigned long long)cx.clk[i]);
}

template <int CONV, bool ZU>
__global__ void __launch_bounds__(NT, 2)
residual_kernel(int64_t ncells, int64_t nrows, CellArgs A, const double* __restrict__ x, double* __restrict__ r, KParams P) {
  extern __shared__ __align__(16) double smem[];
  CellCtx cx = make_ctx(smem);
  prep_init(cx, A);
  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    __syncthreads();
    cell_prep<2>(cx, cell, cell + gridDim.x < ncells ? cell + gridDim.x : -1, A, x, nullptr);
    residual_points<CONV, ZU>(cx, cell, P, nullptr);
    __syncthreads();
    const bool solid = P.cell_solid != nullptr && P.cell_solid[cell] != 0;
    for (int job = threadIdx.x >> 5; job < NRESJOBS; job += NT / 32) residual_rows_job(cx, job, solid, nrows, r);
  }
}

// ---------------------------------------------------------------------------------------------
static KParams make_kparams(const mhd_params_t& p) {
  KParams k;
  k.alpha = p.alpha; k.beta = p.beta; k.gamma = p.gamma; k.sigma = p.sigma; k.zeta_u = p.zeta_u; k.zeta_j = p.zeta_j;
  for (int i = 0; i < 3; i++) { k.B[i] = p.B[i]; k.f[i] = p.f[i]; k.g[i] = p.g[i]; }
  k.cell_solid = nullptr;
  k.cell_sigma = nullptr;
  k.dbg = 0;
  k.clk_out = nullptr;
  return k;
}

static CellArgs make_cell_args(const mhd_operator* op) {
  CellArgs a;
  a.tab = op->d_tables; a.ptab = op->d_ptab; a.coords = op->d_coords; a.cell_nodes = op->d_cell_nodes;
  a.gids = op->d_pgids; a.rowstart = op->d_rowstart; a.perm = op->d_perm; a.jsign = op->d_jsign; a.dirv = op->d_dir;
  return a;
}

static int sm_count() {
  return device_sm_count();
}

template <class Kern>
static int set_smem(Kern k) {
  MHD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return 0;
}

// Clearing nzval (1.18 GB on cfg2, 0.12 ms) and the residual runs on a side stream so that it overlaps the host->device
// copy of the state when the caller passes host buffers: begin_clear is called BEFORE that copy is enqueued, the kernel
// launch waits for it.  The side stream first waits for everything already enqueued on the library stream (an earlier
// SpMV / solve may still read the matrix).
static cudaStream_t s_clear_stream = nullptr;
static cudaEvent_t s_ev_in = nullptr, s_ev_done = nullptr;
static int s_clear_device = -1;

static bool s_clear_on_side = false;

// side = true: on the side stream (the caller is about to enqueue a host->device copy on the library stream and wants the
// clearing to overlap it); side = false: in line on the library stream -- no cross-stream hand-over, which costs ~10 us per
// assembly when there is nothing to overlap (device-resident callers)
int begin_clear(mhd_operator* op, double* d_r, bool matrix, bool side) {
  cudaStream_t st = g_stream;
  if (side) {
    if (s_clear_device != g_device) {  // (re)created per mhd_init
      MHD_CUDA(cudaStreamCreateWithFlags(&s_clear_stream, cudaStreamNonBlocking));
      MHD_CUDA(cudaEventCreateWithFlags(&s_ev_in, cudaEventDisableTiming));
      MHD_CUDA(cudaEventCreateWithFlags(&s_ev_done, cudaEventDisableTiming));
      s_clear_device = g_device;
    }
    MHD_CUDA(cudaEventRecord(s_ev_in, g_stream));
    MHD_CUDA(cudaStreamWaitEvent(s_clear_stream, s_ev_in, 0));
    st = s_clear_stream;
  }
  if (matrix && op->jac_version == 7) {
    MHD_TRY(v7_zero_shared(op, st, d_r));  // only the sectors that hold shared nnz (+ the residual, same launch)
  } else {
    if (matrix) MHD_CUDA(cudaMemsetAsync(op->d_nzval, 0, (size_t)op->nnz * sizeof(double), st));
    if (d_r) MHD_CUDA(cudaMemsetAsync(d_r, 0, (size_t)op->nrows * sizeof(double), st));
  }
  if (side) MHD_CUDA(cudaEventRecord(s_ev_done, s_clear_stream));
  s_clear_on_side = side;
  op->clear_pending = true;
  return 0;
}

int end_clear() {  // the compute stream waits for the clearing started by begin_clear (nothing to do when it ran in line)
  if (s_clear_on_side) MHD_CUDA(cudaStreamWaitEvent(g_stream, s_ev_done, 0));
  s_clear_on_side = false;
  return 0;
}

void assembly_finalize() {
  if (s_clear_device >= 0) {
    cudaStreamDestroy(s_clear_stream);
    cudaEventDestroy(s_ev_in);
    cudaEventDestroy(s_ev_done);
    s_clear_device = -1;
  }
}

// d_r != nullptr: fused residual + Jacobian (residual_and_jacobian!)
int launch_jacobian(mhd_operator* op, const double* d_x, double* d_r) {
  if (!op->clear_pending) MHD_TRY(begin_clear(op, d_r, true, false));
  op->clear_pending = false;
  MHD_TRY(end_clear());
  KParams P = make_kparams(op->prm);
  P.cell_solid = op->d_cell_solid;
  P.cell_sigma = op->d_cell_sigma;
  const int conv = op->prm.convection;
  const bool zu = op->prm.zeta_u != 0.0;
  static int ctas_per_sm = 0;  // resident CTAs per SM the kernel is built for (MHD_JAC_CTAS_PER_SM: A/B builds only)
  if (!ctas_per_sm) {
    const char* e = getenv("MHD_JAC_CTAS_PER_SM");
    ctas_per_sm = e ? atoi(e) : 2;
    if (ctas_per_sm < 1) ctas_per_sm = 2;
  }
  const int64_t grid64 = (int64_t)sm_count() * ctas_per_sm;
  const unsigned grid = (unsigned)(op->ncells < grid64 ? op->ncells : grid64);
#define JK(C, Z, R)                                                                                          \
  do {                                                                                                       \
    MHD_TRY(set_smem(jacobian_kernel<C, Z, R>));                                                             \
    jacobian_kernel<C, Z, R><<<grid, NT, SMEM_BYTES, g_stream>>>(op->ncells, op->nrows, make_cell_args(op), d_x,    \
        op->d_map, op->d_nzval, d_r, P);                                                                      \
  } while (0)
#define JKR(C, Z) do { if (d_r) JK(C, Z, true); else JK(C, Z, false); } while (0)
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("MHD_JAC_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  P.dbg = dbg;
  static unsigned long long* d_clk = nullptr;
  if ((dbg & 16) && !d_clk) MHD_CUDA(cudaMalloc((void**)&d_clk, 16 * sizeof(unsigned long long)));
  if (dbg & 16) MHD_CUDA(cudaMemsetAsync(d_clk, 0, 16 * sizeof(unsigned long long), g_stream));
  P.clk_out = d_clk;
  prof_begin(PROF_JAC);
  if (conv == 0 && !zu) JKR(0, false);
  else if (conv == 0 && zu) JKR(0, true);
  else if (conv == 1 && !zu) JKR(1, false);
  else if (conv == 1 && zu) JKR(1, true);
  else if (conv == 2 && !zu) JKR(2, false);
  else JKR(2, true);
#undef JKR
#undef JK
  prof_end(PROF_JAC);
  MHD_LAUNCH_CHECK();
  if (dbg & 16) {
    unsigned long long h[16];
    MHD_CUDA(cudaMemcpyAsync(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
    MHD_CUDA(cudaStreamSynchronize(g_stream));
    const double per = 1.0 / (double)op->ncells;
    fprintf(stderr, "[mhd phase clocks / cell] top %.0f  loads(bulk issue %.0f, own %.0f, wait %.0f)  geometry(math %.0f, wait %.0f)  panels %.0f  velgrad+UG %.0f  residual %.0f  uu %.0f  pool+tail %.0f\n",
            h[0] * per, h[10] * per, h[11] * per, h[1] * per, h[9] * per, h[2] * per, h[3] * per, h[4] * per, h[5] * per, h[6] * per, h[7] * per);
  }
  return 0;
}

int launch_residual(mhd_operator* op, const double* d_x, double* d_r) {
  MHD_CUDA(cudaMemsetAsync(d_r, 0, (size_t)op->nrows * sizeof(double), g_stream));
  KParams P = make_kparams(op->prm);
  P.cell_solid = op->d_cell_solid;
  P.cell_sigma = op->d_cell_sigma;
  const int conv = op->prm.convection;
  const bool zu = op->prm.zeta_u != 0.0;
  const int64_t grid64 = (int64_t)sm_count() * 2;
  const unsigned grid = (unsigned)(op->ncells < grid64 ? op->ncells : grid64);
#define RK(C, Z)                                                                                             \
  do {                                                                                                       \
    MHD_TRY(set_smem(residual_kernel<C, Z>));                                                                \
    residual_kernel<C, Z><<<grid, NT, SMEM_BYTES, g_stream>>>(op->ncells, op->nrows, make_cell_args(op), d_x, d_r, P); \
  } while (0)
  prof_begin(PROF_RES);
  if (conv == 0 && !zu) RK(0, false);
  else if (conv == 0 && zu) RK(0, true);
  else if (!zu) RK(1, false);   // picard and newton share the residual
  else RK(1, true);
#undef RK
  prof_end(PROF_RES);
  MHD_LAUNCH_CHECK();
  return 0;
}

}  // namespace mhd
