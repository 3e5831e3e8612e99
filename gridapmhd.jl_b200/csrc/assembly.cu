// FP64 sm_100a assembly kernels: cell-wise integration of the Jacobian and residual of the inductionless MHD
// H1-HDiv weak form, scattered into CSR / the residual vector through the precomputed map.
//
// Integrands: jac_fluid_h1_hdiv (src/weakforms.jl:283-312), res_fluid_h1_hdiv (src/weakforms.jl:255-281),
// conv (weakforms.jl:670), local projection (weakforms.jl:672-681).  Notation of SURVEY.md Appendix A.
//
// Design (one persistent CTA per SM slot, one cell at a time):
//   prep   : geometry Jacobians at the 27 Gauss points (registers -> smem), physical gradients of the Q2 basis,
//            Piola-mapped RT basis, all pre-scaled by sqrt(w_q |det J_q|) so every block is a plain product
//            sum_k A[k][m] B[k][n] of two shared-memory panels;
//   blocks : register-tiled FP64 panel products (4x4 per thread) into a shared staging buffer;
//   scatter: row-major sweep of each block section of the 16-bit scatter map (coalesced map reads, runs of
//            consecutive nnz), plain stores for single-contribution nnz, RED.ADD.F64 otherwise.
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
  MHD_TRY(dev_alloc(&op->d_tables, T_TOTAL));
  MHD_TRY(h2d(op->d_tables, h.data(), T_TOTAL));
  MHD_CUDA(cudaStreamSynchronize(g_stream));
  return 0;
}

struct KParams {
  double alpha, beta, gamma, sigma, zeta_u, zeta_j;
  double B[3], f[3], g[3];
  const uint8_t* cell_solid;  // null: no solid sub-domain
  const double* cell_sigma;
};

constexpr int NT = 256;  // threads per CTA

// ---- shared-memory plan (doubles)
constexpr int LDN = 28;                       // padded leading dimension of 27-wide panels
constexpr int S_G = 0;                        // [81][28]  sqrt(w) dN_a/dx_i, row = q*3+i
constexpr int S_UG = S_G + 81 * LDN;          // [27][28]  sqrt(w) (u_q . grad N_b)
constexpr int S_XB = S_G;                     // [27][108] sqrt(w) (psi_m x B)_c, row q, col c*36+m (aliases G,UG)
constexpr int S_N = S_UG + 27 * LDN;          // [27][28]  sqrt(w) N_a
constexpr int S_PSI = S_N + 27 * LDN;         // [81][36]  sqrt(w) Piola(psi_m)_i, row = q*3+i
constexpr int S_DIV = S_PSI + 81 * 36;        // [27][36]  sqrt(w) div psi_m   (contiguous after PSI)
constexpr int S_PP = S_DIV + 27 * 36;         // [27][4]
constexpr int S_CHI = S_PP + 27 * 4;          // [27][8]
constexpr int S_T = S_CHI + 27 * 8;           // [27][9]   (d_d u_c)(q), index d*3+c
constexpr int LDT = 10;                       // padded row of the velocity-gradient table
constexpr int S_ST = S_T + 27 * LDT + 2;      // staging
constexpr int ST_SIZE = 2070 + 1296;
constexpr int S_J = S_ST + ST_SIZE;           // [27][9]
constexpr int S_INV = S_J + 243;              // [27][9]
constexpr int S_DET = S_INV + 243;            // [27]
constexpr int S_SW = S_DET + 27;              // [27]
constexpr int S_X = S_SW + 27;                // [8][3]
constexpr int S_U = S_X + 24;                 // [129] local state
constexpr int S_SG = S_U + 130;               // [36] sign
constexpr int S_E = S_SG + 36;                // [4][81]
constexpr int S_MI = S_E + 324;               // [16]
constexpr int S_UQ = S_MI + 16;               // [27][3] u at q (unweighted)
constexpr int S_SC = S_UQ + 82;               // [108] scale vector for the jj product
constexpr int S_END = S_SC + 108;
constexpr int SMEM_BYTES = S_END * 8 + NLOC * 8 /*row starts*/ + 132 * 4 /*gids*/;

// staging sub-buffers of phase 1
constexpr int ST_D = 0;      // [81][4]  D[(c,a)][k] = sum_q w pi_k d_c N_a
constexpr int ST_S = 324;    // [27][27]
constexpr int ST_C = 1053;   // [27][27]
constexpr int ST_JF = 1782;  // [36][8]
constexpr int ST_JJ = 2070;  // [36][36]   (uj/ju results reuse [0,2916) after phase 0 is scattered)

// ---------------------------------------------------------------------------------------------
// register-tiled panel product:  C[batch][m][n] = sum_k A[k*lda + m] * (sc ? sc[k*scs] : 1) * B[k*ldb + n]
// A thread owns TM x TN outputs arranged as PAIRS of adjacent rows/columns: rows {2(tm + MT*p), +1}, p < TM/2 and
// columns {2(tn + NTL*p), +1}, p < TN/2.  Adjacent lanes therefore read adjacent 16-byte words of a panel row
// (LDS.128, bank-conflict free) while lanes that share tm read the same A word (broadcast).  Panels need
// lda >= TM*MT, ldb >= TN*NTL (even) and 16-byte aligned rows.  tile t -> thread (t + toff) % NT.
// store(batch, m, n, value) is called for in-range entries.
template <int M, int N, int TM, int TN, bool SCALE, class Store>
__device__ __forceinline__ void panel_product(int nbatch, const double* __restrict__ A, int lda, int a_bs,
                                              const double* __restrict__ B, int ldb, int b_bs, int K,
                                              const double* __restrict__ sc, int scs, int sc_bs, int toff,
                                              Store store) {
  static_assert(TM % 2 == 0 && TN % 2 == 0, "tiles are built from pairs");
  constexpr int MT = (M + TM - 1) / TM, NTL = (N + TN - 1) / TN;
  constexpr int PM = TM / 2, PN = TN / 2;
  const int ntiles = nbatch * MT * NTL;
  int t0 = (int)threadIdx.x - toff;
  t0 %= NT;
  if (t0 < 0) t0 += NT;
  for (int t = t0; t < ntiles; t += NT) {
    const int batch = t / (MT * NTL);
    const int r = t - batch * (MT * NTL);
    const int tm = r / NTL, tn = r - tm * NTL;
    const double* a = A + batch * a_bs + 2 * tm;
    const double* b = B + batch * b_bs + 2 * tn;
    const double* s = SCALE ? sc + batch * sc_bs : nullptr;
    double acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
      for (int j = 0; j < TN; j++) acc[i][j] = 0.0;
#pragma unroll 3
    for (int k = 0; k < K; k++) {
      double av[TM], bv[TN];
#pragma unroll
      for (int p = 0; p < PM; p++) {
        const double2 v = *reinterpret_cast<const double2*>(a + k * lda + 2 * MT * p);
        av[2 * p] = v.x;
        av[2 * p + 1] = v.y;
      }
#pragma unroll
      for (int p = 0; p < PN; p++) {
        const double2 v = *reinterpret_cast<const double2*>(b + k * ldb + 2 * NTL * p);
        bv[2 * p] = v.x;
        bv[2 * p + 1] = v.y;
      }
      if (SCALE) {
        const double sk = s[k * scs];
#pragma unroll
        for (int i = 0; i < TM; i++) av[i] *= sk;
      }
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
      for (int j = 0; j < TN; j++) {
        const int m = 2 * (tm + MT * (i / 2)) + (i & 1), n = 2 * (tn + NTL * (j / 2)) + (j & 1);
        if (m < M && n < N) store(batch, m, n, acc[i][j]);
      }
  }
}

// Unrolled scatter sweep: U map codes are loaded before any of them is used (memory-level parallelism; the sweep
// is otherwise bound by the latency of the 2-byte map loads).  getcode(e) reads the map code of sweep entry e,
// f(e, rowstart&, value&) supplies the nnz row offset and the value.
template <int U, class G, class F>
__device__ __forceinline__ void scatter_generic(double* __restrict__ nz, int count, G getcode, F f) {
  for (int base = 0; base < count; base += NT * U) {
    uint16_t c[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int e = base + k * NT + (int)threadIdx.x;
      c[k] = e < count ? getcode(e) : MAP_SKIP;
    }
#pragma unroll
    for (int k = 0; k < U; k++) {
      if (c[k] == MAP_SKIP) continue;
      const int e = base + k * NT + (int)threadIdx.x;
      long long rowstart;
      double v;
      f(e, rowstart, v);
      double* p = nz + rowstart + (c[k] & 0x7FFF);
      if (c[k] & MAP_EXCL) *p = v;
      else atomicAdd(p, v);
    }
  }
}

// Split form of the sweep for single-pass sections (count <= U * NT): the map codes are fetched into registers
// BEFORE the panel products that produce the values, so their global-load latency is hidden by the math.
template <int U>
struct Codes {
  uint16_t c[U];
};
template <int U>
__device__ __forceinline__ void load_codes(Codes<U>& cd, const uint16_t* __restrict__ codes, int count) {
#pragma unroll
  for (int k = 0; k < U; k++) {
    const int e = k * NT + (int)threadIdx.x;
    cd.c[k] = e < count ? __ldg(codes + e) : MAP_SKIP;
  }
}
template <int U, class F>
__device__ __forceinline__ void scatter_loaded(double* __restrict__ nz, const Codes<U>& cd, F f) {
#pragma unroll
  for (int k = 0; k < U; k++) {
    const uint16_t c = cd.c[k];
    if (c == MAP_SKIP) continue;
    const int e = k * NT + (int)threadIdx.x;
    long long rowstart;
    double v;
    f(e, rowstart, v);
    double* p = nz + rowstart + (c & 0x7FFF);
    if (c & MAP_EXCL) *p = v;
    else atomicAdd(p, v);
  }
}

// ---------------------------------------------------------------------------------------------
// FP64 tensor-core path: mma.sync.m8n8k4 (DMMA).  Same peak rate as the FP64 FMA pipe on B200, but an 8x8x4 product
// costs one operand word per lane and operand fragments are shared by all tiles of a warp's output block, which
// halves the shared-memory wavefronts of the panel products (the kernel is L1TEX-bound, not FP64- or HBM-bound).
#ifndef MHD_NO_MMA
constexpr bool USE_MMA = true;
#else
constexpr bool USE_MMA = false;
#endif

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

template <int MT, int NTL, bool SCALE, class Store>
__device__ __forceinline__ void warp_mma_product(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                                 int K, const double* __restrict__ sc, int scs, int m0, int n0, int M, int N,
                                                 Store store) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lk = lane & 3;
  double acc[MT][NTL][2];
  warp_mma_acc<MT, NTL, SCALE>(A, lda, B, ldb, K, sc, scs, m0, n0, acc);
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NTL; j++)
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const int m = m0 + 8 * i + lr, n = n0 + 8 * j + 2 * lk + r;
        if (m < M && n < N) store(m, n, acc[i][j][r]);
      }
}

// Velocity gradient at the 27 Gauss points through the tensor cores:
//   Gq[q][d*3+c] = sum_b G'[(q,d)][b] u_c,b   (= sqrt(w) d_d u_c),   T[q][d*3+c] = Gq / sqrt(w)
// an 81 x 3 x 27 product: A = G' read row-major ((q,d) rows, b contiguous: conflict-free fragment loads since
// LDN = 12 mod 16), B = the local velocity dofs.  11 row tiles dealt to the 8 warps.  Computed ONCE per cell and
// shared by the residual and the Newton block of the fused kernel.
__device__ __forceinline__ void velocity_gradient_mma(const double* __restrict__ G, const double* __restrict__ U,
                                                      const double* __restrict__ sw, double* __restrict__ Gq,
                                                      double* __restrict__ T, int ldt) {
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
          if (T) T[q * ldt + d * 3 + c] = v / sw[q];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
struct CellCtx {
  double* sm;
  long long* row;  // [129] nnz offset of each local row (-1: dropped)
  int32_t* gid;    // [129]
};

// loads + geometry + mapped bases.  NEED_STATE: 0 none, 1 u only, 2 all fields
template <int NEED_STATE>
__device__ __forceinline__ void cell_prep(const CellCtx& cx, int64_t cell, const double* __restrict__ tab,
                                          const double* __restrict__ coords, const int32_t* __restrict__ cell_nodes,
                                          const int32_t* __restrict__ gids, const int8_t* __restrict__ jsign,
                                          const double* __restrict__ dirv, const double* __restrict__ x) {
  double* sm = cx.sm;
  const int tid = threadIdx.x;
  for (int i = tid; i < NLOC; i += NT) {
    const int32_t g = gids[cell * NLOC + i];
    cx.gid[i] = g;
    if (NEED_STATE == 2 || (NEED_STATE == 1 && i < NU)) sm[S_U + i] = g >= 0 ? x[g] : dirv[-g - 1];
  }
  if (tid < 24) sm[S_X + tid] = coords[(int64_t)cell_nodes[cell * 8 + tid / 3] * 3 + tid % 3];
  if (tid >= 32 && tid < 32 + NJ) sm[S_SG + tid - 32] = (double)jsign[cell * NJ + tid - 32];
  __syncthreads();
  if (tid < NQ) {
    const int q = tid;
    double J[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) J[i][k] = 0.0;
    for (int v = 0; v < 8; v++) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double xv = sm[S_X + v * 3 + i];
#pragma unroll
        for (int k = 0; k < 3; k++) J[i][k] = fma(xv, tab[T_GG + (q * 8 + v) * 3 + k], J[i][k]);
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
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        sm[S_J + q * 9 + i * 3 + k] = J[i][k];
        sm[S_INV + q * 9 + k * 3 + i] = inv[k][i];
      }
    sm[S_DET + q] = det;
    sm[S_SW + q] = sqrt(tab[T_W + q] * fabs(det));
  }
  __syncthreads();
  // Q2 panels
  for (int idx = tid; idx < NQ * LDN; idx += NT) {
    const int q = idx / LDN, a = idx - q * LDN;
    if (a < 27) {
      const double sw = sm[S_SW + q];
      const double d0 = tab[T_DNU + (q * 27 + a) * 3 + 0], d1 = tab[T_DNU + (q * 27 + a) * 3 + 1],
                   d2 = tab[T_DNU + (q * 27 + a) * 3 + 2];
      const double* inv = sm + S_INV + q * 9;
#pragma unroll
      for (int i = 0; i < 3; i++)
        sm[S_G + (q * 3 + i) * LDN + a] = sw * (d0 * inv[0 * 3 + i] + d1 * inv[1 * 3 + i] + d2 * inv[2 * 3 + i]);
      sm[S_N + q * LDN + a] = sw * tab[T_NU + q * 27 + a];
    } else {
#pragma unroll
      for (int i = 0; i < 3; i++) sm[S_G + (q * 3 + i) * LDN + a] = 0.0;
      sm[S_N + q * LDN + a] = 0.0;
      sm[S_UG + q * LDN + a] = 0.0;
    }
  }
  // RT panels (contravariant Piola + sign flip)
  for (int idx = tid; idx < NQ * NJ; idx += NT) {
    const int q = idx / NJ, m = idx - q * NJ;
    const double s = sm[S_SG + m] * sm[S_SW + q] / sm[S_DET + q];
    const double p0 = tab[T_PSI + (q * 36 + m) * 3 + 0], p1 = tab[T_PSI + (q * 36 + m) * 3 + 1],
                 p2 = tab[T_PSI + (q * 36 + m) * 3 + 2];
    const double* J = sm + S_J + q * 9;
#pragma unroll
    for (int i = 0; i < 3; i++) sm[S_PSI + (q * 3 + i) * NJ + m] = s * (J[i * 3 + 0] * p0 + J[i * 3 + 1] * p1 + J[i * 3 + 2] * p2);
    sm[S_DIV + q * NJ + m] = s * tab[T_DPSI + q * 36 + m];
  }
  for (int idx = tid; idx < NQ * 4; idx += NT) sm[S_PP + idx] = sm[S_SW + idx / 4] * tab[T_PP + idx];
  for (int idx = tid; idx < NQ * 8; idx += NT) sm[S_CHI + idx] = sm[S_SW + idx / 8] * tab[T_CHI + idx];
  if (NEED_STATE >= 1) {
    // u at the quadrature points (unweighted)
    if (tid < 81) {
      const int q = tid / 3, i = tid - q * 3;
      double s = 0.0;
      for (int a = 0; a < 27; a++) s = fma(tab[T_NU + q * 27 + a], sm[S_U + i * 27 + a], s);
      sm[S_UQ + tid] = s;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void scatter_entry(double* __restrict__ nz, long long rowstart, uint16_t code, double v) {
  if (code == MAP_SKIP) return;
  double* p = nz + rowstart + (code & 0x7FFF);
  if (code & MAP_EXCL) *p = v;
  else atomicAdd(p, v);
}

template <int CONV, bool ZU>
__device__ __forceinline__ void cell_residual(const CellCtx& cx, int64_t cell, int64_t nrows, double* __restrict__ r,
                                              const KParams& P);
constexpr int RES_GQ = 81 + 243 + 81 + 27 * 4;  // offset of the velocity-gradient table inside the residual scratch

// =============================================================================================
// Jacobian kernel.  CONV: 0 none, 1 picard, 2 newton.  ZU: zeta_u != 0.
// RES: also assemble the residual at the same state (residual_and_jacobian!), sharing the cell preparation.
template <int CONV, bool ZU, bool RES>
__global__ void __launch_bounds__(NT, 2)
jacobian_kernel(int64_t ncells, int64_t nrows, const double* __restrict__ tab, const double* __restrict__ coords,
                const int32_t* __restrict__ cell_nodes, const int32_t* __restrict__ gids,
                const int8_t* __restrict__ jsign, const double* __restrict__ dirv, const double* __restrict__ x,
                const int64_t* __restrict__ rowptr, const uint16_t* __restrict__ map, double* __restrict__ nz,
                double* __restrict__ rvec, KParams P) {
  extern __shared__ __align__(16) double smem[];
  CellCtx cx;
  cx.sm = smem;
  cx.row = (long long*)(smem + S_END);
  cx.gid = (int32_t*)(cx.row + NLOC);
  double* sm = smem;
  const int tid = threadIdx.x;
  double* St = sm + S_ST;

  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    __syncthreads();  // previous cell's scatter done before its tables are overwritten
    {
      // pull the NEXT cell's scatter map (30 KB), dof ids and vertex ids into L2 while this cell is being processed:
      // every later load of them then pays L2 latency instead of DRAM latency inside the barrier-separated phases
      const int64_t nxt = cell + gridDim.x;
      if (nxt < ncells) {
        const char* m0 = reinterpret_cast<const char*>(map + nxt * NENT_PAD);
        if (tid * 128 < NENT_PAD * 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(m0 + tid * 128));
        if (tid < 5) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(gids + nxt * NLOC) + tid * 128));
        if (tid == 5) asm volatile("prefetch.global.L2 [%0];" ::"l"(cell_nodes + nxt * 8));
        if (tid == 6) asm volatile("prefetch.global.L2 [%0];" ::"l"(jsign + nxt * NJ));
      }
    }
    cell_prep<(RES ? 2 : (CONV > 0 ? 1 : 0))>(cx, cell, tab, coords, cell_nodes, gids, jsign, dirv, x);
    for (int i = tid; i < NLOC; i += NT) {
      const int32_t g = cx.gid[i];
      cx.row[i] = (g >= 0 && g < nrows) ? (long long)rowptr[g] : -1;
    }
    if (RES || CONV == 2) {
      // velocity gradient once per cell: Gq for the residual (its scratch area), T for the Newton block
      velocity_gradient_mma(sm + S_G, sm + S_U, sm + S_SW, RES ? sm + S_ST + RES_GQ : nullptr,
                            CONV == 2 ? sm + S_T : nullptr, LDT);
    }
    if (RES) {
      __syncthreads();
      cell_residual<(CONV > 0 ? 1 : 0), ZU>(cx, cell, nrows, rvec, P);
      __syncthreads();  // the residual's scratch lives in the staging area that phase 0 overwrites
    }
    if (CONV > 0) {
      // UG[q][b] = sqrt(w) u_q . grad N_b ;  T[q][d*3+c] = d_d u_c (unweighted)
      for (int idx = tid; idx < NQ * 27; idx += NT) {
        const int q = idx / 27, b = idx - q * 27;
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) s = fma(sm[S_UQ + q * 3 + i], sm[S_G + (q * 3 + i) * LDN + b], s);
        sm[S_UG + q * LDN + b] = s;
      }
    }
    if (tid < 108) sm[S_SC + tid] = tid < 81 ? 1.0 : P.zeta_j;
    __syncthreads();

    const uint16_t* cmap = map + cell * NENT_PAD;
    const long long* row = cx.row;
    // solid cells (jac_solid_h1_hdiv, weakforms.jl:327-338): own conductivity, +phi div j instead of -div j phi; their
    // u/p dofs are absent (map codes SKIP), so the u-related products are skipped
    const bool solid = P.cell_solid != nullptr && P.cell_solid[cell] != 0;
    const double sig_c = solid ? P.cell_sigma[cell] : P.sigma;
    const double fj_sign = solid ? 1.0 : -1.0;
    // map codes of the sections scattered right after phase 0, fetched now (latency hidden by the panel products)
    Codes<2> c_up, c_pu, c_jf, c_fj;
    Codes<6> c_jj;
    load_codes(c_up, cmap + SEC_UP, NU * NP);
    load_codes(c_pu, cmap + SEC_PU, NP * NU);
    load_codes(c_jf, cmap + SEC_JF, NJ * NF);
    load_codes(c_fj, cmap + SEC_FJ, NF * NJ);
    load_codes(c_jj, cmap + SEC_JJ, NJ * NJ);

    // ---------------- phase 0: D (up), j-phi, jj, S, C (+ pressure mass matrix), all independent panel products
    if (USE_MMA) {
      // 12 warp jobs, heaviest first, dealt round-robin to the 8 warps
      const int warp = tid >> 5;
      const bool zj = P.zeta_j != 0.0;
      for (int job = warp; job < 12; job += 8) {
        if (job < 2) {
          // S[a][b] = sum_{q,i} G[(q,i)][a] G[(q,i)][b], two column halves
          warp_mma_product<4, 2, false>(sm + S_G, LDN, sm + S_G, LDN, 81, nullptr, 0, 0, 16 * job, 27, 27,
                                        [&](int a, int b, double v) { St[ST_S + a * 27 + b] = v; });
        } else if (job < 7) {
          // jj = sum_{q,i} Psi Psi + zeta_j sum_q Div Div (Psi and Div panels are contiguous), one 8-column strip each
          const int n0 = 8 * (job - 2);
          if (zj)
            warp_mma_product<5, 1, true>(sm + S_PSI, NJ, sm + S_PSI, NJ, 108, sm + S_SC, 1, 0, n0, NJ, NJ,
                                         [&](int m, int n, double v) { St[ST_JJ + m * NJ + n] = v; });
          else
            warp_mma_product<5, 1, false>(sm + S_PSI, NJ, sm + S_PSI, NJ, 81, nullptr, 0, 0, n0, NJ, NJ,
                                          [&](int m, int n, double v) { St[ST_JJ + m * NJ + n] = v; });
        } else if (job == 7) {
          // C[a][b] = sum_q N[q][a] UG[q][b]
          if (CONV > 0)
            warp_mma_product<4, 4, false>(sm + S_N, LDN, sm + S_UG, LDN, NQ, nullptr, 0, 0, 0, 27, 27,
                                          [&](int a, int b, double v) { St[ST_C + a * 27 + b] = v; });
        } else if (job == 8) {
          // JF[m][l] = sum_q Div[q][m] Chi[q][l]
          warp_mma_product<5, 1, false>(sm + S_DIV, NJ, sm + S_CHI, 8, NQ, nullptr, 0, 0, 0, NJ, NF,
                                        [&](int m, int l, double v) { St[ST_JF + m * 8 + l] = v; });
        } else {
          // D[(c,a)][k] = sum_q G[(q,c)][a] Pp[q][k]
          const int c = job - 9;
          warp_mma_product<4, 1, false>(sm + S_G + c * LDN, 3 * LDN, sm + S_PP, 4, NQ, nullptr, 0, 0, 0, 27, NP,
                                        [&](int a, int k, double v) { St[ST_D + (c * 27 + a) * 4 + k] = v; });
        }
      }
    } else {
    // D[(c,a)][k] = sum_q G[(q,c)][a] Pp[q][k]   (batch = c)
    panel_product<27, 4, 4, 4, false>(3, sm + S_G, 3 * LDN, LDN, sm + S_PP, 4, 0, NQ, nullptr, 0, 0, 0,
                                      [&](int c, int a, int k, double v) { St[ST_D + (c * 27 + a) * 4 + k] = v; });
    // JF[m][l] = sum_q Div[q][m] Chi[q][l]
    panel_product<36, 8, 4, 4, false>(1, sm + S_DIV, NJ, 0, sm + S_CHI, 8, 0, NQ, nullptr, 0, 0, 21,
                                      [&](int, int m, int l, double v) { St[ST_JF + m * 8 + l] = v; });
    // jj = sum_{q,i} Psi Psi + zeta_j sum_q Div Div  (K = 81 or 108: Psi and Div panels are contiguous)
    if (P.zeta_j != 0.0)
      panel_product<36, 36, 2, 4, true>(1, sm + S_PSI, NJ, 0, sm + S_PSI, NJ, 0, 108, sm + S_SC, 1, 0, 39,
                                        [&](int, int m, int n, double v) { St[ST_JJ + m * NJ + n] = v; });
    else
      panel_product<36, 36, 2, 4, false>(1, sm + S_PSI, NJ, 0, sm + S_PSI, NJ, 0, 81, nullptr, 0, 0, 39,
                                         [&](int, int m, int n, double v) { St[ST_JJ + m * NJ + n] = v; });
    {
      // S[a][b] = sum_{q,i} G[(q,i)][a] G[(q,i)][b] ; C[a][b] = sum_q N[q][a] UG[q][b]
      panel_product<27, 27, 2, 4, false>(1, sm + S_G, LDN, 0, sm + S_G, LDN, 0, 81, nullptr, 0, 0, 39 + 162,
                                         [&](int, int a, int b, double v) { St[ST_S + a * 27 + b] = v; });
      if (CONV > 0)
        panel_product<27, 27, 4, 4, false>(1, sm + S_N, LDN, 0, sm + S_UG, LDN, 0, NQ, nullptr, 0, 0, 39 + 162 + 98,
                                           [&](int, int a, int b, double v) { St[ST_C + a * 27 + b] = v; });
    }
    }
    if (ZU && tid >= NT - 16) {
      const int kl = tid - (NT - 16), k = kl >> 2, l = kl & 3;
      double s = 0.0;
      for (int q = 0; q < NQ; q++) s = fma(sm[S_PP + q * 4 + k], sm[S_PP + q * 4 + l], s);
      sm[S_MI + kl] = s;
    }
    __syncthreads();
    if (ZU) {
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
        for (int l = 0; l < 4; l++) s = fma(sm[S_MI + k * 4 + l], St[ST_D + i * 4 + l], s);
        sm[S_E + k * 81 + i] = s;
      }
      __syncthreads();
    }

    // ---------------- uu
    if (solid) {
      // no u block on solid cells
    } else if (CONV == 2 && USE_MMA) {
      // Newton: K[(a,c),(b,d)] = delta_cd (beta S_ab + alpha C_ab) + alpha sum_q N_a N_b (d_d u_c)(q)  [+ zeta_u term].
      // Perfectly balanced over the 8 warps: warp w owns row tile w/2 and the column-tile pair w%2 of EVERY component
      // pair (c,d).  Its operand fragments of N' are loaded once (7 k-steps: 7 + 14 words per lane) and reused for
      // the 9 products N'^T diag(T_dc) N'; accumulators are scattered straight from registers (the 4 map codes of
      // a product are fetched before its MMAs), S and C come from the staging buffer.
      const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lk = lane & 3;
      const int mt = warp >> 1, np = warp & 1;
      double fa[7], fb[7][2];
#pragma unroll
      for (int ks = 0; ks < 7; ks++) {
        const int kk = 4 * ks + lk;
        const bool valid = kk < NQ;
        const int kc = valid ? kk : 0;
        const double va = sm[S_N + kc * LDN + 8 * mt + lr];
        const double vb0 = sm[S_N + kc * LDN + 16 * np + lr], vb1 = sm[S_N + kc * LDN + 16 * np + 8 + lr];
        fa[ks] = valid ? va : 0.0;
        fb[ks][0] = valid ? vb0 : 0.0;
        fb[ks][1] = valid ? vb1 : 0.0;
      }
      const int a = 8 * mt + lr;
      // map codes are fetched one product ahead (their DRAM latency is longer than the 14 MMAs of a product)
      auto load_uu_codes = [&](int dc, uint16_t (&code)[2][2]) {
        const int d = dc / 3, c = dc - d * 3;
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
          for (int r = 0; r < 2; r++) {
            const int b = 16 * np + 8 * j + 2 * lk + r;
            code[j][r] = (a < 27 && b < 27) ? __ldg(cmap + SEC_UU + (c * 27 + a) * NU + d * 27 + b) : MAP_SKIP;
          }
      };
      uint16_t code_next[2][2];
      load_uu_codes(0, code_next);
#pragma unroll 1
      for (int dc = 0; dc < 9; dc++) {
        const int d = dc / 3, c = dc - d * 3;
        uint16_t code[2][2];
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
          for (int r = 0; r < 2; r++) code[j][r] = code_next[j][r];
        if (dc < 8) load_uu_codes(dc + 1, code_next);
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int ks = 0; ks < 7; ks++) {
          const int kk = 4 * ks + lk;
          const double t = sm[S_T + (kk < NQ ? kk : 0) * LDT + dc];
          dmma884(acc[0][0], acc[0][1], fa[ks], fb[ks][0] * t);
          dmma884(acc[1][0], acc[1][1], fa[ks], fb[ks][1] * t);
        }
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
          for (int r = 0; r < 2; r++) {
            const uint16_t cd = code[j][r];
            if (cd == MAP_SKIP) continue;
            const int b = 16 * np + 8 * j + 2 * lk + r;
            const int li = c * 27 + a, lj = d * 27 + b;
            double v = P.alpha * acc[j][r];
            if (c == d) v += P.beta * St[ST_S + a * 27 + b] + P.alpha * St[ST_C + a * 27 + b];
            if (ZU) {
              double z = 0.0;
#pragma unroll
              for (int k = 0; k < 4; k++) z = fma(St[ST_D + li * 4 + k], sm[S_E + k * 81 + lj], z);
              v = fma(P.zeta_u, z, v);
            }
            double* pz = nz + row[li] + (cd & 0x7FFF);
            if (cd & MAP_EXCL) *pz = v;
            else atomicAdd(pz, v);
          }
      }
    } else     if (CONV == 2) {
      // Newton: a thread owns the 2x2 (a,b) node tile of all 9 component blocks:
      //   K[(a,c),(b,d)] = delta_cd (beta S_ab + alpha C_ab) + alpha sum_q N_a N_b (d_d u_c)(q)  [+ zeta_u term]
      // accumulated in registers and scattered straight from them (no staging, no extra barrier).
      if (tid < 196) {
        const int ta = tid / 14, tb = tid - ta * 14;
        const double* Ga = sm + S_G + 2 * ta;
        const double* Gb = sm + S_G + 2 * tb;
        const double* Na = sm + S_N + 2 * ta;
        const double* Nb = sm + S_N + 2 * tb;
        const double* Ub = sm + S_UG + 2 * tb;
        double sS[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, sC[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        double nw[9][2][2];
#pragma unroll
        for (int dc = 0; dc < 9; dc++) nw[dc][0][0] = nw[dc][0][1] = nw[dc][1][0] = nw[dc][1][1] = 0.0;
#pragma unroll 1
        for (int q = 0; q < NQ; q++) {
#pragma unroll
          for (int i = 0; i < 3; i++) {
            const double2 ga = *reinterpret_cast<const double2*>(Ga + (q * 3 + i) * LDN);
            const double2 gb = *reinterpret_cast<const double2*>(Gb + (q * 3 + i) * LDN);
            sS[0][0] = fma(ga.x, gb.x, sS[0][0]);
            sS[0][1] = fma(ga.x, gb.y, sS[0][1]);
            sS[1][0] = fma(ga.y, gb.x, sS[1][0]);
            sS[1][1] = fma(ga.y, gb.y, sS[1][1]);
          }
          const double2 na = *reinterpret_cast<const double2*>(Na + q * LDN);
          const double2 nb = *reinterpret_cast<const double2*>(Nb + q * LDN);
          const double2 ub = *reinterpret_cast<const double2*>(Ub + q * LDN);
          sC[0][0] = fma(na.x, ub.x, sC[0][0]);
          sC[0][1] = fma(na.x, ub.y, sC[0][1]);
          sC[1][0] = fma(na.y, ub.x, sC[1][0]);
          sC[1][1] = fma(na.y, ub.y, sC[1][1]);
          const double p00 = na.x * nb.x, p01 = na.x * nb.y, p10 = na.y * nb.x, p11 = na.y * nb.y;
          const double* Tq = sm + S_T + q * LDT;
#pragma unroll
          for (int dc = 0; dc < 9; dc++) {
            const double t = Tq[dc];
            nw[dc][0][0] = fma(t, p00, nw[dc][0][0]);
            nw[dc][0][1] = fma(t, p01, nw[dc][0][1]);
            nw[dc][1][0] = fma(t, p10, nw[dc][1][0]);
            nw[dc][1][1] = fma(t, p11, nw[dc][1][1]);
          }
        }
        // 36 entries per thread, swept one column component d at a time: 12 map codes are loaded first, then
        // the 12 values are scattered
#pragma unroll
        for (int d = 0; d < 3; d++) {
          uint16_t code[3][2][2];
#pragma unroll
          for (int c = 0; c < 3; c++)
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
              for (int j = 0; j < 2; j++) {
                const int a = 2 * ta + i, b = 2 * tb + j;
                code[c][i][j] = (a < 27 && b < 27) ? __ldg(cmap + SEC_UU + (c * 27 + a) * NU + d * 27 + b) : MAP_SKIP;
              }
#pragma unroll
          for (int c = 0; c < 3; c++)
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
              for (int j = 0; j < 2; j++) {
                const uint16_t cd = code[c][i][j];
                if (cd == MAP_SKIP) continue;
                const int li = c * 27 + 2 * ta + i, lj = d * 27 + 2 * tb + j;
                double v = P.alpha * nw[d * 3 + c][i][j];
                if (c == d) v += P.beta * sS[i][j] + P.alpha * sC[i][j];
                if (ZU) {
                  double z = 0.0;
#pragma unroll
                  for (int k = 0; k < 4; k++) z = fma(St[ST_D + li * 4 + k], sm[S_E + k * 81 + lj], z);
                  v = fma(P.zeta_u, z, v);
                }
                double* pz = nz + row[li] + (cd & 0x7FFF);
                if (cd & MAP_EXCL) *pz = v;
                else atomicAdd(pz, v);
              }
        }
      }
    } else if (!ZU) {
      // only the three diagonal component blocks carry values: beta S + alpha C
      scatter_generic<8>(nz, 3 * 729,
          [&](int idx) { const int c = idx / 729, ab = idx - c * 729, a = ab / 27, b = ab - a * 27;
                         return cmap[SEC_UU + (c * 27 + a) * NU + c * 27 + b]; },
          [&](int idx, long long& rs, double& v) {
            const int c = idx / 729, ab = idx - c * 729, a = ab / 27;
            rs = row[c * 27 + a];
            v = P.beta * St[ST_S + ab];
            if (CONV > 0) v = fma(P.alpha, St[ST_C + ab], v);
          });
    } else {
      scatter_generic<8>(nz, NU * NU, [&](int e) { return cmap[SEC_UU + e]; },
          [&](int e, long long& rs, double& v) {
            const int li = e / NU, lj = e - li * NU;
            const int c = li / 27, a = li - c * 27, d = lj / 27, b = lj - d * 27;
            double z = 0.0;
#pragma unroll
            for (int k = 0; k < 4; k++) z = fma(St[ST_D + li * 4 + k], sm[S_E + k * 81 + lj], z);
            z *= P.zeta_u;
            if (c == d) {
              z = fma(P.beta, St[ST_S + a * 27 + b], z);
              if (CONV > 0) z = fma(P.alpha, St[ST_C + a * 27 + b], z);
            }
            rs = row[li];
            v = z;
          });
    }
    // ---------------- up: K_up[(c,a)][k] = -D ; pu: K_pu[k][(d,b)] = -D
    scatter_loaded(nz, c_up, [&](int e, long long& rs, double& v) { rs = row[e >> 2]; v = -St[ST_D + e]; });
    scatter_loaded(nz, c_pu, [&](int e, long long& rs, double& v) { const int k = e / NU, i = e - k * NU; rs = row[OFF_P + k]; v = -St[ST_D + i * 4 + k]; });
    // ---------------- j-phi: -sigma JF[m][l] ; phi-j: -JF[m][l]
    scatter_loaded(nz, c_jf, [&](int e, long long& rs, double& v) { rs = row[OFF_J + (e >> 3)]; v = -sig_c * St[ST_JF + e]; });
    scatter_loaded(nz, c_fj, [&](int e, long long& rs, double& v) { const int l = e / NJ, m = e - l * NJ; rs = row[OFF_F + l]; v = fj_sign * St[ST_JF + m * 8 + l]; });
    // ---------------- jj
    scatter_loaded(nz, c_jj, [&](int e, long long& rs, double& v) { rs = row[OFF_J + e / NJ]; v = St[ST_JJ + e]; });
    // map codes of the uj / ju sections, fetched before the XB fill and the uj product
    if (solid) continue;  // CTA-uniform: no uj / ju blocks on solid cells (the loop-top barrier follows)
    Codes<12> c_uj, c_ju;
    load_codes(c_uj, cmap + SEC_UJ, NU * NJ);
    load_codes(c_ju, cmap + SEC_JU, NJ * NU);
    __syncthreads();

    // ---------------- uj / ju.  XB[q][c*36+m] = sqrt(w) (psi_m x B)_c (overwrites G, UG)
    for (int idx = tid; idx < NQ * NJ; idx += NT) {
      const int q = idx / NJ, m = idx - q * NJ;
      const double p0 = sm[S_PSI + (q * 3 + 0) * NJ + m], p1 = sm[S_PSI + (q * 3 + 1) * NJ + m],
                   p2 = sm[S_PSI + (q * 3 + 2) * NJ + m];
      sm[S_XB + q * 108 + 0 * 36 + m] = p1 * P.B[2] - p2 * P.B[1];
      sm[S_XB + q * 108 + 1 * 36 + m] = p2 * P.B[0] - p0 * P.B[2];
      sm[S_XB + q * 108 + 2 * 36 + m] = p0 * P.B[1] - p1 * P.B[0];
    }
    __syncthreads();
    // R[c][a][m] = sum_q N[q][a] XB[q][c*36+m]
    if (USE_MMA) {
      const int warp = tid >> 5;
      if (warp < 7)
        warp_mma_product<4, 2, false>(sm + S_N, LDN, sm + S_XB, 108, NQ, nullptr, 0, 0, 16 * warp, 27, 108,
                                      [&](int a, int n, double v) { const int c = n / NJ; St[(c * 27 + a) * NJ + (n - c * NJ)] = v; });
    } else {
      panel_product<27, 36, 4, 4, false>(3, sm + S_N, LDN, 0, sm + S_XB, 108, 36, NQ, nullptr, 0, 0, 0,
                                         [&](int c, int a, int m, double v) { St[(c * 27 + a) * NJ + m] = v; });
    }
    __syncthreads();
    // K_uj[(c,a)][m] = -gamma R ; K_ju[m][(d,b)] = +sigma R[d][b][m]
    scatter_loaded(nz, c_uj, [&](int e, long long& rs, double& v) { rs = row[e / NJ]; v = -P.gamma * St[e]; });
    scatter_loaded(nz, c_ju, [&](int e, long long& rs, double& v) { const int m = e / NU, i = e - m * NU; rs = row[OFF_J + m]; v = sig_c * St[i * NJ + m]; });
  }
}

// =============================================================================================
// Residual of one cell: res_fluid_h1_hdiv (src/weakforms.jl:255-281).  Needs cell_prep<2> (all panels + the full
// local state); uses the first ~820 doubles of the staging area and leaves the panels untouched.
template <int CONV, bool ZU>
__device__ __forceinline__ void cell_residual(const CellCtx& cx, int64_t cell, int64_t nrows, double* __restrict__ r,
                                              const KParams& P) {
  double* sm = cx.sm;
  const int tid = threadIdx.x;
  // res_solid_h1_hdiv (weakforms.jl:314-325) on solid cells: sigma of the cell, +phi-test * div j; u = 0 there
  const bool solid = P.cell_solid != nullptr && P.cell_solid[cell] != 0;
  const double sig_c = solid ? P.cell_sigma[cell] : P.sigma;
  const double f_sign = solid ? 1.0 : -1.0;
  // per-q coefficient tables in the staging area
  double* Fu = sm + S_ST;         // [27][3]  coefficient of N'[q][a] in r_u[(c,a)]
  double* Gu = Fu + 81;           // [27][9]  coefficient of G'[(q,d)][a], index d*3+c
  double* Fj = Gu + 243;          // [27][3]  coefficient of Psi'[(q,i)][m]
  double* Dj = Fj + 81;           // [27]     coefficient of Div'[q][m]
  double* Dv = Dj + 27;           // [27]     sqrt(w) div u
  double* Jd = Dv + 27;           // [27]     sqrt(w) div j
  double* Pq = Jd + 27;           // [27]     sqrt(w) p
  double* Gq = Pq + 27;           // [27][9]  sqrt(w) d_d u_c
  double* Pr = Gq + 243;          // [27]     sqrt(w) Pi_p(div u)
  double* Rh = Pr + 27;           // [4] rhs / coefficients of the projection
  {
    const double* U = sm + S_U;
    // Gq (velocity gradient at q) has been filled by velocity_gradient_mma; div u, p at q
    if (tid < NQ) {
      const int q = tid;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) s = fma(sm[S_PP + q * 4 + k], U[OFF_P + k], s);
      Pq[q] = s;
      double dj = 0.0;
      for (int m = 0; m < NJ; m++) dj = fma(sm[S_DIV + q * NJ + m], U[OFF_J + m], dj);
      Jd[q] = dj;
    }
    __syncthreads();
    if (tid < NQ) Dv[tid] = Gq[tid * 9 + 0] + Gq[tid * 9 + 4] + Gq[tid * 9 + 8];
    __syncthreads();
    if (ZU) {
      if (tid < 16) {
        const int k = tid >> 2, l = tid & 3;
        double s = 0.0;
        for (int q = 0; q < NQ; q++) s = fma(sm[S_PP + q * 4 + k], sm[S_PP + q * 4 + l], s);
        sm[S_MI + tid] = s;
      }
      if (tid >= 32 && tid < 36) {
        const int k = tid - 32;
        double s = 0.0;
        for (int q = 0; q < NQ; q++) s = fma(sm[S_PP + q * 4 + k], Dv[q], s);
        Rh[k] = s;
      }
      __syncthreads();
      if (tid == 0) {
        double a[4][5];
        for (int i = 0; i < 4; i++) {
          for (int j = 0; j < 4; j++) a[i][j] = sm[S_MI + i * 4 + j];
          a[i][4] = Rh[i];
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
        for (int i = 0; i < 4; i++) Rh[i] = a[i][4];
      }
      __syncthreads();
      if (tid < NQ) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 4; k++) s = fma(sm[S_PP + tid * 4 + k], Rh[k], s);
        Pr[tid] = s;
      }
      __syncthreads();
    }
    if (tid < NQ) {
      const int q = tid;
      const double sw = sm[S_SW + q];
      // weighted u, j, phi at q
      double uq[3], jq[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        uq[i] = sm[S_UQ + q * 3 + i] * sw;
        double s = 0.0;
        for (int m = 0; m < NJ; m++) s = fma(sm[S_PSI + (q * 3 + i) * NJ + m], U[OFF_J + m], s);
        jq[i] = s;
      }
      double fq = 0.0;
#pragma unroll
      for (int l = 0; l < 8; l++) fq = fma(sm[S_CHI + q * 8 + l], U[OFF_F + l], fq);
      const double jxB[3] = {jq[1] * P.B[2] - jq[2] * P.B[1], jq[2] * P.B[0] - jq[0] * P.B[2], jq[0] * P.B[1] - jq[1] * P.B[0]};
      const double uxB[3] = {uq[1] * P.B[2] - uq[2] * P.B[1], uq[2] * P.B[0] - uq[0] * P.B[2], uq[0] * P.B[1] - uq[1] * P.B[0]};
#pragma unroll
      for (int c = 0; c < 3; c++) {
        double v = -P.gamma * jxB[c] - sw * P.f[c];
        if (CONV > 0) {
          // sqrt(w) (u . grad) u_c = sum_d u_d (sqrt(w) d_d u_c)
          double cv = 0.0;
#pragma unroll
          for (int d = 0; d < 3; d++) cv = fma(sm[S_UQ + q * 3 + d], Gq[q * 9 + d * 3 + c], cv);
          v = fma(P.alpha, cv, v);
        }
        Fu[q * 3 + c] = v;
        Fj[q * 3 + c] = jq[c] - sig_c * uxB[c] - sw * P.g[c];
      }
#pragma unroll
      for (int d = 0; d < 3; d++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
          double v = P.beta * Gq[q * 9 + d * 3 + c];
          if (d == c) {
            v -= Pq[q];
            if (ZU) v = fma(P.zeta_u, Pr[q], v);
          }
          Gu[q * 9 + d * 3 + c] = v;
        }
      Dj[q] = P.zeta_j * Jd[q] - sig_c * fq;
    }
    __syncthreads();
    // rows
    if (tid < NLOC) {
      const int i = tid;
      double s = 0.0;
      if (i < NU) {
        const int c = i / 27, a = i - c * 27;
        for (int q = 0; q < NQ; q++) {
          s = fma(sm[S_N + q * LDN + a], Fu[q * 3 + c], s);
#pragma unroll
          for (int d = 0; d < 3; d++) s = fma(sm[S_G + (q * 3 + d) * LDN + a], Gu[q * 9 + d * 3 + c], s);
        }
      } else if (i < OFF_J) {
        const int k = i - OFF_P;
        for (int q = 0; q < NQ; q++) s = fma(sm[S_PP + q * 4 + k], Dv[q], s);
        s = -s;
      } else if (i < OFF_F) {
        const int m = i - OFF_J;
        for (int q = 0; q < NQ; q++) {
#pragma unroll
          for (int d = 0; d < 3; d++) s = fma(sm[S_PSI + (q * 3 + d) * NJ + m], Fj[q * 3 + d], s);
          s = fma(sm[S_DIV + q * NJ + m], Dj[q], s);
        }
      } else {
        const int l = i - OFF_F;
        for (int q = 0; q < NQ; q++) s = fma(sm[S_CHI + q * 8 + l], Jd[q], s);
        s = f_sign * s;
      }
      const int32_t g = cx.gid[i];
      if (g >= 0 && g < nrows) atomicAdd(&r[g], s);
    }
  }
}

template <int CONV, bool ZU>
__global__ void __launch_bounds__(NT, 2)
residual_kernel(int64_t ncells, int64_t nrows, const double* __restrict__ tab, const double* __restrict__ coords,
                const int32_t* __restrict__ cell_nodes, const int32_t* __restrict__ gids,
                const int8_t* __restrict__ jsign, const double* __restrict__ dirv, const double* __restrict__ x,
                double* __restrict__ r, KParams P) {
  extern __shared__ __align__(16) double smem[];
  CellCtx cx;
  cx.sm = smem;
  cx.row = (long long*)(smem + S_END);
  cx.gid = (int32_t*)(cx.row + NLOC);
  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    __syncthreads();
    cell_prep<2>(cx, cell, tab, coords, cell_nodes, gids, jsign, dirv, x);
    velocity_gradient_mma(cx.sm + S_G, cx.sm + S_U, cx.sm + S_SW, cx.sm + S_ST + RES_GQ, nullptr, 0);
    __syncthreads();
    cell_residual<CONV, ZU>(cx, cell, nrows, r, P);
  }
}

// =============================================================================================
// Jacobian kernel, single-phase variant ("v4").  After the cell preparation there is ONE barrier-free phase: every
// warp runs a fixed list of tensor-core jobs and scatters each job's accumulators straight from registers (map
// codes are fetched before the job's MMAs).  No staging buffer, no barrier between products and scatter:
//   * uu job (every warp): warp w owns row tile w/2 and the column-tile pair w%2 of the 27x27 node space for ALL
//     9 component pairs: S = G'^T G' and C = N'^T UG' of that tile stay in registers and are added on the three
//     diagonal pairs; the Newton products N'^T diag(T_dc) N' reuse one set of N' fragments;
//   * pooled jobs: jj (5 column strips), uj/ju (7 column pairs, psi x B formed on the fly from the Psi panel),
//     j-phi/phi-j, up/pu (3 components), dealt statically: w0..w2: jj+D, w3: jj+JF, w4: jj+uj, w5..w7: 2 uj.
// The barrier-heavy staged kernel above remains for zeta_u != 0 (its rank-4 update needs D and M_p^-1 D first).
__device__ __forceinline__ void scatter_reg(double* __restrict__ nz, long long rowstart, uint16_t cd, double v) {
  if (cd == MAP_SKIP) return;
  double* p = nz + rowstart + (cd & 0x7FFF);
  if (cd & MAP_EXCL) *p = v;
  else atomicAdd(p, v);
}

template <int CONV, bool RES>
__global__ void __launch_bounds__(NT, 2)
jacobian_kernel_v4(int64_t ncells, int64_t nrows, const double* __restrict__ tab, const double* __restrict__ coords,
                   const int32_t* __restrict__ cell_nodes, const int32_t* __restrict__ gids,
                   const int8_t* __restrict__ jsign, const double* __restrict__ dirv, const double* __restrict__ x,
                   const int64_t* __restrict__ rowptr, const uint16_t* __restrict__ map, double* __restrict__ nz,
                   double* __restrict__ rvec, KParams P) {
  extern __shared__ __align__(16) double smem[];
  CellCtx cx;
  cx.sm = smem;
  cx.row = (long long*)(smem + S_END);
  cx.gid = (int32_t*)(cx.row + NLOC);
  double* sm = smem;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lk = lane & 3;

  for (int64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
    __syncthreads();  // every warp is done with the previous cell's panels
    {
      const int64_t nxt = cell + gridDim.x;  // next cell's map / ids into L2 (see jacobian_kernel)
      if (nxt < ncells) {
        const char* m0 = reinterpret_cast<const char*>(map + nxt * NENT_PAD);
        if (tid * 128 < NENT_PAD * 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(m0 + tid * 128));
        if (tid < 5) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(gids + nxt * NLOC) + tid * 128));
        if (tid == 5) asm volatile("prefetch.global.L2 [%0];" ::"l"(cell_nodes + nxt * 8));
        if (tid == 6) asm volatile("prefetch.global.L2 [%0];" ::"l"(jsign + nxt * NJ));
      }
    }
    cell_prep<(RES ? 2 : (CONV > 0 ? 1 : 0))>(cx, cell, tab, coords, cell_nodes, gids, jsign, dirv, x);
    for (int i = tid; i < NLOC; i += NT) {
      const int32_t g = cx.gid[i];
      cx.row[i] = (g >= 0 && g < nrows) ? (long long)rowptr[g] : -1;
    }
    if (RES || CONV == 2)
      velocity_gradient_mma(sm + S_G, sm + S_U, sm + S_SW, RES ? sm + S_ST + RES_GQ : nullptr, CONV == 2 ? sm + S_T : nullptr, LDT);
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
    if (RES) {
      __syncthreads();
      cell_residual<(CONV > 0 ? 1 : 0), false>(cx, cell, nrows, rvec, P);  // touches only its scratch area
    }
    __syncthreads();

    const uint16_t* cmap = map + cell * NENT_PAD;
    const long long* row = cx.row;
    const bool solid = P.cell_solid != nullptr && P.cell_solid[cell] != 0;
    const double sig_c = solid ? P.cell_sigma[cell] : P.sigma;
    const double fj_sign = solid ? 1.0 : -1.0;

    // ------------------------------------------------------------------ uu job (all warps; none on solid cells)
    if (!solid) {
      const int mt = warp >> 1, np = warp & 1;
      const int a = 8 * mt + lr;
      double base[1][2][2];
      warp_mma_acc<1, 2, false>(sm + S_G, LDN, sm + S_G, LDN, 81, nullptr, 0, 8 * mt, 16 * np, base);
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int r = 0; r < 2; r++) base[0][j][r] *= P.beta;
      if (CONV > 0) {
        double cc[1][2][2];
        warp_mma_acc<1, 2, false>(sm + S_N, LDN, sm + S_UG, LDN, NQ, nullptr, 0, 8 * mt, 16 * np, cc);
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
          for (int r = 0; r < 2; r++) base[0][j][r] = fma(P.alpha, cc[0][j][r], base[0][j][r]);
      }
      auto load_uu_codes = [&](int c, int d, uint16_t (&code)[2][2]) {
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
          for (int r = 0; r < 2; r++) {
            const int b = 16 * np + 8 * j + 2 * lk + r;
            code[j][r] = (a < 27 && b < 27) ? __ldg(cmap + SEC_UU + (c * 27 + a) * NU + d * 27 + b) : MAP_SKIP;
          }
      };
      if (CONV == 2) {
        double fa[7], fb[7][2];
#pragma unroll
        for (int ks = 0; ks < 7; ks++) {
          const int kk = 4 * ks + lk;
          const bool valid = kk < NQ;
          const int kc = valid ? kk : 0;
          const double va = sm[S_N + kc * LDN + 8 * mt + lr];
          const double vb0 = sm[S_N + kc * LDN + 16 * np + lr], vb1 = sm[S_N + kc * LDN + 16 * np + 8 + lr];
          fa[ks] = valid ? va : 0.0;
          fb[ks][0] = valid ? vb0 : 0.0;
          fb[ks][1] = valid ? vb1 : 0.0;
        }
        uint16_t code_next[2][2];
        load_uu_codes(0, 0, code_next);
#pragma unroll 1
        for (int dc = 0; dc < 9; dc++) {
          const int d = dc / 3, c = dc - d * 3;
          uint16_t code[2][2];
#pragma unroll
          for (int j = 0; j < 2; j++)
#pragma unroll
            for (int r = 0; r < 2; r++) code[j][r] = code_next[j][r];
          if (dc < 8) load_uu_codes((dc + 1) % 3, (dc + 1) / 3, code_next);
          double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
          for (int ks = 0; ks < 7; ks++) {
            const int kk = 4 * ks + lk;
            const double t = sm[S_T + (kk < NQ ? kk : 0) * LDT + dc];
            dmma884(acc[0][0], acc[0][1], fa[ks], fb[ks][0] * t);
            dmma884(acc[1][0], acc[1][1], fa[ks], fb[ks][1] * t);
          }
#pragma unroll
          for (int j = 0; j < 2; j++)
#pragma unroll
            for (int r = 0; r < 2; r++) {
              double v = P.alpha * acc[j][r];
              if (c == d) v += base[0][j][r];
              if (a < 27) scatter_reg(nz, row[c * 27 + a], code[j][r], v);
            }
        }
      } else {
        // none / picard: only the three diagonal component blocks carry values
#pragma unroll 1
        for (int c = 0; c < 3; c++) {
          uint16_t code[2][2];
          load_uu_codes(c, c, code);
#pragma unroll
          for (int j = 0; j < 2; j++)
#pragma unroll
            for (int r = 0; r < 2; r++)
              if (a < 27) scatter_reg(nz, row[c * 27 + a], code[j][r], base[0][j][r]);
        }
      }
    }

    // ------------------------------------------------------------------ pooled jobs
    // jj strip s: columns 8s..8s+7
    auto job_jj = [&](int s_) {
      double acc[5][1][2];
      uint16_t code[5][2];
#pragma unroll
      for (int i = 0; i < 5; i++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const int m = 8 * i + lr, n = 8 * s_ + 2 * lk + r;
          code[i][r] = (m < NJ && n < NJ) ? __ldg(cmap + SEC_JJ + m * NJ + n) : MAP_SKIP;
        }
      if (P.zeta_j != 0.0) warp_mma_acc<5, 1, true>(sm + S_PSI, NJ, sm + S_PSI, NJ, 108, sm + S_SC, 1, 0, 8 * s_, acc);
      else warp_mma_acc<5, 1, false>(sm + S_PSI, NJ, sm + S_PSI, NJ, 81, nullptr, 0, 0, 8 * s_, acc);
#pragma unroll
      for (int i = 0; i < 5; i++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const int m = 8 * i + lr;
          if (m < NJ) scatter_reg(nz, row[OFF_J + m], code[i][r], acc[i][0][r]);
        }
    };
    // j-phi / phi-j
    auto job_jf = [&]() {
      double acc[5][1][2];
      uint16_t cjf[5][2], cfj[5][2];
#pragma unroll
      for (int i = 0; i < 5; i++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const int m = 8 * i + lr, l = 2 * lk + r;
          const bool ok = m < NJ;
          cjf[i][r] = ok ? __ldg(cmap + SEC_JF + m * NF + l) : MAP_SKIP;
          cfj[i][r] = ok ? __ldg(cmap + SEC_FJ + l * NJ + m) : MAP_SKIP;
        }
      warp_mma_acc<5, 1, false>(sm + S_DIV, NJ, sm + S_CHI, 8, NQ, nullptr, 0, 0, 0, acc);
#pragma unroll
      for (int i = 0; i < 5; i++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const int m = 8 * i + lr, l = 2 * lk + r;
          if (m < NJ) {
            scatter_reg(nz, row[OFF_J + m], cjf[i][r], -sig_c * acc[i][0][r]);
            scatter_reg(nz, row[OFF_F + l], cfj[i][r], fj_sign * acc[i][0][r]);
          }
        }
    };
    // up / pu, component c
    auto job_d = [&](int c) {
      double acc[4][1][2];
      uint16_t cup[4][2], cpu[4][2];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const int a = 8 * i + lr, k = 2 * lk + r;
          const bool ok = a < 27 && k < NP;
          cup[i][r] = ok ? __ldg(cmap + SEC_UP + (c * 27 + a) * NP + k) : MAP_SKIP;
          cpu[i][r] = ok ? __ldg(cmap + SEC_PU + k * NU + c * 27 + a) : MAP_SKIP;
        }
      warp_mma_acc<4, 1, false>(sm + S_G + c * LDN, 3 * LDN, sm + S_PP, 4, NQ, nullptr, 0, 0, 0, acc);
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const int a = 8 * i + lr, k = 2 * lk + r;
          if (a < 27 && k < NP) {
            scatter_reg(nz, row[c * 27 + a], cup[i][r], -acc[i][0][r]);
            scatter_reg(nz, row[OFF_P + k], cpu[i][r], -acc[i][0][r]);
          }
        }
    };
    // uj / ju, columns n = 16u .. 16u+15 of the 108 (c,m) columns; B fragment = sqrt(w) (psi_m x B)_c formed on the fly
    auto job_uj = [&](int u) {
      double acc[4][2][2];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
      // this lane's B-fragment columns (one per column tile)
      int pb1[2], pb2[2];
      double w1[2], w2[2];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int nb = 16 * u + 8 * j + lr;
        const int cb = nb < 108 ? nb / NJ : 0, mb = nb < 108 ? nb - cb * NJ : 0;
        const int c1 = (cb + 1) % 3, c2 = (cb + 2) % 3;
        pb1[j] = c1 * NJ + mb;
        pb2[j] = c2 * NJ + mb;
        w1[j] = nb < 108 ? P.B[c2] : 0.0;   // (psi x B)_c = psi_{c+1} B_{c+2} - psi_{c+2} B_{c+1}
        w2[j] = nb < 108 ? -P.B[c1] : 0.0;
      }
#pragma unroll 2
      for (int k0 = 0; k0 < 28; k0 += 4) {
        const int kk = k0 + lk;
        const bool valid = kk < NQ;
        const int kc = valid ? kk : 0;
        double av[4], bv[2];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const double v = sm[S_N + kc * LDN + 8 * i + lr];
          av[i] = valid ? v : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const double v = sm[S_PSI + kc * 3 * NJ + pb1[j]] * w1[j] + sm[S_PSI + kc * 3 * NJ + pb2[j]] * w2[j];
          bv[j] = valid ? v : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
      }
      // scatter: K_uj[(c,a)][m] = -gamma R ; K_ju[m][(c,a)] = +sigma R
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const int n = 16 * u + 8 * j + 2 * lk + r;
          if (n >= 108) continue;
          const int c = n / NJ, m = n - c * NJ;
          uint16_t cuj[4], cju[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int a = 8 * i + lr;
            cuj[i] = a < 27 ? __ldg(cmap + SEC_UJ + (c * 27 + a) * NJ + m) : MAP_SKIP;
            cju[i] = a < 27 ? __ldg(cmap + SEC_JU + m * NU + c * 27 + a) : MAP_SKIP;
          }
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int a = 8 * i + lr;
            if (a < 27) {
              scatter_reg(nz, row[c * 27 + a], cuj[i], -P.gamma * acc[i][j][r]);
              scatter_reg(nz, row[OFF_J + m], cju[i], sig_c * acc[i][j][r]);
            }
          }
        }
    };
    // static deal (costs in MMAs: jj 105(140), uj 56, JF 35, D 28)
    if (warp < 5) job_jj(warp);
    if (solid) {
      if (warp == 5) job_jf();
    } else {
      if (warp < 3) job_d(warp);
      else if (warp == 3) job_jf();
      else if (warp == 4) job_uj(0);
      else {
        job_uj(2 * warp - 9);   // w5: 1,2  w6: 3,4  w7: 5,6
        job_uj(2 * warp - 8);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
static KParams make_kparams(const mhd_params_t& p) {
  KParams k;
  k.alpha = p.alpha; k.beta = p.beta; k.gamma = p.gamma; k.sigma = p.sigma; k.zeta_u = p.zeta_u; k.zeta_j = p.zeta_j;
  for (int i = 0; i < 3; i++) { k.B[i] = p.B[i]; k.f[i] = p.f[i]; k.g[i] = p.g[i]; }
  k.cell_solid = nullptr;
  k.cell_sigma = nullptr;
  return k;
}

static int g_sm_count = 0;
static int sm_count() {
  if (!g_sm_count) cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, g_device);
  return g_sm_count ? g_sm_count : 148;
}

template <class Kern>
static int set_smem(Kern k) {
  MHD_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return 0;
}

// d_r != nullptr: fused residual + Jacobian (residual_and_jacobian!)
int launch_jacobian(mhd_operator* op, const double* d_x, double* d_r) {
  MHD_CUDA(cudaMemsetAsync(op->d_nzval, 0, (size_t)op->nnz * sizeof(double), g_stream));
  if (d_r) MHD_CUDA(cudaMemsetAsync(d_r, 0, (size_t)op->nrows * sizeof(double), g_stream));
  KParams P = make_kparams(op->prm);
  P.cell_solid = op->d_cell_solid;
  P.cell_sigma = op->d_cell_sigma;
  const int conv = op->prm.convection;
  const bool zu = op->prm.zeta_u != 0.0;
  const int64_t grid64 = (int64_t)sm_count() * 2;
  const unsigned grid = (unsigned)(op->ncells < grid64 ? op->ncells : grid64);
#define JK(C, Z, R)                                                                                          \
  do {                                                                                                       \
    MHD_TRY(set_smem(jacobian_kernel<C, Z, R>));                                                             \
    jacobian_kernel<C, Z, R><<<grid, NT, SMEM_BYTES, g_stream>>>(op->ncells, op->nrows, op->d_tables, op->d_coords, \
        op->d_cell_nodes, op->d_gids, op->d_jsign, op->d_dir, d_x, op->d_rowptr, op->d_map, op->d_nzval, d_r, P); \
  } while (0)
#define JKR(C, Z) do { if (d_r) JK(C, Z, true); else JK(C, Z, false); } while (0)
  static int staged = -1;
  if (staged < 0) staged = getenv("MHD_JAC_STAGED") ? 1 : 0;
#define JK4(C, R)                                                                                            \
  do {                                                                                                       \
    MHD_TRY(set_smem(jacobian_kernel_v4<C, R>));                                                             \
    jacobian_kernel_v4<C, R><<<grid, NT, SMEM_BYTES, g_stream>>>(op->ncells, op->nrows, op->d_tables, op->d_coords, \
        op->d_cell_nodes, op->d_gids, op->d_jsign, op->d_dir, d_x, op->d_rowptr, op->d_map, op->d_nzval, d_r, P); \
  } while (0)
#define JK4R(C) do { if (d_r) JK4(C, true); else JK4(C, false); } while (0)
  prof_begin(PROF_JAC);
  if (!zu && !staged) {
    if (conv == 0) JK4R(0);
    else if (conv == 1) JK4R(1);
    else JK4R(2);
  } else if (conv == 0 && !zu) JKR(0, false);
  else if (conv == 0 && zu) JKR(0, true);
  else if (conv == 1 && !zu) JKR(1, false);
  else if (conv == 1 && zu) JKR(1, true);
  else if (conv == 2 && !zu) JKR(2, false);
  else JKR(2, true);
#undef JK4R
#undef JK4
#undef JKR
#undef JK
  prof_end(PROF_JAC);
  MHD_LAUNCH_CHECK();
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
    residual_kernel<C, Z><<<grid, NT, SMEM_BYTES, g_stream>>>(op->ncells, op->nrows, op->d_tables, op->d_coords, \
        op->d_cell_nodes, op->d_gids, op->d_jsign, op->d_dir, d_x, d_r, P);                                   \
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
