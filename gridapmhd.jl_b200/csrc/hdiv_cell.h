// Structure-exploiting Jacobian of the H1-HDiv formulation ("v6"): jac_fluid_h1_hdiv / jac_solid_h1_hdiv
// (src/weakforms.jl:283-312, :327-338) per cell, written like h1h1_cell.h as barrier-separated PHASES of a cooperative
// thread array so that tests/emul/emul_hdiv.cpp runs the same code on the CPU against the oracle.
//
// STATUS: verified on the CPU emulation only; selected with MHD_JAC_V6=1 (default: the tensor-core kernel of assembly.cu).
// It has not been timed on a B200 yet.
//
// What it exploits (232 k FMA per fluid cell instead of ~430 k in the panel-product kernel):
//   * uu by sum factorisation (sumfac_uu.h): Newton mass-type term, stiffness with the full J^-1 J^-T, Picard convection
//     as 21 (12 without Newton, 9 without convection) coefficient fields x three 1-D contractions;
//   * K_ju = -(sigma/gamma) K_uj^T for constant B, sigma, gamma: both come from V[a][m] = sum_q w N_a psi_m (3 components);
//   * jj over the symmetric half, j-phi / phi-j and up / pu from one integral each.
// Local numbering of a cell (UNPERMUTED, unlike assembly.cu): u (a + 27 c) | p (81 + k) | j (85 + m) | phi (121 + l).
#pragma once
#include <stdint.h>

#include "sumfac_uu.h"

#ifdef __CUDACC__
#define MHD_6HD __host__ __device__ __forceinline__
#define MHD_6UNROLL _Pragma("unroll")
#else
#define MHD_6HD inline
#define MHD_6UNROLL
#endif

namespace mhd {
namespace h6 {

constexpr int NQ = 27, NU = 81, NP = 4, NJ = 36, NF = 8;
constexpr int OFF_P = 81, OFF_J = 85, OFF_F = 121, NLOC = 129;

// ---- enumeration of the touched entries = order of this kernel's u16 scatter map
//   uu   : e = SEC_UU + (c*3+d)*729 + a*27 + b        row c*27+a,  col d*27+b
//   up   : e = SEC_UP + (c*27+a)*4 + k                row c*27+a,  col 81+k        pu: row 81+k, col c*27+a
//   uj   : e = SEC_UJ + c*972 + a*36 + m              row c*27+a,  col 85+m        ju: row 85+m, col c*27+a
//   jj   : e = SEC_JJ + m*36 + n                      row 85+m,    col 85+n
//   jphi : e = SEC_JF + m*8 + l                       row 85+m,    col 121+l       phi-j: row 121+l, col 85+m
constexpr int SEC_UU = 0;
constexpr int SEC_UP = SEC_UU + 9 * 729;
constexpr int SEC_PU = SEC_UP + NU * NP;
constexpr int SEC_UJ = SEC_PU + NU * NP;
constexpr int SEC_JU = SEC_UJ + 3 * 27 * NJ;
constexpr int SEC_JJ = SEC_JU + 3 * 27 * NJ;
constexpr int SEC_JF = SEC_JJ + NJ * NJ;
constexpr int SEC_FJ = SEC_JF + NJ * NF;
constexpr int NENT = SEC_FJ + NJ * NF;           // 14 913
constexpr int NENT_PAD = (NENT + 31) / 32 * 32;

inline void entry_rowcol(int e, int* li, int* lj) {
  if (e < SEC_UP) {
    const int cd = e / 729, ab = e % 729;
    *li = (cd / 3) * 27 + ab / 27;
    *lj = (cd % 3) * 27 + ab % 27;
  } else if (e < SEC_PU) {
    *li = (e - SEC_UP) / 4;
    *lj = OFF_P + (e - SEC_UP) % 4;
  } else if (e < SEC_UJ) {
    *li = OFF_P + (e - SEC_PU) % 4;
    *lj = (e - SEC_PU) / 4;
  } else if (e < SEC_JU) {
    const int r = e - SEC_UJ;
    *li = (r / 972) * 27 + (r % 972) / 36;
    *lj = OFF_J + r % 36;
  } else if (e < SEC_JJ) {
    const int r = e - SEC_JU;
    *li = OFF_J + r % 36;
    *lj = (r / 972) * 27 + (r % 972) / 36;
  } else if (e < SEC_JF) {
    *li = OFF_J + (e - SEC_JJ) / 36;
    *lj = OFF_J + (e - SEC_JJ) % 36;
  } else if (e < SEC_FJ) {
    *li = OFF_J + (e - SEC_JF) / 8;
    *lj = OFF_F + (e - SEC_JF) % 8;
  } else {
    *li = OFF_F + (e - SEC_FJ) % 8;
    *lj = OFF_J + (e - SEC_FJ) / 8;
  }
}

// packed reference tables = the T_* layout of common.h (op->d_tables)
constexpr int T_W = 0;
constexpr int T_GG = T_W + NQ;               // [27][8][3]
constexpr int T_NU = T_GG + NQ * 24;         // [27][27]
constexpr int T_DNU = T_NU + NQ * 27;        // [27][27][3]
constexpr int T_PP = T_DNU + NQ * 81;        // [27][4]
constexpr int T_PSI = T_PP + NQ * 4;         // [27][36][3]
constexpr int T_DPSI = T_PSI + NQ * 108;     // [27][36]
constexpr int T_CHI = T_DPSI + NQ * 36;      // [27][8]

struct Params {
  double alpha, beta, gamma, sigma, zeta_u, zeta_j;
  double B[3];
};

struct Shared {
  sf::Work W;            // coefficient fields + intermediates of the sum-factorised uu block
  double X[24];
  double stu[NU];        // local velocity values (Newton / Picard only)
  double sgn[NJ];        // RT sign flips
  double J[NQ][9];       // J[q][i*3+k] = d x_i / d xi_k
  double invJ[NQ][9];    // invJ[q][k*3+i] = d xi_k / d x_i
  double wdet[NQ];       // w |det J|
  double idet[NQ];       // 1 / det J
  double N[NQ][27];
  double Psi[NQ][3][NJ]; // Piola-mapped, signed RT basis: Psi[q][i][m]
  double Dv[NQ][NJ];     // its divergence
  double uq[NQ][3];
  double gur[NQ][9];     // reference gradient of u: gur[q][k*3+c] = sum_a d N_a / d xi_k u_c,a
  double D[4][NU];       // D[k][c*27+a] = int pi_k d_c N_a
  double E[4][NU];       // Mp^{-1} D
  double Minv[16];
  double sigma_cell;     // conductivity used in the j-phi block (cell_sigma on solid cells)
  double phi_sign;       // -1 on fluid cells (- div dj w), +1 on solid cells (+ w div dj)
  long long rowstart[NLOC];
  int32_t gid[NLOC];
};

// ------------------------------------------------------------------ phase 0: gather
MHD_6HD void phase_load(Shared& S, int tid, int nt, const double* coords, const int32_t* cell_nodes8, const int32_t* gids129,
                        const long long* rowstart129, const int8_t* jsign36, const double* dir, const double* x, const double* tab,
                        bool need_u, bool solid, double sigma_cell, double sigma_fluid) {
  for (int i = tid; i < 24; i += nt) S.X[i] = coords[(long long)cell_nodes8[i / 3] * 3 + i % 3];
  for (int i = tid; i < NLOC; i += nt) {
    const int32_t g = gids129[i];
    S.gid[i] = g;
    S.rowstart[i] = rowstart129 ? rowstart129[i] : -1;
    if (i < NU) S.stu[i] = need_u ? (g >= 0 ? x[g] : dir[-(long long)g - 1]) : 0.0;
  }
  for (int i = tid; i < NJ; i += nt) S.sgn[i] = (double)jsign36[i];
  for (int i = tid; i < NQ * 27; i += nt) S.N[i / 27][i % 27] = tab[T_NU + i];
  if (tid == 0) {
    S.sigma_cell = solid ? sigma_cell : sigma_fluid;
    S.phi_sign = solid ? 1.0 : -1.0;
  }
}

// ------------------------------------------------------------------ phase 1: geometry at the quadrature points
MHD_6HD void phase_geometry(Shared& S, int tid, int nt, const double* tab) {
  for (int q = tid; q < NQ; q += nt) {
    double* J = S.J[q];
    for (int i = 0; i < 9; i++) J[i] = 0.0;
    for (int v = 0; v < 8; v++)
      for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) J[i * 3 + k] += S.X[v * 3 + i] * tab[T_GG + (q * 8 + v) * 3 + k];
    const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
    const double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
    const double id = 1.0 / det;
    double* I = S.invJ[q];
    I[0] = c00 * id;
    I[1] = (J[2] * J[7] - J[1] * J[8]) * id;
    I[2] = (J[1] * J[5] - J[2] * J[4]) * id;
    I[3] = c01 * id;
    I[4] = (J[0] * J[8] - J[2] * J[6]) * id;
    I[5] = (J[2] * J[3] - J[0] * J[5]) * id;
    I[6] = c02 * id;
    I[7] = (J[1] * J[6] - J[0] * J[7]) * id;
    I[8] = (J[0] * J[4] - J[1] * J[3]) * id;
    S.idet[q] = id;
    S.wdet[q] = tab[T_W + q] * (det < 0.0 ? -det : det);
  }
}

// ------------------------------------------------------------------ phase 2: Piola map of the RT basis, u and its reference gradient
template <int CONV>
MHD_6HD void phase_mapped_bases(Shared& S, int tid, int nt, const double* tab) {
  for (int idx = tid; idx < NQ * NJ; idx += nt) {
    const int q = idx / NJ, m = idx % NJ;
    const double s = S.sgn[m] * S.idet[q];
    const double* J = S.J[q];
    const double r0 = tab[T_PSI + (q * NJ + m) * 3 + 0], r1 = tab[T_PSI + (q * NJ + m) * 3 + 1], r2 = tab[T_PSI + (q * NJ + m) * 3 + 2];
    for (int i = 0; i < 3; i++) S.Psi[q][i][m] = s * (J[i * 3 + 0] * r0 + J[i * 3 + 1] * r1 + J[i * 3 + 2] * r2);
    S.Dv[q][m] = s * tab[T_DPSI + q * NJ + m];
  }
  if (CONV != 0)
    for (int idx = tid; idx < NQ * 12; idx += nt) {
      const int q = idx / 12, r = idx % 12;
      double s = 0.0;
      if (r < 3) {
        for (int a = 0; a < 27; a++) s += S.N[q][a] * S.stu[r * 27 + a];
        S.uq[q][r] = s;
      } else {
        const int k = (r - 3) / 3, c = (r - 3) % 3;
        for (int a = 0; a < 27; a++) s += tab[T_DNU + (q * 27 + a) * 3 + k] * S.stu[c * 27 + a];
        S.gur[q][k * 3 + c] = s;
      }
    }
}

MHD_6HD void invert4(const double* M, double* out) {
  double a[4][8];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      a[i][j] = M[i * 4 + j];
      a[i][4 + j] = i == j ? 1.0 : 0.0;
    }
  for (int c = 0; c < 4; c++) {
    int p = c;
    double best = a[c][c] < 0 ? -a[c][c] : a[c][c];
    for (int r = c + 1; r < 4; r++) {
      const double v = a[r][c] < 0 ? -a[r][c] : a[r][c];
      if (v > best) { best = v; p = r; }
    }
    if (p != c)
      for (int j = 0; j < 8; j++) { const double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
    const double ip = 1.0 / a[c][c];
    for (int j = 0; j < 8; j++) a[c][j] *= ip;
    for (int r = 0; r < 4; r++)
      if (r != c) {
        const double fct = a[r][c];
        for (int j = 0; j < 8; j++) a[r][j] -= fct * a[c][j];
      }
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) out[i * 4 + j] = a[i][4 + j];
}

// ------------------------------------------------------------------ phase 3: coefficient fields of the uu block, D, Mp^{-1}
// fields (sumfac_uu.h): 0..8 mass type (Newton), 9..17 stiffness, 18..20 convection; unused fields are zeroed
template <int CONV, bool ZU>
MHD_6HD void phase_coefficients(Shared& S, int tid, int nt, const Params& P, const double* tab) {
  for (int idx = tid; idx < sf::NFIELD * NQ; idx += nt) {
    const int f = idx / NQ, q = idx % NQ;
    const double* I = S.invJ[q];
    double v = 0.0;
    if (f < 9) {
      if (CONV == 2) {
        // M_cd = alpha w d_d u_c, d_d u_c = sum_k invJ[k][d] gur[k][c]
        const int c = f / 3, d = f % 3;
        v = P.alpha * S.wdet[q] * (I[0 * 3 + d] * S.gur[q][0 * 3 + c] + I[1 * 3 + d] * S.gur[q][1 * 3 + c] + I[2 * 3 + d] * S.gur[q][2 * 3 + c]);
      }
    } else if (f < 18) {
      const int m = (f - 9) / 3, n = (f - 9) % 3;
      v = P.beta * S.wdet[q] * (I[m * 3 + 0] * I[n * 3 + 0] + I[m * 3 + 1] * I[n * 3 + 1] + I[m * 3 + 2] * I[n * 3 + 2]);
    } else if (CONV != 0) {
      const int n = f - 18;
      v = P.alpha * S.wdet[q] * (I[n * 3 + 0] * S.uq[q][0] + I[n * 3 + 1] * S.uq[q][1] + I[n * 3 + 2] * S.uq[q][2]);
    }
    S.W.C[f][q] = v;
  }
  for (int idx = tid; idx < NP * NU; idx += nt) {
    const int k = idx / NU, ca = idx % NU, c = ca / 27, a = ca % 27;
    double s = 0.0;
    for (int q = 0; q < NQ; q++) {
      const double* I = S.invJ[q];
      const double* g = tab + T_DNU + (q * 27 + a) * 3;
      s += S.wdet[q] * tab[T_PP + q * 4 + k] * (g[0] * I[0 * 3 + c] + g[1] * I[1 * 3 + c] + g[2] * I[2 * 3 + c]);
    }
    S.D[k][ca] = s;
  }
  if (ZU && tid == nt - 1) {
    double Mp[16];
    for (int i = 0; i < 16; i++) {
      double s = 0.0;
      for (int q = 0; q < NQ; q++) s += S.wdet[q] * tab[T_PP + q * 4 + i / 4] * tab[T_PP + q * 4 + i % 4];
      Mp[i] = s;
    }
    invert4(Mp, S.Minv);
  }
}

MHD_6HD void phase_projection(Shared& S, int tid, int nt) {
  for (int idx = tid; idx < NP * NU; idx += nt) {
    const int k = idx / NU, ca = idx % NU;
    S.E[k][ca] = S.Minv[k * 4 + 0] * S.D[0][ca] + S.Minv[k * 4 + 1] * S.D[1][ca] + S.Minv[k * 4 + 2] * S.D[2][ca] +
                 S.Minv[k * 4 + 3] * S.D[3][ca];
  }
}

// ------------------------------------------------------------------ entries
// job list (slots padded to whole warps):
//   [0,736)     uu   729 pairs (a,b): third contraction of the sum factorisation, 9 entries each (+ zeta_u D' E)
//   [736,832)   uj   81 jobs: 3 x 4 (a,m) tiles x 3 components -> uj and ju      (a in {3ta..3ta+2}, m in {tm + 9i})
//   [832,928)   jj   78 jobs: 3 x 3 tiles of the upper triangle, mirrored            (m in {tm+12i}, n in {tn+12j}, tm <= tn)
//   [928,1216)  jphi 288 (m,l) pairs -> j-phi and phi-j
//   [1216,1540) up   324 (c,a,k) -> up and pu
constexpr int JOB_UU = 0, JOB_UJ = 736, JOB_JJ = 832, JOB_JF = 928, JOB_UP = 1216, JOB_END = 1540;

template <bool ZU, class Store>
MHD_6HD void job_uu(const Shared& S, const sf::Tables& T, int ab, const Params& P, Store& store) {
  const int a = ab / 27, b = ab % 27;
  const int ii = 3 * T.ijk[a][0] + T.ijk[b][0], jj = 3 * T.ijk[a][1] + T.ijk[b][1], kk = 3 * T.ijk[a][2] + T.ijk[b][2];
  const int t2 = (jj * 9 + kk) * 3;
  double val[9], s = 0.0;
  MHD_6UNROLL
  for (int f = 0; f < sf::NFIELD; f++) {
    const double* p = T.P[sf::pair_index(f, 0)][ii];
    const double* t = S.W.T2[f] + t2;
    const double v = p[0] * t[0] + p[1] * t[1] + p[2] * t[2];
    if (f < 9) val[f] = v;
    else s += v;
  }
  MHD_6UNROLL
  for (int cd = 0; cd < 9; cd++) {
    const int c = cd / 3, d = cd % 3;
    double v = val[cd] + (c == d ? s : 0.0);
    const int ra = c * 27 + a, cb = d * 27 + b;
    if (ZU) v += P.zeta_u * (S.D[0][ra] * S.E[0][cb] + S.D[1][ra] * S.E[1][cb] + S.D[2][ra] * S.E[2][cb] + S.D[3][ra] * S.E[3][cb]);
    store(SEC_UU + cd * 729 + ab, ra, cb, v);
  }
}

template <class Store>
MHD_6HD void job_uj(const Shared& S, int job, const Params& P, Store& store) {
  const int ta = job / 9, tm = job % 9;
  double acc[3][4][3];
  MHD_6UNROLL
  for (int i = 0; i < 3; i++)
    MHD_6UNROLL
    for (int j = 0; j < 4; j++)
      MHD_6UNROLL
      for (int k = 0; k < 3; k++) acc[i][j][k] = 0.0;
  for (int q = 0; q < NQ; q++) {
    const double w = S.wdet[q];
    double wn[3];
    MHD_6UNROLL
    for (int i = 0; i < 3; i++) wn[i] = w * S.N[q][3 * ta + i];
    MHD_6UNROLL
    for (int k = 0; k < 3; k++) {
      double g[4];
      MHD_6UNROLL
      for (int j = 0; j < 4; j++) g[j] = S.Psi[q][k][tm + 9 * j];
      MHD_6UNROLL
      for (int i = 0; i < 3; i++)
        MHD_6UNROLL
        for (int j = 0; j < 4; j++) acc[i][j][k] += wn[i] * g[j];
    }
  }
  // uj: -gamma (dj x B).v = -gamma N_a (psi_m x B)_c ; ju: -sigma (du x B).s = +sigma N_b (psi_m x B)_d
  MHD_6UNROLL
  for (int i = 0; i < 3; i++)
    MHD_6UNROLL
    for (int j = 0; j < 4; j++) {
      const int a = 3 * ta + i, m = tm + 9 * j;
      const double* v = acc[i][j];
      const double x[3] = {v[1] * P.B[2] - v[2] * P.B[1], v[2] * P.B[0] - v[0] * P.B[2], v[0] * P.B[1] - v[1] * P.B[0]};
      MHD_6UNROLL
      for (int c = 0; c < 3; c++) {
        store(SEC_UJ + c * 972 + a * NJ + m, c * 27 + a, OFF_J + m, -P.gamma * x[c]);
        store(SEC_JU + c * 972 + a * NJ + m, OFF_J + m, c * 27 + a, P.sigma * x[c]);
      }
    }
}

// jj: 3 x 3 tiles of the upper triangle (m in {tm, tm+12, tm+24}, n in {tn, tn+12, tn+24}, tm <= tn < 12): 78 jobs.
// (4 x 4 tiles would need fewer operand loads but leave only 45 threads on the longest jobs of the cell.)
template <bool ZJ, class Store>
MHD_6HD void job_jj(const Shared& S, int job, const Params& P, Store& store) {
  int tm = 0, r = job;
  while (r >= 12 - tm) {
    r -= 12 - tm;
    tm++;
  }
  const int tn = tm + r;
  double acc[3][3];
  MHD_6UNROLL
  for (int i = 0; i < 3; i++)
    MHD_6UNROLL
    for (int j = 0; j < 3; j++) acc[i][j] = 0.0;
  for (int q = 0; q < NQ; q++) {
    const double w = S.wdet[q];
    MHD_6UNROLL
    for (int k = 0; k < 3; k++) {
      double gm[3], gn[3];
      MHD_6UNROLL
      for (int i = 0; i < 3; i++) {
        gm[i] = w * S.Psi[q][k][tm + 12 * i];
        gn[i] = S.Psi[q][k][tn + 12 * i];
      }
      MHD_6UNROLL
      for (int i = 0; i < 3; i++)
        MHD_6UNROLL
        for (int j = 0; j < 3; j++) acc[i][j] += gm[i] * gn[j];
    }
    if (ZJ) {  // zeta_j div dj div s
      double gm[3], gn[3];
      MHD_6UNROLL
      for (int i = 0; i < 3; i++) {
        gm[i] = w * P.zeta_j * S.Dv[q][tm + 12 * i];
        gn[i] = S.Dv[q][tn + 12 * i];
      }
      MHD_6UNROLL
      for (int i = 0; i < 3; i++)
        MHD_6UNROLL
        for (int j = 0; j < 3; j++) acc[i][j] += gm[i] * gn[j];
    }
  }
  MHD_6UNROLL
  for (int i = 0; i < 3; i++)
    MHD_6UNROLL
    for (int j = 0; j < 3; j++) {
      const int m = tm + 12 * i, n = tn + 12 * j;
      store(SEC_JJ + m * NJ + n, OFF_J + m, OFF_J + n, acc[i][j]);
      if (tm != tn) store(SEC_JJ + n * NJ + m, OFF_J + n, OFF_J + m, acc[i][j]);
    }
}

template <int CONV, bool ZU, bool ZJ, class Store>
MHD_6HD void phase_entries(const Shared& S, const sf::Tables& T, int tid, int nt, const Params& P, const double* tab, Store& store) {
  for (int slot = tid; slot < JOB_END; slot += nt) {
    if (slot < JOB_UJ) {
      if (slot < 729) job_uu<ZU>(S, T, slot, P, store);
    } else if (slot < JOB_JJ) {
      if (slot - JOB_UJ < 81) job_uj(S, slot - JOB_UJ, P, store);
    } else if (slot < JOB_JF) {
      if (slot - JOB_JJ < 78) job_jj<ZJ>(S, slot - JOB_JJ, P, store);
    } else if (slot < JOB_UP) {
      // j-phi: -sigma dphi div s ; phi-j: -(+ on solid cells) div dj w
      const int idx = slot - JOB_JF, m = idx / NF, l = idx % NF;
      double s = 0.0;
      for (int q = 0; q < NQ; q++) s += S.wdet[q] * tab[T_CHI + q * NF + l] * S.Dv[q][m];
      store(SEC_JF + idx, OFF_J + m, OFF_F + l, -S.sigma_cell * s);
      store(SEC_FJ + idx, OFF_F + l, OFF_J + m, S.phi_sign * s);
    } else {
      const int idx = slot - JOB_UP, ca = idx / 4, k = idx % 4;
      const double v = -S.D[k][ca];
      store(SEC_UP + idx, ca, OFF_P + k, v);
      store(SEC_PU + idx, OFF_P + k, ca, v);
    }
  }
}

// host: the 1-D factors of the tensor-product Q2 tables.  ijk[a] = (i,j,k) of node a.  By the partition of unity of the
// 1-D basis, l_i(x_q1) = sum over the nodes with first index i of N_a at the point (q1, 0, 0), likewise for the derivative.
// Returns the largest deviation of the 3-D tables from the products of the factors (tensor structure check).
inline double derive_tensor_tables(const double* nu /*[27][27]*/, const double* dnu /*[27][27][3]*/, const int8_t* ijk /*[27][3]*/,
                                   sf::Tables* T) {
  double l[2][3][3];  // l[m][i][q]
  for (int i = 0; i < 3; i++)
    for (int q = 0; q < 3; q++) {
      double v = 0.0, d = 0.0;
      for (int a = 0; a < 27; a++)
        if (ijk[a * 3 + 0] == i) {
          v += nu[q * 27 + a];
          d += dnu[(q * 27 + a) * 3 + 0];
        }
      l[0][i][q] = v;
      l[1][i][q] = d;
    }
  for (int m = 0; m < 2; m++)
    for (int n = 0; n < 2; n++)
      for (int i = 0; i < 3; i++)
        for (int i2 = 0; i2 < 3; i2++)
          for (int q = 0; q < 3; q++) T->P[2 * m + n][3 * i + i2][q] = l[m][i][q] * l[n][i2][q];
  for (int a = 0; a < 27; a++)
    for (int d = 0; d < 3; d++) T->ijk[a][d] = ijk[a * 3 + d];
  double dev = 0.0;
  for (int q = 0; q < 27; q++) {
    const int q1 = q % 3, q2 = (q / 3) % 3, q3 = q / 9;
    for (int a = 0; a < 27; a++) {
      const int i = ijk[a * 3], j = ijk[a * 3 + 1], k = ijk[a * 3 + 2];
      const double ref[4] = {l[0][i][q1] * l[0][j][q2] * l[0][k][q3], l[1][i][q1] * l[0][j][q2] * l[0][k][q3],
                             l[0][i][q1] * l[1][j][q2] * l[0][k][q3], l[0][i][q1] * l[0][j][q2] * l[1][k][q3]};
      const double got[4] = {nu[q * 27 + a], dnu[(q * 27 + a) * 3], dnu[(q * 27 + a) * 3 + 1], dnu[(q * 27 + a) * 3 + 2]};
      for (int t = 0; t < 4; t++) {
        const double e = ref[t] - got[t];
        if ((e < 0 ? -e : e) > dev) dev = e < 0 ? -e : e;
      }
    }
  }
  return dev;
}

}  // namespace h6
}  // namespace mhd
