// Per-cell arithmetic of the H1-H1 formulation (u Q2, p P1disc, phi Q3 continuous; the current is eliminated):
// jac_fluid_h1_h1 / res_fluid_h1_h1 (src/weakforms.jl:440-466, :415-438) and, on solid cells, jac/res_solid_h1_h1
// (:468-478; only the phi-phi Laplacian survives there because the u and p dofs of a solid cell are absent).
//
// The code is organised as PHASES of a cooperative thread array working on one cell: every phase is a function of
// (shared cell data, thread id, thread count) and phases are separated by a barrier.  The CUDA kernels in h1h1.cu run
// the phases with a CTA per cell; tests/emul_h1h1.cpp compiles this very header with g++ and runs the same phases with
// a loop over the thread ids, so the arithmetic and the entry enumeration of the scatter map are checked against the
// oracle without a GPU.
//
// Local numbering of a cell: u (a + 27 c, a = Q2 node, c = component) | p (81 + k) | phi (85 + l), 149 dofs.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MHD_HD __host__ __device__ __forceinline__
#define MHD_UNROLL _Pragma("unroll")
#else
#define MHD_HD inline
#define MHD_UNROLL
#endif

// product and sum rounded separately, like the reference (see hdiv7_cell.h): the Jacobian-matrix sum cancels |X| / h digits
#ifndef MHD_MULADD_REF
#ifdef __CUDA_ARCH__
#define MHD_MULADD_REF(a, b, c) __dadd_rn(__dmul_rn((a), (b)), (c))
#else
#define MHD_MULADD_REF(a, b, c) ((a) * (b) + (c))
#endif
#endif
namespace mhd {
namespace h1 {

constexpr int NQ = 27;
constexpr int NU = 81, NP = 4, NF = 64;
constexpr int OFF_U = 0, OFF_P = 81, OFF_F = 85;
constexpr int NLOC = 149;

// ---- enumeration of the touched entries of a cell = order of the u16 scatter map (6 blocks; pp, p-phi, phi-p are
// never inserted).  Chosen so that consecutive threads of a phase read consecutive map codes.
//   uu    : [(c,d)][a][b]          e = SEC_UU + (c*3+d)*729 + a*27 + b      row c*27+a, col d*27+b
//   up    : [(c,a)][k]             e = SEC_UP + (c*27+a)*4 + k              row c*27+a, col 81+k
//   pu    : [(c,a)][k]             e = SEC_PU + (c*27+a)*4 + k              row 81+k,   col c*27+a
//   u-phi : [c][a][l]              e = SEC_UF + c*1728 + a*64 + l           row c*27+a, col 85+l
//   phi-u : [c][a][l]              e = SEC_FU + c*1728 + a*64 + l           row 85+l,   col c*27+a
//   phi-phi: [l][m]                e = SEC_FF + l*64 + m                    row 85+l,   col 85+m
constexpr int SEC_UU = 0;
constexpr int SEC_UP = SEC_UU + 9 * 729;
constexpr int SEC_PU = SEC_UP + NU * NP;
constexpr int SEC_UF = SEC_PU + NU * NP;
constexpr int SEC_FU = SEC_UF + 3 * 27 * NF;
constexpr int SEC_FF = SEC_FU + 3 * 27 * NF;
constexpr int NENT = SEC_FF + NF * NF;            // 21673 entries per cell
constexpr int NENT_PAD = (NENT + 31) / 32 * 32;   // per-cell stride of the u16 map

// ---- packed reference tables (doubles), gradients stored direction-major so that the basis index is contiguous
constexpr int T_W = 0;                      // [27]
constexpr int T_GG = T_W + NQ;              // [27][8][3]   d(vertex function v)/d xi_k
constexpr int T_N = T_GG + NQ * 24;         // [27][27]     N_a
constexpr int T_DN = T_N + NQ * 27;         // [27][3][27]  d N_a / d xi_k
constexpr int T_PP = T_DN + NQ * 81;        // [27][4]
constexpr int T_DF = T_PP + NQ * 4;         // [27][3][64]  d phi_l / d xi_k
constexpr int T_TOTAL = T_DF + NQ * 192;

// host: pack the caller's tables (mhd_tables_h1h1_t layouts) into T[T_TOTAL]
inline void pack_tables(const double* w, const double* geo_grad, const double* u_val, const double* u_grad, const double* p_val,
                        const double* phi_grad, double* T) {
  for (int q = 0; q < NQ; q++) {
    T[T_W + q] = w[q];
    for (int i = 0; i < 24; i++) T[T_GG + q * 24 + i] = geo_grad[q * 24 + i];
    for (int a = 0; a < 27; a++) {
      T[T_N + q * 27 + a] = u_val[q * 27 + a];
      for (int k = 0; k < 3; k++) T[T_DN + (q * 3 + k) * 27 + a] = u_grad[(q * 27 + a) * 3 + k];
    }
    for (int k = 0; k < 4; k++) T[T_PP + q * 4 + k] = p_val[q * 4 + k];
    for (int l = 0; l < 64; l++)
      for (int k = 0; k < 3; k++) T[T_DF + (q * 3 + k) * 64 + l] = phi_grad[(q * 64 + l) * 3 + k];
  }
}

struct Params {
  double alpha, beta, gamma, zeta_u;
  double B[3], f[3];
};

// Cell data shared by the threads of a cell (shared memory on the device).
struct Shared {
  double X[24];
  double st[NLOC];
  double invJ[NQ][9];   // invJ[q][k*3+i] = d xi_k / d x_i
  double wdet[NQ];
  double N[NQ][27];
  double gU[NQ][3][27];  // physical d_i N_a
  double gF[NQ][3][64];  // physical d_i phi_l
  double pp[NQ][4];
  double uq[NQ][3];
  double gu[NQ][9];      // gu[q][d*3+c] = d_d u_c
  double adv[NQ][27];    // u . grad N_b
  double Mw[NQ][9];      // wdet * (gamma (|B|^2 delta_cd - B_c B_d) + alpha d_d u_c), index c*3+d
  double D[4][NU];       // D[k][c*27+a] = int pi_k d_c N_a
  double E[4][NU];       // Mp^{-1} D
  double Minv[16];
  // residual only
  double pq[NQ], proj[NQ], Fp[NQ];
  double gphi[NQ][3], Fu[NQ][3], Gf[NQ][3];
  double Gu[NQ][9];      // coefficient of d_d N_a in row (a,c): index c*3+d
  double rhs[4], coef[4];
  long long rowstart[NLOC];
  int32_t gid[NLOC];
};

// ------------------------------------------------------------------ phase 0: gather
// gids: >= 0 local free id, < 0: -(index into dir)-1.  rowstart may be null (residual).
MHD_HD void phase_load(Shared& S, int tid, int nt, const double* coords, const int32_t* cell_nodes8, const int32_t* gids149,
                       const long long* rowstart149, const double* dir, const double* x, const double* tab) {
  for (int i = tid; i < 24; i += nt) S.X[i] = coords[(long long)cell_nodes8[i / 3] * 3 + i % 3];
  for (int i = tid; i < NLOC; i += nt) {
    const int32_t g = gids149[i];
    S.gid[i] = g;
    S.st[i] = g >= 0 ? x[g] : dir[-(long long)g - 1];
    S.rowstart[i] = rowstart149 ? rowstart149[i] : -1;
  }
  for (int i = tid; i < NQ * 27; i += nt) S.N[i / 27][i % 27] = tab[T_N + i];
  for (int i = tid; i < NQ * 4; i += nt) S.pp[i / 4][i % 4] = tab[T_PP + i];
}

// ------------------------------------------------------------------ phase 1: geometry at the quadrature points
MHD_HD void phase_geometry(Shared& S, int tid, int nt, const double* tab) {
  for (int q = tid; q < NQ; q += nt) {
    double J[9];  // J[i*3+k] = d x_i / d xi_k
    for (int i = 0; i < 9; i++) J[i] = 0.0;
    for (int v = 0; v < 8; v++)
      for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) J[i * 3 + k] = MHD_MULADD_REF(S.X[v * 3 + i], tab[T_GG + (q * 8 + v) * 3 + k], J[i * 3 + k]);
    const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
    const double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
    const double id = 1.0 / det;
    // inverse: invJ[k][i] = cofactor(J)[i][k] / det
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
    S.wdet[q] = tab[T_W + q] * (det < 0.0 ? -det : det);
  }
}

// ------------------------------------------------------------------ phase 2: physical gradients
MHD_HD void phase_gradients(Shared& S, int tid, int nt, const double* tab) {
  for (int idx = tid; idx < NQ * 27; idx += nt) {
    const int q = idx / 27, a = idx % 27;
    const double* I = S.invJ[q];
    const double r0 = tab[T_DN + (q * 3 + 0) * 27 + a], r1 = tab[T_DN + (q * 3 + 1) * 27 + a], r2 = tab[T_DN + (q * 3 + 2) * 27 + a];
    for (int i = 0; i < 3; i++) S.gU[q][i][a] = r0 * I[i] + r1 * I[3 + i] + r2 * I[6 + i];
  }
  for (int idx = tid; idx < NQ * 64; idx += nt) {
    const int q = idx / 64, l = idx % 64;
    const double* I = S.invJ[q];
    const double r0 = tab[T_DF + (q * 3 + 0) * 64 + l], r1 = tab[T_DF + (q * 3 + 1) * 64 + l], r2 = tab[T_DF + (q * 3 + 2) * 64 + l];
    for (int i = 0; i < 3; i++) S.gF[q][i][l] = r0 * I[i] + r1 * I[3 + i] + r2 * I[6 + i];
  }
}

// ------------------------------------------------------------------ phase 3: u and grad u at the points
MHD_HD void phase_point_values(Shared& S, int tid, int nt) {
  for (int idx = tid; idx < NQ * 12; idx += nt) {
    const int q = idx / 12, r = idx % 12;
    double s = 0.0;
    if (r < 3) {
      for (int a = 0; a < 27; a++) s += S.N[q][a] * S.st[r * 27 + a];
      S.uq[q][r] = s;
    } else {
      const int d = (r - 3) / 3, c = (r - 3) % 3;
      for (int a = 0; a < 27; a++) s += S.gU[q][d][a] * S.st[c * 27 + a];
      S.gu[q][d * 3 + c] = s;
    }
  }
}

// 4x4 inverse by Gauss-Jordan with partial pivoting (Mp is SPD)
MHD_HD void invert4(const double* M, double* out) {
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

// ------------------------------------------------------------------ phase 4 (Jacobian): point coefficients, D, Mp^{-1}
template <int CONV, bool ZU>
MHD_HD void phase_jac_coefficients(Shared& S, int tid, int nt, const Params& P) {
  if (CONV != 0)
    for (int idx = tid; idx < NQ * 27; idx += nt) {
      const int q = idx / 27, b = idx % 27;
      S.adv[q][b] = S.uq[q][0] * S.gU[q][0][b] + S.uq[q][1] * S.gU[q][1][b] + S.uq[q][2] * S.gU[q][2][b];
    }
  const double B2 = P.B[0] * P.B[0] + P.B[1] * P.B[1] + P.B[2] * P.B[2];
  for (int idx = tid; idx < NQ * 9; idx += nt) {
    const int q = idx / 9, c = (idx % 9) / 3, d = idx % 3;
    double m = P.gamma * ((c == d ? B2 : 0.0) - P.B[c] * P.B[d]);
    if (CONV == 2) m += P.alpha * S.gu[q][d * 3 + c];
    S.Mw[q][c * 3 + d] = S.wdet[q] * m;
  }
  for (int idx = tid; idx < NP * NU; idx += nt) {
    const int k = idx / NU, ca = idx % NU, c = ca / 27, a = ca % 27;
    double s = 0.0;
    for (int q = 0; q < NQ; q++) s += S.wdet[q] * S.pp[q][k] * S.gU[q][c][a];
    S.D[k][ca] = s;
  }
  if (ZU && tid == nt - 1) {
    double Mp[16];
    for (int i = 0; i < 16; i++) {
      double s = 0.0;
      for (int q = 0; q < NQ; q++) s += S.wdet[q] * S.pp[q][i / 4] * S.pp[q][i % 4];
      Mp[i] = s;
    }
    invert4(Mp, S.Minv);
  }
}

// ------------------------------------------------------------------ phase 5 (Jacobian, zeta_u != 0): E = Mp^{-1} D
MHD_HD void phase_jac_projection(Shared& S, int tid, int nt) {
  for (int idx = tid; idx < NP * NU; idx += nt) {
    const int k = idx / NU, ca = idx % NU;
    S.E[k][ca] = S.Minv[k * 4 + 0] * S.D[0][ca] + S.Minv[k * 4 + 1] * S.D[1][ca] + S.Minv[k * 4 + 2] * S.D[2][ca] +
                 S.Minv[k * 4 + 3] * S.D[3][ca];
  }
}

// ------------------------------------------------------------------ phase 6 (Jacobian): the entries
// store(e, li, lj, v): entry e of the enumeration = local (row li, column lj), value v.
//
// The phase is bound by shared-memory operand reads (ncu: L1TEX 87 % with one entry per thread), so every thread owns a
// REGISTER TILE of entries and each operand it loads feeds several FMAs.  Tiles are strided (a, a+14 | l, l+16, ...) so
// that the lanes of a warp read consecutive shared-memory words and consecutive map codes.  One job list per cell,
// job kinds padded to whole warps:
//   slots [0,224)   uu      196 jobs: 2 x 2 (a,b) node pairs x 9 components      (a in {ta, ta+14}, b in {tb, tb+14})
//   slots [224,384) phi-phi 136 jobs: 4 x 4 tiles of the upper triangle, mirrored (l in {tl+16i}, m in {tm+16j}, tl <= tm)
//   slots [384,544) u-phi   144 jobs: 3 x 4 (a,l) pairs x 3 directions -> u-phi and phi-u (a in {3ta..3ta+2}, l in {tl+16i})
//   slots [544,868) up, pu  324 jobs: one (c,a,k) each (the integrals are the D of the coefficient phase)
constexpr int JOB_UU = 0, JOB_FF = 224, JOB_UF = 384, JOB_UP = 544, JOB_END = 868;
constexpr int NJOB_UU = 196, NJOB_FF = 136, NJOB_UF = 144;

template <int CONV, bool ZU, class Store>
MHD_HD void job_uu(const Shared& S, int job, const Params& P, Store& store) {
  const int ta = job / 14, tb = job % 14;
  const int a[2] = {ta, ta + 14 < 27 ? ta + 14 : 26}, b[2] = {tb, tb + 14 < 27 ? tb + 14 : 26};
  const int na_ok = ta + 14 < 27 ? 2 : 1, nb_ok = tb + 14 < 27 ? 2 : 1;
  double s[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  double mass[2][2] = {{0.0, 0.0}, {0.0, 0.0}};  // CONV != 2: sum_q wdet N_a N_b
  double t[2][2][9];                              // CONV == 2: sum_q N_a N_b Mw_cd
  if (CONV == 2)
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++)
        for (int k = 0; k < 9; k++) t[i][j][k] = 0.0;
  for (int q = 0; q < NQ; q++) {
    const double w = S.wdet[q];
    double g[2][2], nn[2][2];
    {
      double na[2], ga[2][3], nb[2], gb[2][3], ab[2];
      MHD_UNROLL
      for (int i = 0; i < 2; i++) {
        na[i] = S.N[q][a[i]];
        nb[i] = S.N[q][b[i]];
        for (int k = 0; k < 3; k++) {
          ga[i][k] = S.gU[q][k][a[i]];
          gb[i][k] = S.gU[q][k][b[i]];
        }
        ab[i] = CONV != 0 ? P.alpha * S.adv[q][b[i]] : 0.0;
      }
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
          double v = P.beta * (ga[i][0] * gb[j][0] + ga[i][1] * gb[j][1] + ga[i][2] * gb[j][2]);
          if (CONV != 0) v += na[i] * ab[j];
          g[i][j] = v;
          nn[i][j] = na[i] * nb[j];
        }
    }
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) {
        s[i][j] += w * g[i][j];
        if (CONV != 2) mass[i][j] += w * nn[i][j];
      }
    if (CONV == 2)
      MHD_UNROLL
      for (int k = 0; k < 9; k++) {
        const double m = S.Mw[q][k];
        for (int i = 0; i < 2; i++)
          for (int j = 0; j < 2; j++) t[i][j][k] += nn[i][j] * m;
      }
  }
  const double B2 = P.B[0] * P.B[0] + P.B[1] * P.B[1] + P.B[2] * P.B[2];
  MHD_UNROLL
  for (int i = 0; i < 2; i++)
    MHD_UNROLL
    for (int j = 0; j < 2; j++)
      MHD_UNROLL
      for (int cd = 0; cd < 9; cd++) {
        if (i < na_ok && j < nb_ok) {  // static loop bounds: the accumulators stay in registers
          const int c = cd / 3, d = cd % 3;
          double v = CONV == 2 ? t[i][j][c * 3 + d] : P.gamma * ((c == d ? B2 : 0.0) - P.B[c] * P.B[d]) * mass[i][j];
          if (c == d) v += s[i][j];
          const int ra = c * 27 + a[i], cb = d * 27 + b[j];
          if (ZU)
            v += P.zeta_u * (S.D[0][ra] * S.E[0][cb] + S.D[1][ra] * S.E[1][cb] + S.D[2][ra] * S.E[2][cb] + S.D[3][ra] * S.E[3][cb]);
          store(SEC_UU + (c * 3 + d) * 729 + a[i] * 27 + b[j], ra, cb, v);
        }
      }
}

template <class Store>
MHD_HD void job_ff(const Shared& S, int job, Store& store) {
  int tl = 0, r = job;
  while (r >= 16 - tl) {
    r -= 16 - tl;
    tl++;
  }
  const int tm = tl + r;
  double acc[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
  for (int q = 0; q < NQ; q++) {
    const double w = S.wdet[q];
    MHD_UNROLL
    for (int k = 0; k < 3; k++) {
      double gl[4], gm[4];
      for (int i = 0; i < 4; i++) {
        gl[i] = w * S.gF[q][k][tl + 16 * i];
        gm[i] = S.gF[q][k][tm + 16 * i];
      }
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) acc[i][j] += gl[i] * gm[j];
    }
  }
  MHD_UNROLL
  for (int i = 0; i < 4; i++)
    MHD_UNROLL
    for (int j = 0; j < 4; j++) {
      const int l = tl + 16 * i, m = tm + 16 * j;
      store(SEC_FF + l * NF + m, OFF_F + l, OFF_F + m, acc[i][j]);
      if (tl != tm) store(SEC_FF + m * NF + l, OFF_F + m, OFF_F + l, acc[i][j]);
    }
}

template <class Store>
MHD_HD void job_uf(const Shared& S, int job, const Params& P, Store& store) {
  const int ta = job / 16, tl = job % 16;
  double acc[3][4][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 3; k++) acc[i][j][k] = 0.0;
  for (int q = 0; q < NQ; q++) {
    const double w = S.wdet[q];
    double wn[3];
    for (int i = 0; i < 3; i++) wn[i] = w * S.N[q][3 * ta + i];
    MHD_UNROLL
    for (int k = 0; k < 3; k++) {
      double g[4];
      for (int j = 0; j < 4; j++) g[j] = S.gF[q][k][tl + 16 * j];
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) acc[i][j][k] += wn[i] * g[j];
    }
  }
  // u-phi: -gamma grad dphi.(v x B) = -gamma N_a (B x grad phi_l)_c ; phi-u: -(du x B).grad w = -N_a (B x grad phi_l)_c
  MHD_UNROLL
  for (int i = 0; i < 3; i++)
    MHD_UNROLL
    for (int j = 0; j < 4; j++) {
      const int a = 3 * ta + i, l = tl + 16 * j;
      const double* v = acc[i][j];
      const double x[3] = {P.B[1] * v[2] - P.B[2] * v[1], P.B[2] * v[0] - P.B[0] * v[2], P.B[0] * v[1] - P.B[1] * v[0]};
      for (int c = 0; c < 3; c++) {
        store(SEC_UF + c * 27 * NF + a * NF + l, c * 27 + a, OFF_F + l, -P.gamma * x[c]);
        store(SEC_FU + c * 27 * NF + a * NF + l, OFF_F + l, c * 27 + a, -x[c]);
      }
    }
}

template <int CONV, bool ZU, class Store>
MHD_HD void phase_jac_entries(const Shared& S, int tid, int nt, const Params& P, Store& store) {
  for (int slot = tid; slot < JOB_END; slot += nt) {
    if (slot < JOB_FF) {
      // uu: beta grad du : grad v + gamma (du x B).(v x B) + zeta_u Pi_p(du) div v + alpha v.(conv(u,grad du) + conv(du,grad u))
      if (slot - JOB_UU < NJOB_UU) job_uu<CONV, ZU>(S, slot - JOB_UU, P, store);
    } else if (slot < JOB_UF) {
      // phi-phi: grad dphi . grad w (symmetric)
      if (slot - JOB_FF < NJOB_FF) job_ff(S, slot - JOB_FF, store);
    } else if (slot < JOB_UP) {
      if (slot - JOB_UF < NJOB_UF) job_uf(S, slot - JOB_UF, P, store);
    } else {
      // up: -dp div v ; pu: -div du q
      const int idx = slot - JOB_UP, ca = idx / 4, k = idx % 4;
      const double v = -S.D[k][ca];
      store(SEC_UP + idx, ca, OFF_P + k, v);
      store(SEC_PU + idx, OFF_P + k, ca, v);
    }
  }
}

// (row, col) of entry e -- the symbolic phase builds the scatter map from this (h1h1.cu: entry_order)
inline void entry_rowcol(int e, int* li, int* lj) {
  if (e < SEC_UP) {
    const int cd = e / 729, ab = e % 729;
    *li = (cd / 3) * 27 + ab / 27;
    *lj = (cd % 3) * 27 + ab % 27;
  } else if (e < SEC_PU) {
    *li = (e - SEC_UP) / 4;
    *lj = OFF_P + (e - SEC_UP) % 4;
  } else if (e < SEC_UF) {
    *li = OFF_P + (e - SEC_PU) % 4;
    *lj = (e - SEC_PU) / 4;
  } else if (e < SEC_FU) {
    const int r = e - SEC_UF, c = r / (27 * NF), al = r % (27 * NF);
    *li = c * 27 + al / NF;
    *lj = OFF_F + al % NF;
  } else if (e < SEC_FF) {
    const int r = e - SEC_FU, c = r / (27 * NF), al = r % (27 * NF);
    *li = OFF_F + al % NF;
    *lj = c * 27 + al / NF;
  } else {
    *li = OFF_F + (e - SEC_FF) / NF;
    *lj = OFF_F + (e - SEC_FF) % NF;
  }
}

// ------------------------------------------------------------------ residual phases
// phase R4: p and grad phi at the points, the projection right-hand side and Mp^{-1}
template <bool ZU>
MHD_HD void phase_res_points(Shared& S, int tid, int nt) {
  for (int idx = tid; idx < NQ * 4; idx += nt) {
    const int q = idx / 4, r = idx % 4;
    double s = 0.0;
    if (r == 0) {
      for (int k = 0; k < 4; k++) s += S.pp[q][k] * S.st[OFF_P + k];
      S.pq[q] = s;
    } else {
      for (int l = 0; l < NF; l++) s += S.gF[q][r - 1][l] * S.st[OFF_F + l];
      S.gphi[q][r - 1] = s;
    }
  }
  if (ZU) {
    if (tid < 4) {
      double s = 0.0;
      for (int q = 0; q < NQ; q++) s += S.wdet[q] * S.pp[q][tid] * (S.gu[q][0] + S.gu[q][4] + S.gu[q][8]);
      S.rhs[tid] = s;
    }
    if (tid == nt - 1) {
      double Mp[16];
      for (int i = 0; i < 16; i++) {
        double s = 0.0;
        for (int q = 0; q < NQ; q++) s += S.wdet[q] * S.pp[q][i / 4] * S.pp[q][i % 4];
        Mp[i] = s;
      }
      invert4(Mp, S.Minv);
    }
  }
}

// phase R5: the integrand coefficients at the points
template <int CONV, bool ZU>
MHD_HD void phase_res_coefficients(Shared& S, int tid, int nt, const Params& P) {
  for (int q = tid; q < NQ; q += nt) {
    const double w = S.wdet[q];
    const double* u = S.uq[q];
    const double* g = S.gu[q];
    const double* gp = S.gphi[q];
    const double divu = g[0] + g[4] + g[8];
    double proj = 0.0;
    if (ZU) {
      for (int k = 0; k < 4; k++) {
        const double ck = S.Minv[k * 4 + 0] * S.rhs[0] + S.Minv[k * 4 + 1] * S.rhs[1] + S.Minv[k * 4 + 2] * S.rhs[2] +
                          S.Minv[k * 4 + 3] * S.rhs[3];
        proj += ck * S.pp[q][k];
      }
    }
    const double uB[3] = {u[1] * P.B[2] - u[2] * P.B[1], u[2] * P.B[0] - u[0] * P.B[2], u[0] * P.B[1] - u[1] * P.B[0]};
    // B x (u x B) and B x grad phi
    const double BuB[3] = {P.B[1] * uB[2] - P.B[2] * uB[1], P.B[2] * uB[0] - P.B[0] * uB[2], P.B[0] * uB[1] - P.B[1] * uB[0]};
    const double Bgp[3] = {P.B[1] * gp[2] - P.B[2] * gp[1], P.B[2] * gp[0] - P.B[0] * gp[2], P.B[0] * gp[1] - P.B[1] * gp[0]};
    for (int c = 0; c < 3; c++) {
      double fu = P.gamma * (BuB[c] - Bgp[c]) - P.f[c];
      if (CONV != 0) fu += P.alpha * (u[0] * g[0 * 3 + c] + u[1] * g[1 * 3 + c] + u[2] * g[2 * 3 + c]);
      S.Fu[q][c] = w * fu;
      for (int d = 0; d < 3; d++)
        S.Gu[q][c * 3 + d] = w * (P.beta * g[d * 3 + c] + (c == d ? (ZU ? P.zeta_u * proj : 0.0) - S.pq[q] : 0.0));
      S.Gf[q][c] = w * (gp[c] - uB[c]);
    }
    S.Fp[q] = -w * divu;
  }
}

// phase R6: the rows; add(li, v)
template <class Add>
MHD_HD void phase_res_rows(const Shared& S, int tid, int nt, Add& add) {
  for (int i = tid; i < NLOC; i += nt) {
    double s = 0.0;
    if (i < NU) {
      const int c = i / 27, a = i % 27;
      for (int q = 0; q < NQ; q++)
        s += S.N[q][a] * S.Fu[q][c] + S.gU[q][0][a] * S.Gu[q][c * 3 + 0] + S.gU[q][1][a] * S.Gu[q][c * 3 + 1] +
             S.gU[q][2][a] * S.Gu[q][c * 3 + 2];
    } else if (i < OFF_F) {
      for (int q = 0; q < NQ; q++) s += S.pp[q][i - OFF_P] * S.Fp[q];
    } else {
      const int l = i - OFF_F;
      for (int q = 0; q < NQ; q++) s += S.gF[q][0][l] * S.Gf[q][0] + S.gF[q][1][l] * S.Gf[q][1] + S.gF[q][2][l] * S.Gf[q][2];
    }
    add(i, s);
  }
}

}  // namespace h1
}  // namespace mhd
