// "v7" H1-HDiv Jacobian: cell-independent tables of the fully sum-factorised kernel (hdiv7_cell.h, hdiv7.cu) and their
// DISCOVERY from the plain reference tables of mhd_tables_t.
//
// The C ABI stays basis-agnostic (tabulated values at the 27 Gauss points).  At operator creation the library checks
// numerically whether those tables are tensor products of 1-D factors on the tensor rule q = q0 + 3 q1 + 9 q2 -- true for
// Gridap's HEX elements of src/parameters.jl:436-441,521-525 (Q2 Lagrangian, RT1 with face/interior moments against tensor
// bases, Q1) -- by a rank-1 factorisation of every basis function, clustering of the 1-D factors and "grid completion"
// (the tables of every class are read off the functions that differ from function 0 in one direction only).  Every table
// is then re-synthesised from the factors and compared with the input (tolerance 1e-12 relative); operators whose tables
// fail the check keep the generic tensor-core kernel of assembly.cu.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

namespace mhd {
namespace h7 {

constexpr int NQ = 27;

// Cell-independent data of the kernel (lives in __constant__ memory on the device).
struct Tab7 {
  double w[27];
  double gg[27][24];     // geometry-map gradients [q][v*3+k]
  double pp[27][4];      // pressure basis (generic table)
  double LV[3][3][3];    // Q2 1-D values      [direction][class][q]
  double LD[3][3][3];    // Q2 1-D derivatives
  double RV[3][3][3][3]; // RT: [component k][rotated direction r: actual direction (k+r)%3][class (3 for r = 0, else 2)][q]
  double RD[3][3][3];    // RT: derivative of the r = 0 factor [k][class][q]
  double XV[3][2][3];    // Q1 (phi) 1-D values [direction][class][q]
  double Puu[3][4][9][3];  // Q2 pair tables [direction][2*(row derivative)+(col derivative)][3*i+i'][q]
  double Puj[3][9][3];     // [k][3*a+i][q] = LV[k][a][q] * RV[k][0][i][q]   (last contraction of the u-j blocks)
  uint8_t node_t[27];    // reference Q2 node -> tensor index i0 + 3 i1 + 9 i2
  uint8_t jdof_t[36];    // reference RT dof  -> 12 k + i_0 + 3 (i_1 + 2 i_2)   (rotated classes)
  uint8_t phi_t[8];      // reference Q1 dof  -> l0 + 2 l1 + 4 l2
  uint8_t t_phi[8];      // inverse of phi_t
  uint8_t pad_[1];
};

namespace detail {

inline double absd(double x) { return x < 0 ? -x : x; }

// t[q0 + 3 q1 + 9 q2] -> scale, three factors normalised to 1 at the arg-max; returns the deviation from rank 1
inline double rank1(const double* t, double* scale, double f[3][3], int p[3]) {
  int pm = 0;
  for (int q = 1; q < 27; q++)
    if (absd(t[q]) > absd(t[pm])) pm = q;
  p[0] = pm % 3; p[1] = (pm / 3) % 3; p[2] = pm / 9;
  const double s = t[pm];
  *scale = s;
  if (s == 0.0) return 1e300;
  for (int i = 0; i < 3; i++) {
    f[0][i] = t[i + 3 * p[1] + 9 * p[2]] / s;
    f[1][i] = t[p[0] + 3 * i + 9 * p[2]] / s;
    f[2][i] = t[p[0] + 3 * p[1] + 9 * i] / s;
  }
  double dev = 0.0;
  for (int q = 0; q < 27; q++) dev = fmax(dev, absd(t[q] - s * f[0][q % 3] * f[1][(q / 3) % 3] * f[2][q / 9]));
  return dev / absd(s);
}

// Structure of n scalar functions tab[q*ld + f*stride] on the tensor rule: class index per direction, class tables V[d][class][q]
// (scales folded in by grid completion), optional derivative tables D[d][class][q] from dtab[(q*ld + f*stride)*... ] given by a
// callback.  ncls_expected[d]: required number of classes per direction.  Returns false if the tables are not of that form.
struct ScalarStruct {
  int n = 0;
  int idx[36][3];
  int ncls[3];
  double V[3][3][3];
  double D[3][3][3];
};

template <class Val, class Der>
inline bool discover(int n, Val val /* (q,f) */, Der der /* (q,f,dir) or null-like */, bool has_der, const int ncls_expected[3],
                     ScalarStruct* S, double tol = 1e-11) {
  S->n = n;
  double sc[36], F[36][3][3];
  int P[36][3];
  double amax = 0.0;
  for (int f = 0; f < n; f++) {
    double t[27];
    for (int q = 0; q < 27; q++) { t[q] = val(q, f); amax = fmax(amax, absd(t[q])); }
    if (rank1(t, &sc[f], F[f], P[f]) > tol) return false;
  }
  // cluster the 1-D factors per direction (up to sign)
  double reps[3][3][3];
  for (int d = 0; d < 3; d++) {
    int nc = 0;
    for (int f = 0; f < n; f++) {
      int found = -1;
      for (int c = 0; c < nc && found < 0; c++) {
        double dp = 0.0, dm = 0.0;
        for (int q = 0; q < 3; q++) { dp = fmax(dp, absd(F[f][d][q] - reps[d][c][q])); dm = fmax(dm, absd(F[f][d][q] + reps[d][c][q])); }
        if (dp < 1e-9 || dm < 1e-9) found = c;
      }
      if (found < 0) {
        if (nc == 3) return false;
        for (int q = 0; q < 3; q++) reps[d][nc][q] = F[f][d][q];
        found = nc++;
      }
      S->idx[f][d] = found;
    }
    S->ncls[d] = nc;
    if (nc != ncls_expected[d]) return false;
  }
  if (n != S->ncls[0] * S->ncls[1] * S->ncls[2]) return false;
  // the functions fill the class grid exactly once
  int find[27];
  for (int i = 0; i < 27; i++) find[i] = -1;
  for (int f = 0; f < n; f++) {
    const int key = S->idx[f][0] + 3 * S->idx[f][1] + 9 * S->idx[f][2];
    if (find[key] >= 0) return false;
    find[key] = f;
  }
  // grid completion through function 0: a point q* where every factor of function 0 equals 1
  const int* b = S->idx[0];
  const int qs[3] = {P[0][0], P[0][1], P[0][2]};
  const double v0 = val(qs[0] + 3 * qs[1] + 9 * qs[2], 0);
  memset(S->V, 0, sizeof(S->V));
  memset(S->D, 0, sizeof(S->D));
  for (int d = 0; d < 3; d++)
    for (int c = 0; c < S->ncls[d]; c++) {
      int key[3] = {b[0], b[1], b[2]};
      key[d] = c;
      const int f = find[key[0] + 3 * key[1] + 9 * key[2]];
      for (int q = 0; q < 3; q++) {
        int qq[3] = {qs[0], qs[1], qs[2]};
        qq[d] = q;
        S->V[d][c][q] = val(qq[0] + 3 * qq[1] + 9 * qq[2], f) / (d > 0 ? v0 : 1.0);
      }
    }
  for (int f = 0; f < n; f++)
    for (int q = 0; q < 27; q++) {
      const double r = S->V[0][S->idx[f][0]][q % 3] * S->V[1][S->idx[f][1]][(q / 3) % 3] * S->V[2][S->idx[f][2]][q / 9];
      if (absd(r - val(q, f)) > tol * amax) return false;
    }
  if (has_der) {
    double dmax = 0.0;
    for (int f = 0; f < n; f++)
      for (int q = 0; q < 27; q++)
        for (int d = 0; d < 3; d++) dmax = fmax(dmax, absd(der(q, f, d)));
    for (int d = 0; d < 3; d++)
      for (int c = 0; c < S->ncls[d]; c++) {
        int key[3] = {b[0], b[1], b[2]};
        key[d] = c;
        const int f = find[key[0] + 3 * key[1] + 9 * key[2]];
        double others = 1.0;
        for (int e = 0; e < 3; e++)
          if (e != d) others *= S->V[e][S->idx[f][e]][qs[e]];
        for (int q = 0; q < 3; q++) {
          int qq[3] = {qs[0], qs[1], qs[2]};
          qq[d] = q;
          S->D[d][c][q] = der(qq[0] + 3 * qq[1] + 9 * qq[2], f, d) / others;
        }
      }
    for (int f = 0; f < n; f++)
      for (int q = 0; q < 27; q++)
        for (int d = 0; d < 3; d++) {
          const int qd[3] = {q % 3, (q / 3) % 3, q / 9};
          double r = 1.0;
          for (int e = 0; e < 3; e++) r *= (e == d ? S->D[e] : S->V[e])[S->idx[f][e]][qd[e]];
          if (absd(r - der(q, f, d)) > 10 * tol * dmax) return false;
        }
  }
  return true;
}

}  // namespace detail

// Builds the kernel tables from the plain reference tables (layouts of mhd_tables_t).  Returns false when the tables do not
// have the tensor-product structure the kernel needs (the caller keeps the generic kernel).
inline bool build_tab7(const double* w, const double* geo_grad /*[27][8][3]*/, const double* u_val /*[27][27]*/,
                       const double* u_grad /*[27][27][3]*/, const double* p_val /*[27][4]*/, const double* j_val /*[27][36][3]*/,
                       const double* j_div /*[27][36]*/, const double* phi_val /*[27][8]*/, Tab7* T) {
  using namespace detail;
  memset(T, 0, sizeof(Tab7));
  memcpy(T->w, w, sizeof(T->w));
  memcpy(T->gg, geo_grad, sizeof(T->gg));
  memcpy(T->pp, p_val, sizeof(T->pp));
  // ---- Q2
  ScalarStruct S;
  const int n333[3] = {3, 3, 3};
  if (!discover(27, [&](int q, int f) { return u_val[q * 27 + f]; }, [&](int q, int f, int d) { return u_grad[(q * 27 + f) * 3 + d]; },
                true, n333, &S))
    return false;
  memcpy(T->LV, S.V, sizeof(T->LV));
  memcpy(T->LD, S.D, sizeof(T->LD));
  for (int a = 0; a < 27; a++) T->node_t[a] = (uint8_t)(S.idx[a][0] + 3 * S.idx[a][1] + 9 * S.idx[a][2]);
  for (int d = 0; d < 3; d++)
    for (int m = 0; m < 2; m++)
      for (int n = 0; n < 2; n++)
        for (int i = 0; i < 3; i++)
          for (int i2 = 0; i2 < 3; i2++)
            for (int q = 0; q < 3; q++)
              T->Puu[d][2 * m + n][3 * i + i2][q] = (m ? T->LD : T->LV)[d][i][q] * (n ? T->LD : T->LV)[d][i2][q];
  // ---- RT: every function has one non-zero reference component k; the 12 functions of a component form a (3,2,2) grid
  int cnt[3] = {0, 0, 0}, members[3][12];
  for (int m = 0; m < 36; m++) {
    double mx[3] = {0, 0, 0};
    for (int q = 0; q < 27; q++)
      for (int k = 0; k < 3; k++) mx[k] = fmax(mx[k], absd(j_val[(q * 36 + m) * 3 + k]));
    int k = mx[1] > mx[0] ? 1 : 0;
    if (mx[2] > mx[k]) k = 2;
    for (int e = 0; e < 3; e++)
      if (e != k && mx[e] > 1e-12 * mx[k]) return false;
    if (cnt[k] == 12) return false;
    members[k][cnt[k]++] = m;
  }
  for (int k = 0; k < 3; k++) {
    if (cnt[k] != 12) return false;
    int ne[3] = {2, 2, 2};
    ne[k] = 3;
    const int* ms = members[k];
    if (!discover(12, [&](int q, int f) { return j_val[(q * 36 + ms[f]) * 3 + k]; },
                  [&](int q, int f, int d) { return d == k ? j_div[q * 36 + ms[f]] : 0.0; }, false, ne, &S))
      return false;
    for (int r = 0; r < 3; r++) {
      const int d = (k + r) % 3;
      for (int c = 0; c < S.ncls[d]; c++)
        for (int q = 0; q < 3; q++) T->RV[k][r][c][q] = S.V[d][c][q];
    }
    // derivative of the quadratic factor from the divergence table, through function 0 of the component
    {
      int P[3];
      double sc, F[3][3], t[27];
      for (int q = 0; q < 27; q++) t[q] = j_val[(q * 36 + ms[0]) * 3 + k];
      rank1(t, &sc, F, P);
      int find[27];
      for (int i = 0; i < 27; i++) find[i] = -1;
      for (int f = 0; f < 12; f++) find[S.idx[f][0] + 3 * S.idx[f][1] + 9 * S.idx[f][2]] = f;
      for (int c = 0; c < 3; c++) {
        int key[3] = {S.idx[0][0], S.idx[0][1], S.idx[0][2]};
        key[k] = c;
        const int f = find[key[0] + 3 * key[1] + 9 * key[2]];
        double others = 1.0;
        for (int e = 0; e < 3; e++)
          if (e != k) others *= S.V[e][S.idx[f][e]][P[e]];
        for (int q = 0; q < 3; q++) {
          int qq[3] = {P[0], P[1], P[2]};
          qq[k] = q;
          T->RD[k][c][q] = j_div[(qq[0] + 3 * qq[1] + 9 * qq[2]) * 36 + ms[f]] / others;
        }
      }
      double dmax = 0.0;
      for (int q = 0; q < 27 * 36; q++) dmax = fmax(dmax, absd(j_div[q]));
      for (int f = 0; f < 12; f++)
        for (int q = 0; q < 27; q++) {
          const int qd[3] = {q % 3, (q / 3) % 3, q / 9};
          double r = 1.0;
          for (int e = 0; e < 3; e++) r *= e == k ? T->RD[k][S.idx[f][e]][qd[e]] : S.V[e][S.idx[f][e]][qd[e]];
          if (absd(r - j_div[q * 36 + ms[f]]) > 1e-10 * dmax) return false;
        }
    }
    for (int f = 0; f < 12; f++) {
      const int i0 = S.idx[f][k], i1 = S.idx[f][(k + 1) % 3], i2 = S.idx[f][(k + 2) % 3];
      T->jdof_t[ms[f]] = (uint8_t)(12 * k + i0 + 3 * (i1 + 2 * i2));
    }
    for (int a = 0; a < 3; a++)
      for (int i = 0; i < 3; i++)
        for (int q = 0; q < 3; q++) T->Puj[k][3 * a + i][q] = T->LV[k][a][q] * T->RV[k][0][i][q];
  }
  // ---- Q1 (phi)
  const int n222[3] = {2, 2, 2};
  if (!discover(8, [&](int q, int f) { return phi_val[q * 8 + f]; }, [&](int, int, int) { return 0.0; }, false, n222, &S)) return false;
  for (int d = 0; d < 3; d++)
    for (int c = 0; c < 2; c++)
      for (int q = 0; q < 3; q++) T->XV[d][c][q] = S.V[d][c][q];
  for (int l = 0; l < 8; l++) {
    T->phi_t[l] = (uint8_t)(S.idx[l][0] + 2 * S.idx[l][1] + 4 * S.idx[l][2]);
    T->t_phi[T->phi_t[l]] = (uint8_t)l;
  }
  return true;
}

}  // namespace h7
}  // namespace mhd
