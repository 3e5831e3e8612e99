/* CPU ORACLE in C (test infrastructure + timed CPU baseline; never linked into the product library).
 *
 * Plain-C restatement of the reference hot path, the same algorithm as oracle/mhd_oracle.py:
 *   - per-cell, per-quadrature-point integration of jac_fluid_h1_hdiv / res_fluid_h1_hdiv
 *     (/root/reference/src/weakforms.jl:283-312, 255-281; conv :670; local projection :672-681)
 *   - insertion into an existing sorted CSR by searching each (i,j) in its row, as Gridap's add_entry! does on
 *     re-assembly (jacobian!(A,op,x); reached from /root/reference/src/main.jl:163,275)
 *   - CSR SpMV, dot, axpy as the FGMRES of /root/reference/src/Solvers/badia2024.jl:40 issues them.
 * Parity pinning: see the header of mhd_oracle.py (solution-level pins; entry-level parity is unpinned upstream).
 * The reference itself (Julia + Gridap) cannot be built or run in this environment.
 * OpenMP over cells / rows: `cores` in bench.py's cpu_baseline = omp_get_max_threads().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NQ 27
#define NLOC 129
#define OFF_P 81
#define OFF_J 85
#define OFF_F 121

typedef struct {
  double alpha, beta, gamma, sigma, zeta_u, zeta_j;
  double B[3], f[3], g[3];
  int32_t convection; /* 0 none, 1 picard, 2 newton */
} oracle_params_t;

typedef struct {
  const double *w, *geo_grad, *u_val, *u_grad, *p_val, *j_val, *j_div, *phi_val;
} oracle_tables_t;

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bench.py --impl reference: torchrun exports OMP_NUM_THREADS=1, the CPU arm uses every host core at every N */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* geometry + mapped bases at one quadrature point */
static void point_bases(const oracle_tables_t* T, int q, const double* X, const int8_t* sign, double* wq, double gN[27][3],
                        double psi[36][3], double dpsi[36]) {
  double J[3][3] = {{0}}, inv[3][3];
  for (int v = 0; v < 8; v++)
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < 3; k++) J[i][k] += X[v * 3 + i] * T->geo_grad[(q * 8 + v) * 3 + k];
  double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
         c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  inv[0][0] = c00 / det; inv[1][0] = c01 / det; inv[2][0] = c02 / det;
  inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
  inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
  inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
  inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
  inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
  inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  *wq = T->w[q] * fabs(det);
  for (int a = 0; a < 27; a++)
    for (int i = 0; i < 3; i++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += T->u_grad[(q * 27 + a) * 3 + k] * inv[k][i];
      gN[a][i] = s;
    }
  for (int m = 0; m < 36; m++) {
    double sg = (double)sign[m] / det;
    for (int i = 0; i < 3; i++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += J[i][k] * T->j_val[(q * 36 + m) * 3 + k];
      psi[m][i] = sg * s;
    }
    dpsi[m] = sg * T->j_div[q * 36 + m];
  }
}

static void cross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* dense 129x129 cell matrix (row-major), rows = test, cols = trial */
void oracle_cell_jacobian(const oracle_tables_t* T, const oracle_params_t* P, const double* X, const double* st,
                          const int8_t* sign, double* K) {
  memset(K, 0, sizeof(double) * NLOC * NLOC);
  double D[4][81], Mp[4][4];
  memset(D, 0, sizeof(D));
  memset(Mp, 0, sizeof(Mp));
  for (int q = 0; q < NQ; q++) {
    double w, gN[27][3], psi[36][3], dpsi[36];
    point_bases(T, q, X, sign, &w, gN, psi, dpsi);
    const double* N = T->u_val + q * 27;
    const double* Pp = T->p_val + q * 4;
    const double* Chi = T->phi_val + q * 8;
    double uq[3] = {0, 0, 0}, gu[3][3] = {{0}};
    if (P->convection > 0)
      for (int i = 0; i < 3; i++)
        for (int a = 0; a < 27; a++) {
          uq[i] += N[a] * st[i * 27 + a];
          for (int d = 0; d < 3; d++) gu[d][i] += gN[a][d] * st[i * 27 + a];
        }
    for (int a = 0; a < 27; a++)
      for (int b = 0; b < 27; b++) {
        double s = P->beta * (gN[a][0] * gN[b][0] + gN[a][1] * gN[b][1] + gN[a][2] * gN[b][2]);
        if (P->convection > 0) s += P->alpha * N[a] * (uq[0] * gN[b][0] + uq[1] * gN[b][1] + uq[2] * gN[b][2]);
        for (int c = 0; c < 3; c++) K[(c * 27 + a) * NLOC + c * 27 + b] += w * s;
        if (P->convection == 2)
          for (int c = 0; c < 3; c++)
            for (int d = 0; d < 3; d++) K[(c * 27 + a) * NLOC + d * 27 + b] += w * P->alpha * N[a] * N[b] * gu[d][c];
      }
    for (int c = 0; c < 3; c++)
      for (int a = 0; a < 27; a++)
        for (int k = 0; k < 4; k++) {
          double v = w * Pp[k] * gN[a][c];
          K[(c * 27 + a) * NLOC + OFF_P + k] -= v;
          K[(OFF_P + k) * NLOC + c * 27 + a] -= v;
          D[k][c * 27 + a] += v;
        }
    for (int k = 0; k < 4; k++)
      for (int l = 0; l < 4; l++) Mp[k][l] += w * Pp[k] * Pp[l];
    for (int m = 0; m < 36; m++) {
      double xb[3];
      cross(psi[m], P->B, xb);
      for (int c = 0; c < 3; c++)
        for (int a = 0; a < 27; a++) {
          K[(c * 27 + a) * NLOC + OFF_J + m] -= P->gamma * w * N[a] * xb[c];
          K[(OFF_J + m) * NLOC + c * 27 + a] += P->sigma * w * N[a] * xb[c];
        }
      for (int n = 0; n < 36; n++)
        K[(OFF_J + m) * NLOC + OFF_J + n] +=
            w * (psi[m][0] * psi[n][0] + psi[m][1] * psi[n][1] + psi[m][2] * psi[n][2] + P->zeta_j * dpsi[m] * dpsi[n]);
      for (int l = 0; l < 8; l++) {
        K[(OFF_J + m) * NLOC + OFF_F + l] -= P->sigma * w * Chi[l] * dpsi[m];
        K[(OFF_F + l) * NLOC + OFF_J + m] -= w * Chi[l] * dpsi[m];
      }
    }
  }
  if (P->zeta_u != 0.0) {
    /* zeta_u * D^T Mp^{-1} D : solve Mp E = D by Gauss-Jordan */
    double a[4][4 + 81];
    for (int i = 0; i < 4; i++) {
      for (int j = 0; j < 4; j++) a[i][j] = Mp[i][j];
      for (int j = 0; j < 81; j++) a[i][4 + j] = D[i][j];
    }
    for (int p = 0; p < 4; p++) {
      double ip = 1.0 / a[p][p];
      for (int j = 0; j < 85; j++) a[p][j] *= ip;
      for (int i = 0; i < 4; i++)
        if (i != p) {
          double f = a[i][p];
          for (int j = 0; j < 85; j++) a[i][j] -= f * a[p][j];
        }
    }
    for (int i = 0; i < 81; i++)
      for (int j = 0; j < 81; j++) {
        double s = 0;
        for (int k = 0; k < 4; k++) s += D[k][i] * a[k][4 + j];
        K[i * NLOC + j] += P->zeta_u * s;
      }
  }
}

void oracle_cell_residual(const oracle_tables_t* T, const oracle_params_t* P, const double* X, const double* st,
                          const int8_t* sign, double* R) {
  memset(R, 0, sizeof(double) * NLOC);
  double proj_coef[4] = {0, 0, 0, 0};
  if (P->zeta_u != 0.0) {
    double a[4][5];
    memset(a, 0, sizeof(a));
    for (int q = 0; q < NQ; q++) {
      double w, gN[27][3], psi[36][3], dpsi[36];
      point_bases(T, q, X, sign, &w, gN, psi, dpsi);
      const double* Pp = T->p_val + q * 4;
      double divu = 0;
      for (int i = 0; i < 3; i++)
        for (int b = 0; b < 27; b++) divu += gN[b][i] * st[i * 27 + b];
      for (int k = 0; k < 4; k++) {
        for (int l = 0; l < 4; l++) a[k][l] += w * Pp[k] * Pp[l];
        a[k][4] += w * Pp[k] * divu;
      }
    }
    for (int p = 0; p < 4; p++) {
      double ip = 1.0 / a[p][p];
      for (int j = 0; j < 5; j++) a[p][j] *= ip;
      for (int i = 0; i < 4; i++)
        if (i != p) {
          double f = a[i][p];
          for (int j = 0; j < 5; j++) a[i][j] -= f * a[p][j];
        }
    }
    for (int k = 0; k < 4; k++) proj_coef[k] = a[k][4];
  }
  for (int q = 0; q < NQ; q++) {
    double w, gN[27][3], psi[36][3], dpsi[36];
    point_bases(T, q, X, sign, &w, gN, psi, dpsi);
    const double* N = T->u_val + q * 27;
    const double* Pp = T->p_val + q * 4;
    const double* Chi = T->phi_val + q * 8;
    double uq[3] = {0, 0, 0}, gu[3][3] = {{0}}, jq[3] = {0, 0, 0}, pq = 0, divj = 0, fq = 0, proj = 0;
    for (int i = 0; i < 3; i++)
      for (int a = 0; a < 27; a++) {
        uq[i] += N[a] * st[i * 27 + a];
        for (int d = 0; d < 3; d++) gu[d][i] += gN[a][d] * st[i * 27 + a];
      }
    for (int k = 0; k < 4; k++) {
      pq += Pp[k] * st[OFF_P + k];
      proj += Pp[k] * proj_coef[k];
    }
    for (int m = 0; m < 36; m++) {
      for (int i = 0; i < 3; i++) jq[i] += psi[m][i] * st[OFF_J + m];
      divj += dpsi[m] * st[OFF_J + m];
    }
    for (int l = 0; l < 8; l++) fq += Chi[l] * st[OFF_F + l];
    double divu = gu[0][0] + gu[1][1] + gu[2][2];
    double jxB[3], uxB[3], conv[3] = {0, 0, 0};
    cross(jq, P->B, jxB);
    cross(uq, P->B, uxB);
    if (P->convection > 0)
      for (int c = 0; c < 3; c++)
        for (int d = 0; d < 3; d++) conv[c] += uq[d] * gu[d][c];
    for (int c = 0; c < 3; c++)
      for (int a = 0; a < 27; a++) {
        double s = P->beta * (gu[0][c] * gN[a][0] + gu[1][c] * gN[a][1] + gu[2][c] * gN[a][2]);
        s += P->alpha * N[a] * conv[c];
        s += (P->zeta_u * proj - pq) * gN[a][c];
        s -= P->gamma * N[a] * jxB[c];
        s -= N[a] * P->f[c];
        R[c * 27 + a] += w * s;
      }
    for (int k = 0; k < 4; k++) R[OFF_P + k] -= w * Pp[k] * divu;
    for (int m = 0; m < 36; m++) {
      double s = 0;
      for (int i = 0; i < 3; i++) s += psi[m][i] * (jq[i] - P->sigma * uxB[i] - P->g[i]);
      s += dpsi[m] * (P->zeta_j * divj - P->sigma * fq);
      R[OFF_J + m] += w * s;
    }
    for (int l = 0; l < 8; l++) R[OFF_F + l] -= w * Chi[l] * divj;
  }
}

static int touched(int li, int lj) {
  int fi = li < OFF_P ? 0 : (li < OFF_J ? 1 : (li < OFF_F ? 2 : 3));
  int fj = lj < OFF_P ? 0 : (lj < OFF_J ? 1 : (lj < OFF_F ? 2 : 3));
  static const int m[4][4] = {{1, 1, 1, 0}, {1, 0, 0, 0}, {1, 0, 1, 1}, {0, 0, 1, 0}};
  return m[fi][fj];
}

static void cell_state(const int32_t* g, const double* x, const double* dirv, double* st) {
  for (int i = 0; i < NLOC; i++) st[i] = g[i] >= 0 ? x[g[i]] : dirv[-g[i] - 1];
}

/* numeric assembly into an existing CSR (0-based, sorted). gids: [ncells*129], >=0 free id, <0 -(dirichlet idx+1).
 * cell range [c0,c1) so that bench.py can time a bounded sample. */
void oracle_assemble_jacobian(const oracle_tables_t* T, const oracle_params_t* P, int64_t c0, int64_t c1,
                              const double* coords, const int32_t* cell_nodes, const int32_t* gids, const int8_t* jsign,
                              const double* dirv, const double* x, const int64_t* rowptr, const int64_t* colval,
                              double* nzval) {
#pragma omp parallel
  {
    double* K = (double*)malloc(sizeof(double) * NLOC * NLOC);
#pragma omp for schedule(dynamic, 4)
    for (int64_t c = c0; c < c1; c++) {
      double X[24], st[NLOC];
      for (int v = 0; v < 8; v++)
        for (int i = 0; i < 3; i++) X[v * 3 + i] = coords[(int64_t)cell_nodes[c * 8 + v] * 3 + i];
      const int32_t* g = gids + c * NLOC;
      cell_state(g, x, dirv, st);
      oracle_cell_jacobian(T, P, X, st, jsign + c * 36, K);
      for (int i = 0; i < NLOC; i++) {
        if (g[i] < 0) continue;
        int64_t lo0 = rowptr[g[i]], hi0 = rowptr[g[i] + 1];
        for (int j = 0; j < NLOC; j++) {
          if (g[j] < 0 || !touched(i, j)) continue;
          int64_t lo = lo0, hi = hi0 - 1;
          while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (colval[mid] < g[j]) lo = mid + 1; else hi = mid;
          }
#pragma omp atomic
          nzval[lo] += K[i * NLOC + j];
        }
      }
    }
    free(K);
  }
}

void oracle_assemble_residual(const oracle_tables_t* T, const oracle_params_t* P, int64_t c0, int64_t c1,
                              const double* coords, const int32_t* cell_nodes, const int32_t* gids, const int8_t* jsign,
                              const double* dirv, const double* x, double* r) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t c = c0; c < c1; c++) {
    double X[24], st[NLOC], R[NLOC];
    for (int v = 0; v < 8; v++)
      for (int i = 0; i < 3; i++) X[v * 3 + i] = coords[(int64_t)cell_nodes[c * 8 + v] * 3 + i];
    const int32_t* g = gids + c * NLOC;
    cell_state(g, x, dirv, st);
    oracle_cell_residual(T, P, X, st, jsign + c * 36, R);
    for (int i = 0; i < NLOC; i++)
      if (g[i] >= 0) {
#pragma omp atomic
        r[g[i]] += R[i];
      }
  }
}

void oracle_spmv(int64_t nrows, const int64_t* rowptr, const int64_t* colval, const double* nzval, const double* x,
                 double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nrows; i++) {
    double s = 0;
    for (int64_t p = rowptr[i]; p < rowptr[i + 1]; p++) s += nzval[p] * x[colval[p]];
    y[i] = s;
  }
}

double oracle_dot(int64_t n, const double* x, const double* y) {
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}

void oracle_axpy(int64_t n, double a, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) y[i] += a * x[i];
}
