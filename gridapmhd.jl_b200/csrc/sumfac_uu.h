// The velocity-velocity (uu) block of a cell by SUM FACTORISATION -- building block for the next Jacobian kernels
// (DESIGN.md 7.1a).  NOT yet called by any kernel: the phases below are verified on the CPU (tests/emul/emul_sumfac.cpp,
// tests/test_host.py) and wait for a B200 to be timed against the panel products of assembly.cu / the register tiles of
// h1h1_cell.h.
//
// With the tensor-product Q2 basis N_a(q) = l_i(q1) l_j(q2) l_k(q3), a = (i,j,k), on the tensor Gauss rule q = q1 + 3 q2 + 9 q3
// every uu contribution (jac_fluid_h1_hdiv / jac_fluid_h1_h1, src/weakforms.jl:283-312,440-466) is a sum of terms
//     K_f[a][b] = sum_q  D^{m} N_a(q) D^{n} N_b(q) C_f(q)
// and is contracted one direction at a time (3 159 FMA per field instead of 19 683):
//     T1_f[kk'][q1,q2]   = sum_q3 P_z[kk'][q3] C_f[q1,q2,q3]
//     T2_f[jj'][kk'][q1] = sum_q2 P_y[jj'][q2] T1_f[kk'][q1,q2]
//     K_f [ii'][jj'][kk'] = sum_q1 P_x[ii'][q1] T2_f[jj'][kk'][q1]
// with P[ii'][q] = D^m l_i(q) D^n l_i'(q) (derivative flags per direction and side).  The 21 coefficient fields:
//     f = 0..8   mass type  N_a N_b M_cd          M_cd = w|det J| (gamma (|B|^2 d_cd - B_c B_d) [H1-H1 only] + alpha d_d u_c [Newton])
//     f = 9..17  stiffness  d_m N_a d_n N_b G^mn  G^mn = beta w|det J| sum_i Jinv[m][i] Jinv[n][i]   (reference derivatives)
//     f = 18..20 convection N_a d_n N_b U^n       U^n  = alpha w|det J| sum_i Jinv[n][i] u_i
// and the block is  K[(a,c),(b,d)] = K_{c*3+d}[a][b] + delta_cd sum_{f >= 9} K_f[a][b].
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MHD_SHD __host__ __device__ __forceinline__
#else
#define MHD_SHD inline
#endif

namespace mhd {
namespace sf {

constexpr int NFIELD = 21;

struct Tables {
  double P[4][9][3];   // P[2*m + n][3*i + i'][q] = D^m l_i(x_q) D^n l_i'(x_q), m, n in {0 (value), 1 (derivative)}
  int8_t ijk[27][3];   // Q2 node a -> (i, j, k)
};

struct Work {
  double C[NFIELD][27];    // coefficient fields at the points
  double T1[NFIELD][81];   // [kk' * 9 + q1 + 3 q2]
  double T2[NFIELD][243];  // [(jj' * 9 + kk') * 3 + q1]
};

// derivative flags (a side, b side) of field f in direction ax
MHD_SHD int pair_index(int f, int ax) {
  int m = 0, n = 0;
  if (f >= 9 && f < 18) {
    m = ((f - 9) / 3 == ax);
    n = ((f - 9) % 3 == ax);
  } else if (f >= 18) {
    n = (f - 18 == ax);
  }
  return 2 * m + n;
}

MHD_SHD void phase_stage1(Work& W, const Tables& T, int tid, int nt) {
  for (int it = tid; it < NFIELD * 81; it += nt) {
    const int f = it / 81, r = it % 81, kk = r / 9, q12 = r % 9;
    const double* p = T.P[pair_index(f, 2)][kk];
    W.T1[f][r] = p[0] * W.C[f][q12] + p[1] * W.C[f][q12 + 9] + p[2] * W.C[f][q12 + 18];
  }
}

MHD_SHD void phase_stage2(Work& W, const Tables& T, int tid, int nt) {
  for (int it = tid; it < NFIELD * 243; it += nt) {
    const int f = it / 243, r = it % 243, jj = r / 27, kk = (r / 3) % 9, q1 = r % 3;
    const double* p = T.P[pair_index(f, 1)][jj];
    const double* t1 = W.T1[f] + kk * 9 + q1;
    W.T2[f][r] = p[0] * t1[0] + p[1] * t1[3] + p[2] * t1[6];
  }
}

// store(a, b, c, d, value)
template <class Store>
MHD_SHD void phase_stage3(const Work& W, const Tables& T, int tid, int nt, Store& store) {
  for (int ab = tid; ab < 729; ab += nt) {
    const int a = ab / 27, b = ab % 27;
    const int ii = 3 * T.ijk[a][0] + T.ijk[b][0], jj = 3 * T.ijk[a][1] + T.ijk[b][1], kk = 3 * T.ijk[a][2] + T.ijk[b][2];
    const int t2 = (jj * 9 + kk) * 3;
    double val[9], s = 0.0;
    for (int f = 0; f < NFIELD; f++) {
      const double* p = T.P[pair_index(f, 0)][ii];
      const double* t = W.T2[f] + t2;
      const double v = p[0] * t[0] + p[1] * t[1] + p[2] * t[2];
      if (f < 9) val[f] = v;
      else s += v;
    }
    for (int c = 0; c < 3; c++)
      for (int d = 0; d < 3; d++) store(a, b, c, d, val[c * 3 + d] + (c == d ? s : 0.0));
  }
}

}  // namespace sf
}  // namespace mhd
