// "v7" Jacobian of the H1-HDiv formulation: jac_fluid_h1_hdiv / jac_solid_h1_hdiv (src/weakforms.jl:283-312, :327-338) with
// EVERY block evaluated by sum factorisation, written as barrier-separated PHASES of a cooperative thread array so that
// tests/emul/emul_hdiv7.cpp runs the same code on the CPU against the oracle (forward / reverse thread order, NaN-filled data).
//
// With tensor-product bases on the tensor Gauss rule (hdiv7_tables.h) every block is a sum of terms
//     K[a][b] = sum_q F(q) a(q) b(q),   a = a_0(q_d0) a_1(q_d1) a_2(q_d2),  b likewise,  F a per-cell coefficient field,
// contracted one direction at a time:  T1[(a2,b2)][q_d0,q_d1] -> T2[(a1,b1)][(a2,b2)][q_d0] -> K.   Per fluid cell ~0.13 M FMA
// instead of 0.43 M (panel products of assembly.cu) and operands of a few hundred bytes instead of 27 x 27 panels:
//   uu   21 fields (9 Newton alpha w d_d u_c | 9 stiffness beta w Jinv Jinv^T | 3 convection), merged into 9 + 4 arrays at stage 2
//   uj   9 fields E_ck = w/det (J[c+1][k] B[c+2] - J[c+2][k] B[c+1]) (Piola map and cross product folded into the field),
//        K_uj = -gamma sgn V, K_ju = +sigma sgn V^T from the same V
//   jj   6 fields w/det^2 (J^T J)_kk' over the symmetric half + 1 field zeta_j w/det^2 (div-div)
//   jphi 1 field w/det ;  up: 36 fields w pi_k Jinv[k'][c] contracted with one basis function (the result D also feeds the
//        zeta_u projection term zeta_u D^T Mp^-1 D)
// The values are staged in shared memory in DESTINATION order (permuted local numbering of assembly.cu: every field sorted by
// global id) chunk by chunk -- uu rows of component 0 | 1 | 2, uj, ju, the rest -- and swept out 32 consecutive map entries per
// instruction while the next chunk is being computed (two staging buffers).
#pragma once
#include <stdint.h>

#include "hdiv7_tables.h"

// d x / d xi = sum_v X_v (x) grad N_v cancels |X| / h leading digits (1e4 on the boundary-layer cells of the Hunt meshes), so
// its ROUNDING decides the 13th digit of every matrix entry.  The reference rounds the product and the sum separately (Julia
// never contracts a * b + c on its own; oracle/mhd_oracle.py does the same): do that here too instead of the FMA nvcc would
// emit -- with the FMA the device sits 3e-13 (cfg2) .. 1.1e-12 (256 x 128 mesh) from the oracle, without it 4e-14.
#ifdef __CUDA_ARCH__
#define MHD_MULADD_REF(a, b, c) __dadd_rn(__dmul_rn((a), (b)), (c))
#else
#define MHD_MULADD_REF(a, b, c) ((a) * (b) + (c))
#endif
#ifdef __CUDACC__
#define MHD_7HD __host__ __device__ __forceinline__
#define MHD_7UNROLL _Pragma("unroll")
#define MHD_7NOUNROLL _Pragma("unroll 1")
#else
#define MHD_7HD inline
#define MHD_7UNROLL
#define MHD_7NOUNROLL
#endif

namespace mhd {
namespace h7 {

constexpr int NU = 81, NP = 4, NJ = 36, NF = 8;
constexpr int OFF_P = 81, OFF_J = 85, OFF_F = 121, NLOC = 129;

// ---- enumeration of the touched entries = order of this kernel's u16 scatter map (chunks padded to multiples of 32)
constexpr int pad32(int n) { return (n + 31) / 32 * 32; }
constexpr int CH_UU = 27 * 81, CH_UU_PAD = pad32(CH_UU);         // rows (c, slot) x cols (slot, d), one chunk per row component c
constexpr int CH_UJ = 81 * 36, CH_UJ_PAD = pad32(CH_UJ);
constexpr int E_UU = 0, E_UJ = 3 * CH_UU_PAD, E_JU = E_UJ + CH_UJ_PAD, E_REST = E_JU + CH_UJ_PAD;
constexpr int R_JJ = 0, R_JF = R_JJ + NJ * NJ, R_FJ = R_JF + NJ * NF, R_UP = R_FJ + NF * NJ, R_PU = R_UP + NU * NP, CH_REST = R_PU + NP * NU;
constexpr int CH_REST_PAD = pad32(CH_REST);
constexpr int NENT = E_REST + CH_REST_PAD;  // 15 040 codes per cell, 14 913 of them entries
constexpr int BUF = CH_UJ_PAD;                 // doubles per staging buffer

// (row slot, col slot) of map entry e in the permuted local numbering (u: c*27 + slot | p | j: 85 + slot | phi); false: padding
inline bool entry_rowcol(int e, int* li, int* lj) {
  auto ucol = [](int col) { return (col % 3) * 27 + col / 3; };
  if (e < E_UJ) {
    const int c = e / CH_UU_PAD, i = e % CH_UU_PAD;
    if (i >= CH_UU) return false;
    *li = c * 27 + i / 81;
    *lj = ucol(i % 81);
  } else if (e < E_JU) {
    const int i = e - E_UJ;
    if (i >= CH_UJ) return false;
    *li = i / 36;
    *lj = OFF_J + i % 36;
  } else if (e < E_REST) {
    const int i = e - E_JU;
    if (i >= CH_UJ) return false;
    *li = OFF_J + i / 81;
    *lj = ucol(i % 81);
  } else {
    const int i = e - E_REST;
    if (i >= CH_REST) return false;
    if (i < R_JF) { *li = OFF_J + i / 36; *lj = OFF_J + i % 36; }
    else if (i < R_FJ) { *li = OFF_J + (i - R_JF) / 8; *lj = OFF_F + (i - R_JF) % 8; }
    else if (i < R_UP) { *li = OFF_F + (i - R_FJ) / 36; *lj = OFF_J + (i - R_FJ) % 36; }
    else if (i < R_PU) { *li = (i - R_UP) / 4; *lj = OFF_P + (i - R_UP) % 4; }
    else { *li = OFF_P + (i - R_PU) / 81; *lj = ucol((i - R_PU) % 81); }
  }
  return true;
}

struct Params {
  double alpha, beta, gamma, sigma, zeta_u, zeta_j;
  double B[3];
  double f[3], g[3];  // body forces of the u and j equations (residual only)
};

// small cell-independent tables: SmallDyn = what is looked up with thread-dependent indices (kernel: shared memory, copied once
// per CTA from Tab7); Small7 = SmallDyn + the pair tables, which are only indexed at compile time (kernel: constant memory)
struct SmallDyn {
  double LV[3][3][3], LD[3][3][3];
  double RV[3][3][3][3], RD[3][3][3];
  double XV[3][2][3];
  double pp[27][4];
  uint8_t node_t[27], jdof_t[36], t_phi[8], pad_[1];
};
struct Small7 : SmallDyn {
  double Puu[3][4][9][3];
  double Puj[3][9][3];
};
MHD_7HD void small_from_tab(SmallDyn& s, const Tab7& T, int tid, int nt) {
  for (int i = tid; i < 27; i += nt) { (&s.LV[0][0][0])[i] = (&T.LV[0][0][0])[i]; (&s.LD[0][0][0])[i] = (&T.LD[0][0][0])[i]; (&s.RD[0][0][0])[i] = (&T.RD[0][0][0])[i]; }
  for (int i = tid; i < 81; i += nt) (&s.RV[0][0][0][0])[i] = (&T.RV[0][0][0][0])[i];
  for (int i = tid; i < 18; i += nt) (&s.XV[0][0][0])[i] = (&T.XV[0][0][0])[i];
  for (int i = tid; i < 108; i += nt) (&s.pp[0][0])[i] = (&T.pp[0][0])[i];
  for (int i = tid; i < 27; i += nt) s.node_t[i] = T.node_t[i];
  for (int i = tid; i < 36; i += nt) s.jdof_t[i] = T.jdof_t[i];
  for (int i = tid; i < 8; i += nt) s.t_phi[i] = T.t_phi[i];
}
MHD_7HD void small_from_tab(Small7& s, const Tab7& T, int tid, int nt) {
  small_from_tab(static_cast<SmallDyn&>(s), T, tid, nt);
  for (int i = tid; i < 81; i += nt) (&s.Puj[0][0][0])[i] = (&T.Puj[0][0][0])[i];
  for (int i = tid; i < 324; i += nt) (&s.Puu[0][0][0][0])[i] = (&T.Puu[0][0][0][0])[i];
}

// ---- field / intermediate layouts (doubles)
constexpr int NFIELD = 74;
constexpr int FO_UU = 0, FO_UJ = 21, FO_JJ = 30, FO_DD = 36, FO_JF = 37, FO_UP = 38;
constexpr int T1_UU = 0, T1_UJ = T1_UU + 21 * 81, T1_JA = T1_UJ + 9 * 54, T1_JB = T1_JA + 6 * 36, T1_JF = T1_JB + 6 * 36,
              T1_UP = T1_JF + 3 * 36, T1_END = T1_UP + 36 * 27;  // 3699
constexpr int T2_N = 0, T2_B = T2_N + 9 * 243, T2_UJ = T2_B + 4 * 243, T2_JA = T2_UJ + 9 * 108, T2_JB = T2_JA + 6 * 48,
              T2_JF = T2_JB + 6 * 72, T2_END = T2_JF + 3 * 48;  // 4995
constexpr int R1_DE = BUF;  // D and E live behind staging buffer 0 inside region 1 (T1 is dead by then)
static_assert(R1_DE + 2 * 324 <= T1_END, "D/E do not fit behind staging buffer 0");
constexpr int R3_T2UP = NFIELD * 27;  // 1998
constexpr int R3_END = R3_T2UP + 36 * 27;  // 2970
static_assert(R3_END >= BUF, "staging buffer 1 does not fit region 3");

// residual scratch (Cell7::rs) and the residual's intermediates inside r3[0, 1998) (free between stage 1 and chunk_uu(1))
constexpr int RS_JH = 0, RS_DJ = 81, RS_P = 162, RS_F = 189, RS_DIVU = 216, RS_D = 243, RS_PST = 247, RS_FST = 251, RS_JT = 259, RS_END = 296;
constexpr int RF_G = 0, RF_H = 243, RF_V = 324, RF_S = 405, RF_DJ = 432, RR_A = 459, RR_JP = 783, RR_B = 855, RR_END = 1179;
static_assert(RR_END <= NFIELD * 27, "residual intermediates must stay below T2up in r3");

struct Cell7 {
  double r1[T1_END];   // stage-1 results; later staging buffer 0 [0,BUF) | D [4][81] | E [4][81]
  double r2[T2_END];   // stage-2 results; its head first holds the temporaries of the point evaluation
  double r3[R3_END];   // coefficient fields [74][27] | T2up [36][27]; later staging buffer 1
  double J[27][9];     // J[q][i*3+k] = d x_i / d xi_k
  double invJ[27][9];  // invJ[q][k*3+i] = d xi_k / d x_i
  double W[27];        // w |det J|
  double idet[27];     // 1 / det J
  double X[24];
  double ut[81];       // velocity dofs in tensor order [c][i0 + 3 i1 + 9 i2]
  double uq[27][3];
  double gur[27][9];   // reference gradient of u: gur[q][k*3+c]
  double Mp[16], Minv[16];
  double rs[RS_END];   // residual: point values of j^, div^ j^, p, phi, div u | d | p, phi, j state (offsets RS_*)
  double sgn[36];      // RT sign flip by tensor id
  double sigma_cell, phi_sign;
  long long rowaddr[NLOC];  // permuted numbering: byte address of the first nnz of the row (kernel) / unused (emulation)
  int32_t gid[NLOC];
  uint8_t slot_u[32];  // tensor node index -> slot of the permuted numbering
  uint8_t slot_j[40];  // tensor RT id -> slot
};
MHD_7HD double* cellD(Cell7& S) { return S.r1 + R1_DE; }
MHD_7HD double* cellE(Cell7& S) { return S.r1 + R1_DE + 324; }

MHD_7HD int pow3(int d) { return d == 0 ? 1 : (d == 1 ? 3 : 9); }
MHD_7HD int qdir(int q, int d) { return d == 0 ? q % 3 : (d == 1 ? (q / 3) % 3 : q / 9); }  // index of point q in direction d
// derivative flags (row side, col side) of uu field f in direction ax (fields: 0..8 mass type, 9..17 stiffness (m,n), 18..20 convection n)
MHD_7HD int pat_uu(int f, int ax) {
  int m = 0, n = 0;
  if (f >= 9 && f < 18) { m = ((f - 9) / 3 == ax); n = ((f - 9) % 3 == ax); }
  else if (f >= 18) n = (f - 18 == ax);
  return 2 * m + n;
}
MHD_7HD void jjb_pair(int blk, int* k, int* kp) { *k = blk == 2 ? 1 : 0; *kp = blk == 0 ? 1 : 2; }

// ------------------------------------------------------------------ phase 0: gather (permuted ids, row starts, state, vertices)
// Every gathered VALUE is one item (u 81 | j 36 | p, phi 12 | vertex coordinates 24): load_gather fetches it from global memory,
// load_scatter files it in the cell data.  The kernel runs load_gather for the NEXT cell one barrier interval early (one value
// per thread, in a register) so that no global-memory latency is left at the top of a cell; phase_load is the plain sequence.
constexpr int LOAD_ITEMS = 153;
MHD_7HD double load_gather(int i, const double* coords, const int32_t* cell_nodes8, const int32_t* pgids129, const double* dir, const double* x,
                           bool need_u, bool with_res) {
  int32_t g;
  if (i < 81) {
    if (!need_u) return 0.0;
    g = pgids129[i];
  } else if (i < 129) {
    if (!with_res) return 0.0;
    g = pgids129[i < 117 ? OFF_J + (i - 81) : (i < 121 ? OFF_P + (i - 117) : OFF_F + (i - 121))];
  } else {
    const int k = i - 129;
    return coords[(long long)cell_nodes8[k / 3] * 3 + k % 3];
  }
  return g >= 0 ? x[g] : dir[-(long long)g - 1];
}
MHD_7HD void load_scatter(Cell7& S, const SmallDyn& C, int i, double v, const uint8_t* perm64, bool with_res) {
  if (i < 81) {
    const int c = i / 27, s = i % 27, t = C.node_t[perm64[s]];
    if (c == 0) S.slot_u[t] = (uint8_t)s;
    S.ut[c * 27 + t] = v;
  } else if (i < 117) {
    const int s = i - 81, pm = perm64[27 + s], tj = C.jdof_t[pm & 0x7F];
    S.slot_j[tj] = (uint8_t)s;
    const double sg = (pm & 0x80) ? -1.0 : 1.0;
    S.sgn[tj] = sg;
    if (with_res) S.rs[RS_JT + tj] = sg * v;
  } else if (i < 129) {
    if (with_res) S.rs[RS_PST + (i - 117)] = v;  // p (4) and phi (8) are not permuted
  } else {
    S.X[i - 129] = v;
  }
}
MHD_7HD void load_ids(Cell7& S, int tid, int nt, const int32_t* pgids129, const long long* rowstart129, const double* nzval, bool solid,
                      double sigma_cell, double sigma_fluid) {
  for (int i = tid; i < NLOC; i += nt) {
    S.gid[i] = pgids129[i];
    // device: absolute byte address of the first nnz of the row (dropped rows are never dereferenced: their codes are MAP_SKIP)
    S.rowaddr[i] = rowstart129 ? (long long)(nzval + rowstart129[i]) : -1;
  }
  if (tid == 0) {
    S.sigma_cell = solid ? sigma_cell : sigma_fluid;
    S.phi_sign = solid ? 1.0 : -1.0;
  }
}
MHD_7HD void phase_load(Cell7& S, const SmallDyn& C, int tid, int nt, const double* coords, const int32_t* cell_nodes8, const int32_t* pgids129,
                        const long long* rowstart129, const uint8_t* perm64, const double* dir, const double* x, bool need_u, bool solid,
                        double sigma_cell, double sigma_fluid, const double* nzval = nullptr, bool with_res = false) {
  load_ids(S, tid, nt, pgids129, rowstart129, nzval, solid, sigma_cell, sigma_fluid);
  for (int i = tid; i < LOAD_ITEMS; i += nt)
    load_scatter(S, C, i, load_gather(i, coords, cell_nodes8, pgids129, dir, x, need_u, with_res), perm64, with_res);
}

// ------------------------------------------------------------------ phase 1a: J at the points | first contraction of the point evaluation
template <int CONV>
MHD_7HD void phase_geom_a(Cell7& S, const SmallDyn& C, int tid, int nt, const double* gg /* Tab7::gg */) {
  const int n = 243 + (CONV != 0 ? 162 : 0);
  for (int it = tid; it < n; it += nt) {
    if (it < 243) {
      const int q = it / 9, i = (it / 3) % 3, k = it % 3;
      double s = 0.0;
      MHD_7UNROLL
      for (int v = 0; v < 8; v++) s = MHD_MULADD_REF(S.X[v * 3 + i], gg[q * 24 + v * 3 + k], s);
      S.J[q][i * 3 + k] = s;
    } else {
      // A1[var][c][q0 + 3 a12] = sum_a0 (var ? LD : LV)[0][a0][q0] ut[c][a0 + 3 a12]
      const int r = it - 243, var = r / 81, c = (r / 27) % 3, q0 = r % 3, a12 = (r / 3) % 9;
      const double* tb = var ? C.LD[0][0] : C.LV[0][0];
      const double* u = S.ut + c * 27 + 3 * a12;
      S.r2[(var * 3 + c) * 27 + q0 + 3 * a12] = tb[0 * 3 + q0] * u[0] + tb[1 * 3 + q0] * u[1] + tb[2 * 3 + q0] * u[2];
    }
  }
}

// ------------------------------------------------------------------ phase 1b: inverse / determinant | second contraction
template <int CONV>
MHD_7HD void phase_geom_b(Cell7& S, const SmallDyn& C, int tid, int nt, const double* w) {
  const int n = 27 + (CONV != 0 ? 243 : 0);
  for (int it = tid; it < n; it += nt) {
    if (it < 27) {
      const int q = it;
      const double* J = S.J[q];
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
      S.W[q] = w[q] * (det < 0.0 ? -det : det);
    } else {
      // Bv[var][c][q0 + 3 q1 + 9 a2] = sum_a1 T[a1][q1] A1[src][c][q0 + 3 (a1 + 3 a2)];  var 0: (v,v)  1: (d,v)  2: (v,d)
      const int r = it - 27, var = r / 81, c = (r / 27) % 3, q01 = r % 9, a2 = (r / 9) % 3, q0 = q01 % 3, q1 = q01 / 3;
      const double* tb = var == 2 ? C.LD[1][0] : C.LV[1][0];
      const double* a1 = S.r2 + ((var == 1 ? 1 : 0) * 3 + c) * 27 + q0 + 9 * a2;
      S.r2[162 + (var * 3 + c) * 27 + q01 + 9 * a2] = tb[0 * 3 + q1] * a1[0] + tb[1 * 3 + q1] * a1[3] + tb[2 * 3 + q1] * a1[6];
    }
  }
}

// ------------------------------------------------------------------ phase 2: u and its reference gradient at the points
MHD_7HD void phase_points(Cell7& S, const SmallDyn& C, int tid, int nt) {
  for (int it = tid; it < 324; it += nt) {
    const int kind = it / 81, c = (it / 27) % 3, q = it % 27, q01 = q % 9, q2 = q / 9;
    const double* tb = kind == 3 ? C.LD[2][0] : C.LV[2][0];
    const int var = kind == 1 ? 1 : (kind == 2 ? 2 : 0);
    const double* b = S.r2 + 162 + (var * 3 + c) * 27 + q01;
    const double v = tb[0 * 3 + q2] * b[0] + tb[1 * 3 + q2] * b[9] + tb[2 * 3 + q2] * b[18];
    if (kind == 0) S.uq[q][c] = v;
    else S.gur[q][(kind - 1) * 3 + c] = v;
  }
}

// ------------------------------------------------------------------ phase 3: coefficient fields
// Item = (group, point): a lane owns one quadrature point (27 of 32 lanes), loads its geometry once and evaluates all fields of
// its group from registers; with 256 threads warp w = group w, so the group switch is warp-uniform and every field index a
// compile-time constant.  Groups: 0 Newton (9) | 1 stiffness (9) | 2 convection (3) + jj (6) + div-div + j-phi | 3 uj (9) |
// 4..7 up, pressure function kp = group - 4 (9 each).
template <int CONV, bool ZJ>
MHD_7HD void phase_fields(Cell7& S, const SmallDyn& C, int tid, int nt, const Params& P) {
  double* F = S.r3;
  for (int it = tid; it < 8 * 32; it += nt) {
    const int grp = it >> 5, q = it & 31;
    if (q >= 27) continue;
    const double w = S.W[q];
    if (grp == 0) {
      if (CONV == 2) {
        double I[9], G[9];
        MHD_7UNROLL
        for (int i = 0; i < 9; i++) { I[i] = S.invJ[q][i]; G[i] = S.gur[q][i]; }
        const double aw = P.alpha * w;
        MHD_7UNROLL
        for (int c = 0; c < 3; c++)
          MHD_7UNROLL
          for (int d = 0; d < 3; d++) F[(c * 3 + d) * 27 + q] = aw * (I[0 * 3 + d] * G[0 * 3 + c] + I[1 * 3 + d] * G[1 * 3 + c] + I[2 * 3 + d] * G[2 * 3 + c]);
      } else {
        MHD_7UNROLL
        for (int f = 0; f < 9; f++) F[f * 27 + q] = 0.0;
      }
    } else if (grp == 1) {
      double I[9];
      MHD_7UNROLL
      for (int i = 0; i < 9; i++) I[i] = S.invJ[q][i];
      const double bw = P.beta * w;
      MHD_7UNROLL
      for (int m = 0; m < 3; m++)
        MHD_7UNROLL
        for (int n = m; n < 3; n++) {
          const double v = bw * (I[m * 3 + 0] * I[n * 3 + 0] + I[m * 3 + 1] * I[n * 3 + 1] + I[m * 3 + 2] * I[n * 3 + 2]);
          F[(9 + m * 3 + n) * 27 + q] = v;
          if (n != m) F[(9 + n * 3 + m) * 27 + q] = v;
        }
    } else if (grp == 2) {
      double J[9];
      MHD_7UNROLL
      for (int i = 0; i < 9; i++) J[i] = S.J[q][i];
      const double id = S.idet[q], wii = w * id * id;
      if (CONV != 0) {
        const double aw = P.alpha * w, u0 = S.uq[q][0], u1 = S.uq[q][1], u2 = S.uq[q][2];
        MHD_7UNROLL
        for (int n = 0; n < 3; n++) F[(18 + n) * 27 + q] = aw * (S.invJ[q][n * 3 + 0] * u0 + S.invJ[q][n * 3 + 1] * u1 + S.invJ[q][n * 3 + 2] * u2);
      } else {
        MHD_7UNROLL
        for (int n = 0; n < 3; n++) F[(18 + n) * 27 + q] = 0.0;
      }
      MHD_7UNROLL
      for (int b = 0; b < 6; b++) {
        const int k = b < 3 ? b : (b == 5 ? 1 : 0), kp = b < 3 ? b : (b == 3 ? 1 : 2);  // jjb_pair
        F[(FO_JJ + b) * 27 + q] = wii * (J[0 * 3 + k] * J[0 * 3 + kp] + J[1 * 3 + k] * J[1 * 3 + kp] + J[2 * 3 + k] * J[2 * 3 + kp]);
      }
      F[FO_DD * 27 + q] = ZJ ? P.zeta_j * wii : 0.0;
      F[FO_JF * 27 + q] = w * id;
    } else if (grp == 3) {
      double J[9];
      MHD_7UNROLL
      for (int i = 0; i < 9; i++) J[i] = S.J[q][i];
      const double wi = w * S.idet[q];
      MHD_7UNROLL
      for (int c = 0; c < 3; c++) {
        const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
        MHD_7UNROLL
        for (int k = 0; k < 3; k++) F[(FO_UJ + c * 3 + k) * 27 + q] = wi * (J[c1 * 3 + k] * P.B[c2] - J[c2 * 3 + k] * P.B[c1]);
      }
    } else {
      const int kp = grp - 4;
      const double wp = w * C.pp[q][kp];
      double* o = F + (FO_UP + kp * 9) * 27 + q;
      MHD_7UNROLL
      for (int c = 0; c < 3; c++)
        MHD_7UNROLL
        for (int kk = 0; kk < 3; kk++) o[(c * 3 + kk) * 27] = wp * S.invJ[q][kk * 3 + c];
    }
  }
}

// ------------------------------------------------------------------ phase 4: first contraction (direction d2) of every block
// One item per call (the kernel calls it for item = tid, tid + 256, tid + 512 without a loop: inside a loop ptxas hoists the
// constant-bank table operands of every case into registers on each call -- 850 instructions per cell in stage 1, 1 400 in
// stage 2 -- see chunk_uj_item).
constexpr int STAGE1_ITEMS = 189 + 81 + 54 + 54 + 27 + 324;
template <bool ZJ>
MHD_7HD void stage1_item(Cell7& S, const SmallDyn& C, const Small7& K, int it) {
  constexpr int N_UU = 189, N_UJ = N_UU + 81, N_JA = N_UJ + 54, N_JB = N_JA + 54, N_JF = N_JB + 27, N_UP = N_JF + 324;
  static_assert(N_UP == STAGE1_ITEMS, "stage 1 item count");
  const double* F = S.r3;
  if (it >= N_UP) return;
  {
    if (it < N_UU) {
      // items grouped by the direction-2 derivative pattern of their field so that the pattern switch is (nearly) warp-uniform
      // and every table operand a compile-time constant: pattern 0: Newton 0..8, stiffness 9 10 12 13, convection 18 19 |
      // 1: 11 14 20 | 2: 15 16 | 3: 17
      const int fi = it / 9, r = it % 9;
      int f, pat;
      if (fi < 9) { f = fi; pat = 0; }
      else if (fi < 13) { f = fi + (fi >= 11 ? 1 : 0); pat = 0; }
      else if (fi < 15) { f = fi + 5; pat = 0; }
      else if (fi < 18) { f = fi == 17 ? 20 : 11 + 3 * (fi - 15); pat = 1; }
      else if (fi < 20) { f = fi - 3; pat = 2; }
      else { f = 17; pat = 3; }
      const double x0 = F[f * 27 + r], x1 = F[f * 27 + r + 9], x2 = F[f * 27 + r + 18];
      double* o = S.r1 + T1_UU + f * 81 + r;
#define MHD_7S1(PAT_)                                                                                                   \
  MHD_7UNROLL                                                                                                           \
  for (int ab = 0; ab < 9; ab++) o[ab * 9] = K.Puu[2][PAT_][ab][0] * x0 + K.Puu[2][PAT_][ab][1] * x1 + K.Puu[2][PAT_][ab][2] * x2
      switch (pat) {
        case 0: MHD_7S1(0); break;
        case 1: MHD_7S1(1); break;
        case 2: MHD_7S1(2); break;
        default: MHD_7S1(3); break;
      }
#undef MHD_7S1
    } else if (it < N_UJ) {
      const int i = it - N_UU, inst = i / 9, r = i % 9, k = inst % 3;
      const int s0 = pow3(k), s1 = pow3((k + 1) % 3), d2 = (k + 2) % 3, s2 = pow3(d2);
      const double* f = F + (FO_UJ + inst) * 27 + (r % 3) * s0 + (r / 3) * s1;
      const double x0 = f[0], x1 = f[s2], x2 = f[2 * s2];
      double* o = S.r1 + T1_UJ + inst * 54 + r;
      MHD_7UNROLL
      for (int a = 0; a < 3; a++)
        MHD_7UNROLL
        for (int j = 0; j < 2; j++) {
          const double* la = C.LV[d2][a];
          const double* rb = C.RV[k][2][j];
          o[(a * 2 + j) * 9] = la[0] * rb[0] * x0 + la[1] * rb[1] * x1 + la[2] * rb[2] * x2;
        }
    } else if (it < N_JB) {
      const bool typeb = it >= N_JA;
      const int i = it - (typeb ? N_JA : N_UJ), blk = i / 18, kind = (i / 9) % 2, r = i % 9;
      if (kind == 1 && !ZJ) return;
      int k = blk, kp = blk;
      if (typeb) jjb_pair(blk, &k, &kp);
      const int d0 = k, d1 = typeb ? kp : (k + 1) % 3, d2 = 3 - d0 - d1;
      const int s0 = pow3(d0), s1 = pow3(d1), s2 = pow3(d2);
      const double* f = F + (kind ? FO_DD : FO_JJ + (typeb ? 3 : 0) + blk) * 27 + (r % 3) * s0 + (r / 3) * s1;
      const double x0 = f[0], x1 = f[s2], x2 = f[2 * s2];
      const double* ra = C.RV[k][(d2 - k + 3) % 3][0];
      const double* rb = C.RV[kp][(d2 - kp + 3) % 3][0];
      double* o = S.r1 + (typeb ? T1_JB : T1_JA) + (blk * 2 + kind) * 36 + r;
      MHD_7UNROLL
      for (int a = 0; a < 2; a++)
        MHD_7UNROLL
        for (int b = 0; b < 2; b++)
          o[(a * 2 + b) * 9] = ra[a * 3] * rb[b * 3] * x0 + ra[a * 3 + 1] * rb[b * 3 + 1] * x1 + ra[a * 3 + 2] * rb[b * 3 + 2] * x2;
    } else if (it < N_JF) {
      const int i = it - N_JB, k = i / 9, r = i % 9;
      const int s0 = pow3(k), s1 = pow3((k + 1) % 3), d2 = (k + 2) % 3, s2 = pow3(d2);
      const double* f = F + FO_JF * 27 + (r % 3) * s0 + (r / 3) * s1;
      const double x0 = f[0], x1 = f[s2], x2 = f[2 * s2];
      double* o = S.r1 + T1_JF + k * 36 + r;
      MHD_7UNROLL
      for (int a = 0; a < 2; a++)
        MHD_7UNROLL
        for (int l = 0; l < 2; l++) {
          const double* ra = C.RV[k][2][a];
          const double* xl = C.XV[d2][l];
          o[(a * 2 + l) * 9] = ra[0] * xl[0] * x0 + ra[1] * xl[1] * x1 + ra[2] * xl[2] * x2;
        }
    } else {
      const int i = it - N_JF, inst = i / 9, r = i % 9, kk = inst % 3;
      const double* f = F + (FO_UP + inst) * 27 + r;
      const double x0 = f[0], x1 = f[9], x2 = f[18];
      const double* tb = kk == 2 ? C.LD[2][0] : C.LV[2][0];
      double* o = S.r1 + T1_UP + inst * 27 + r;
      MHD_7UNROLL
      for (int a = 0; a < 3; a++) o[a * 9] = tb[a * 3] * x0 + tb[a * 3 + 1] * x1 + tb[a * 3 + 2] * x2;
    }
  }
}
template <bool ZJ>
MHD_7HD void phase_stage1(Cell7& S, const SmallDyn& C, const Small7& K, int tid, int nt) {
  for (int it = tid; it < STAGE1_ITEMS; it += nt) stage1_item<ZJ>(S, C, K, it);
}

// ------------------------------------------------------------------ phase 5: second contraction (direction d1)
constexpr int STAGE2_ITEMS = 243 + 108 + 162 + 72 + 72 + 36 + 324;
template <int CONV, bool ZJ>
MHD_7HD void stage2_item(Cell7& S, const SmallDyn& C, const Small7& K, int it) {
  constexpr int N_N = 243, N_B = N_N + 108, N_UJ = N_B + 162, N_JA = N_UJ + 72, N_JB = N_JA + 72, N_JF = N_JB + 36, N_UP = N_JF + 324;
  static_assert(N_UP == STAGE2_ITEMS, "stage 2 item count");
  const double* T1 = S.r1;
  if (it >= N_UP) return;
  {
    if (it < N_N) {
      if (CONV != 2) return;
      const int f = it / 27, p2 = (it / 3) % 9, q0 = it % 3;
      const double* x = T1 + T1_UU + f * 81 + p2 * 9 + q0;
      const double x0 = x[0], x1 = x[3], x2 = x[6];
      double* o = S.r2 + T2_N + f * 243 + p2 * 3 + q0;
      MHD_7UNROLL
      for (int ab = 0; ab < 9; ab++) o[ab * 27] = K.Puu[1][0][ab][0] * x0 + K.Puu[1][0][ab][1] * x1 + K.Puu[1][0][ab][2] * x2;
    } else if (it < N_B) {
      const int i = it - N_N, pc = i / 27, p2 = (i / 3) % 9, q0 = i % 3;
      double acc[9];
      MHD_7UNROLL
      for (int ab = 0; ab < 9; ab++) acc[ab] = 0.0;
      // The fields whose direction-0 derivative pattern is pc, grouped by their direction-1 pattern (stiffness f = 9 + 3 m + n has
      // pattern 2 (m == ax) + (n == ax) in direction ax, convection f = 18 + n has (n == ax)); fields of one group are summed
      // before the contraction.  The loop over the pattern is NOT unrolled (27 compile-time table operands per pattern).
      const double* xb = T1 + T1_UU + p2 * 9 + q0;
      // straight-line code, one block per direction-1 pattern (no loop: see stage1_item); f1 / f2 = fields of (pc, pattern)
#define MHD_7NB(PAT_, F1_EXPR, F2_EXPR)                                                                                   \
  do {                                                                                                                    \
    const int f1 = (F1_EXPR), f2 = (F2_EXPR);                                                                              \
    if (f1 >= 0) {                                                                                                        \
      double x0 = xb[f1 * 81], x1 = xb[f1 * 81 + 3], x2 = xb[f1 * 81 + 6];                                                  \
      if (CONV != 0 && f2 >= 0) { x0 += xb[f2 * 81]; x1 += xb[f2 * 81 + 3]; x2 += xb[f2 * 81 + 6]; }                        \
      MHD_7UNROLL                                                                                                         \
      for (int ab = 0; ab < 9; ab++) acc[ab] += K.Puu[1][PAT_][ab][0] * x0 + K.Puu[1][PAT_][ab][1] * x1 + K.Puu[1][PAT_][ab][2] * x2; \
    }                                                                                                                     \
  } while (0)
      MHD_7NB(0, pc == 3 ? 9 : (pc == 2 ? 11 : (pc == 1 ? 15 : 17)), pc == 1 ? 18 : (pc == 0 ? 20 : -1));
      MHD_7NB(1, pc == 2 ? 10 : (pc == 0 ? 16 : -1), pc == 0 ? 19 : -1);
      MHD_7NB(2, pc == 1 ? 12 : (pc == 0 ? 14 : -1), -1);
      MHD_7NB(3, pc == 0 ? 13 : -1, -1);
#undef MHD_7NB
      double* o = S.r2 + T2_B + pc * 243 + p2 * 3 + q0;
      MHD_7UNROLL
      for (int ab = 0; ab < 9; ab++) o[ab * 27] = acc[ab];
    } else if (it < N_UJ) {
      const int i = it - N_B, inst = i / 18, p2 = (i / 3) % 6, q0 = i % 3, k = inst % 3, d1 = (k + 1) % 3;
      const double* x = T1 + T1_UJ + inst * 54 + p2 * 9 + q0;
      const double x0 = x[0], x1 = x[3], x2 = x[6];
      double* o = S.r2 + T2_UJ + inst * 108 + p2 * 3 + q0;
      MHD_7UNROLL
      for (int a = 0; a < 3; a++)
        MHD_7UNROLL
        for (int j = 0; j < 2; j++) {
          const double* la = C.LV[d1][a];
          const double* rb = C.RV[k][1][j];
          o[(a * 2 + j) * 18] = la[0] * rb[0] * x0 + la[1] * rb[1] * x1 + la[2] * rb[2] * x2;
        }
    } else if (it < N_JA) {
      const int i = it - N_UJ, blk = i / 24, kind = (i / 12) % 2, p2 = (i / 3) % 4, q0 = i % 3, k = blk;
      if (kind == 1 && !ZJ) return;
      const double* x = T1 + T1_JA + (blk * 2 + kind) * 36 + p2 * 9 + q0;
      const double x0 = x[0], x1 = x[3], x2 = x[6];
      const double* r = C.RV[k][1][0];
      double* o = S.r2 + T2_JA + (blk * 2 + kind) * 48 + p2 * 3 + q0;
      MHD_7UNROLL
      for (int a = 0; a < 2; a++)
        MHD_7UNROLL
        for (int b = 0; b < 2; b++) o[(a * 2 + b) * 12] = r[a * 3] * r[b * 3] * x0 + r[a * 3 + 1] * r[b * 3 + 1] * x1 + r[a * 3 + 2] * r[b * 3 + 2] * x2;
    } else if (it < N_JB) {
      const int i = it - N_JA, blk = i / 24, kind = (i / 12) % 2, p2 = (i / 3) % 4, q0 = i % 3;
      if (kind == 1 && !ZJ) return;
      int k, kp;
      jjb_pair(blk, &k, &kp);
      const int d1 = kp;
      const double* x = T1 + T1_JB + (blk * 2 + kind) * 36 + p2 * 9 + q0;
      const double x0 = x[0], x1 = x[3], x2 = x[6];
      const double* ra = C.RV[k][(d1 - k + 3) % 3][0];                 // row: a linear factor of component k (2 classes)
      const double* rb = kind ? C.RD[kp][0] : C.RV[kp][0][0];          // col: the quadratic factor of component k' (3 classes)
      double* o = S.r2 + T2_JB + (blk * 2 + kind) * 72 + p2 * 3 + q0;
      MHD_7UNROLL
      for (int a = 0; a < 2; a++)
        MHD_7UNROLL
        for (int b = 0; b < 3; b++) o[(a * 3 + b) * 12] = ra[a * 3] * rb[b * 3] * x0 + ra[a * 3 + 1] * rb[b * 3 + 1] * x1 + ra[a * 3 + 2] * rb[b * 3 + 2] * x2;
    } else if (it < N_JF) {
      const int i = it - N_JB, k = i / 12, p2 = (i / 3) % 4, q0 = i % 3, d1 = (k + 1) % 3;
      const double* x = T1 + T1_JF + k * 36 + p2 * 9 + q0;
      const double x0 = x[0], x1 = x[3], x2 = x[6];
      double* o = S.r2 + T2_JF + k * 48 + p2 * 3 + q0;
      MHD_7UNROLL
      for (int a = 0; a < 2; a++)
        MHD_7UNROLL
        for (int l = 0; l < 2; l++) {
          const double* ra = C.RV[k][1][a];
          const double* xl = C.XV[d1][l];
          o[(a * 2 + l) * 12] = ra[0] * xl[0] * x0 + ra[1] * xl[1] * x1 + ra[2] * xl[2] * x2;
        }
    } else {
      const int i = it - N_JF, inst = i / 9, a2 = (i / 3) % 3, q0 = i % 3, kk = inst % 3;
      const double* x = T1 + T1_UP + inst * 27 + a2 * 9 + q0;
      const double x0 = x[0], x1 = x[3], x2 = x[6];
      const double* tb = kk == 1 ? C.LD[1][0] : C.LV[1][0];
      double* o = S.r3 + R3_T2UP + inst * 27 + a2 * 3 + q0;
      MHD_7UNROLL
      for (int a = 0; a < 3; a++) o[a * 9] = tb[a * 3] * x0 + tb[a * 3 + 1] * x1 + tb[a * 3 + 2] * x2;
    }
  }
}
template <int CONV, bool ZJ>
MHD_7HD void phase_stage2(Cell7& S, const SmallDyn& C, const Small7& K, int tid, int nt) {
  for (int it = tid; it < STAGE2_ITEMS; it += nt) stage2_item<CONV, ZJ>(S, C, K, it);
}

// ------------------------------------------------------------------ phase 5b: D[k][(c, tensor node)] = int pi_k d_c N, pressure mass matrix
// pressure mass matrix (zeta_u projection): 16 entries
MHD_7HD void phase_Mp(Cell7& S, const SmallDyn& C, int tid, int nt) {
  for (int i = tid; i < 16; i += nt) {
    double s = 0.0;
    for (int q = 0; q < 27; q++) s += S.W[q] * C.pp[q][i / 4] * C.pp[q][i % 4];
    S.Mp[i] = s;
  }
}

template <bool ZU>
MHD_7HD void phase_D(Cell7& S, const SmallDyn& C, const Small7& K, int tid, int nt) {
  double* D = cellD(S);
  for (int it = tid; it < 108; it += nt) {
    {
      const int kc = it / 9, a1 = (it / 3) % 3, a2 = it % 3, kp = kc / 3, c = kc % 3;
      double acc[3] = {0.0, 0.0, 0.0};
      MHD_7UNROLL
      for (int kk = 0; kk < 3; kk++) {
        const double* x = S.r3 + R3_T2UP + (kc * 3 + kk) * 27 + a1 * 9 + a2 * 3;
        MHD_7UNROLL
        for (int a0 = 0; a0 < 3; a0++)
          acc[a0] += (kk == 0 ? K.LD[0][a0][0] : K.LV[0][a0][0]) * x[0] + (kk == 0 ? K.LD[0][a0][1] : K.LV[0][a0][1]) * x[1] +
                     (kk == 0 ? K.LD[0][a0][2] : K.LV[0][a0][2]) * x[2];
      }
      MHD_7UNROLL
      for (int a0 = 0; a0 < 3; a0++) D[kp * 81 + c * 27 + a0 + 3 * a1 + 9 * a2] = acc[a0];
    }
  }
}

MHD_7HD void invert4(const double* M, double* out) {
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

MHD_7HD void phase_Minv(Cell7& S, int tid, int nt) {
  if (tid == nt - 1) invert4(S.Mp, S.Minv);
}
MHD_7HD void phase_E(Cell7& S, int tid, int nt, double zeta_u) {
  const double* D = cellD(S);
  double* E = cellE(S);
  for (int it = tid; it < 324; it += nt) {
    const int k = it / 81, ca = it % 81;
    E[it] = zeta_u * (S.Minv[k * 4 + 0] * D[ca] + S.Minv[k * 4 + 1] * D[81 + ca] + S.Minv[k * 4 + 2] * D[162 + ca] + S.Minv[k * 4 + 3] * D[243 + ca]);
  }
}

// ------------------------------------------------------------------ chunks: last contraction (direction d0) -> staging buffer
// uu rows of component c: buf[slot(a) * 81 + 3 slot(b) + d]
// one item (of 243) of a uu chunk; the kernel calls it directly with item = tid (see chunk_uj_item)
template <int CONV, bool ZU>
MHD_7HD void chunk_uu_item(Cell7& S, const SmallDyn& C, const Small7& K, int it, int c, double* buf) {
  const double* D = cellD(S);
  const double* E = cellE(S);
  {
    const int a0 = it / 81, r = it - 81 * a0, p1 = r / 9, p2 = r - 9 * p1, bi = p1 * 27 + p2 * 3;
    const int a1 = p1 / 3, b1 = p1 - 3 * a1, a2 = p2 / 3, b2 = p2 - 3 * a2;
    const int ta = a0 + 3 * a1 + 9 * a2, tb12 = 3 * b1 + 9 * b2;
    // Puu[0][2 m + n][3 a0 + b0][q] = (m ? LD : LV)[0][a0][q] * (n ? LD : LV)[0][b0][q]: fold the row factor (thread-dependent
    // class a0) into the stage-2 data once, the column factor (unrolled b0) is a compile-time table operand
    double rv[3], rd[3];
    MHD_7UNROLL
    for (int q = 0; q < 3; q++) { rv[q] = C.LV[0][a0][q]; rd[q] = C.LD[0][a0][q]; }
    double tb[4][3], tn[3][3];
    MHD_7UNROLL
    for (int pc = 0; pc < 4; pc++)
      MHD_7UNROLL
      for (int q = 0; q < 3; q++) tb[pc][q] = ((pc & 2) ? rd[q] : rv[q]) * S.r2[T2_B + pc * 243 + bi + q];
    if (CONV == 2) {
      MHD_7UNROLL
      for (int d = 0; d < 3; d++)
        MHD_7UNROLL
        for (int q = 0; q < 3; q++) tn[d][q] = rv[q] * S.r2[T2_N + (c * 3 + d) * 243 + bi + q];
    }
    const int rowoff = S.slot_u[ta] * 81;
    MHD_7UNROLL
    for (int b0 = 0; b0 < 3; b0++) {
      const int tbn = b0 + tb12;
      double base = 0.0;
      MHD_7UNROLL
      for (int pc = 0; pc < 4; pc++)
        MHD_7UNROLL
        for (int q = 0; q < 3; q++) base += ((pc & 1) ? K.LD[0][b0][q] : K.LV[0][b0][q]) * tb[pc][q];
      const int col = 3 * S.slot_u[tbn];
      MHD_7UNROLL
      for (int d = 0; d < 3; d++) {
        double v = d == c ? base : 0.0;
        if (CONV == 2) v += K.LV[0][b0][0] * tn[d][0] + K.LV[0][b0][1] * tn[d][1] + K.LV[0][b0][2] * tn[d][2];
        if (ZU) v += D[c * 27 + ta] * E[d * 27 + tbn] + D[81 + c * 27 + ta] * E[81 + d * 27 + tbn] + D[162 + c * 27 + ta] * E[162 + d * 27 + tbn] +
                     D[243 + c * 27 + ta] * E[243 + d * 27 + tbn];
        buf[rowoff + col + d] = v;
      }
    }
  }
}
template <int CONV, bool ZU>
MHD_7HD void chunk_uu(Cell7& S, const SmallDyn& C, const Small7& K, int tid, int nt, int c, double* buf) {
  for (int it = tid; it < 243; it += nt) chunk_uu_item<CONV, ZU>(S, C, K, it, c, buf);
}

// uj (JU = false): buf[(c*27 + slot(a)) * 36 + slot(m)] = -gamma sgn V ;  ju (JU = true): buf[slot(m) * 81 + 3 slot(a) + c] = +sigma sgn V
// (JU is a run-time flag: one copy of the code, the kernel is instruction-cache bound otherwise)
template <int KD>
MHD_7HD void chunk_uj_k(Cell7& S, const Small7& K, const Params& P, double* buf, int c, int p1, int p2, bool JU) {
  constexpr int d1 = (KD + 1) % 3, d2 = (KD + 2) % 3;
  constexpr int s0 = KD == 0 ? 1 : (KD == 1 ? 3 : 9), s1 = d1 == 0 ? 1 : (d1 == 1 ? 3 : 9), s2 = d2 == 0 ? 1 : (d2 == 1 ? 3 : 9);
  const double* x = S.r2 + T2_UJ + (c * 3 + KD) * 108 + p1 * 18 + p2 * 3;
  const int a1 = p1 >> 1, i1 = p1 & 1, a2 = p2 >> 1, i2 = p2 & 1;
  const int ta12 = a1 * s1 + a2 * s2, tj12 = 12 * KD + 3 * (i1 + 2 * i2);
  const double coef = JU ? P.sigma : -P.gamma;
  // destination index = base + su * SU + sj * SJ   (ju: slot_j * 81 + 3 slot_u + c | uj: (c * 27 + slot_u) * 36 + slot_j)
  const int SU = JU ? 3 : 36, SJ = JU ? 81 : 1, base = JU ? c : c * 972;
  double xs[3][3];  // sign, coefficient and stage-2 data folded together
  int oj[3];
  MHD_7UNROLL
  for (int i0 = 0; i0 < 3; i0++) {
    const double sg = coef * S.sgn[tj12 + i0];
    oj[i0] = base + S.slot_j[tj12 + i0] * SJ;
    MHD_7UNROLL
    for (int q = 0; q < 3; q++) xs[i0][q] = sg * x[q];
  }
  MHD_7UNROLL
  for (int a = 0; a < 3; a++) {
    const int ou = S.slot_u[ta12 + a * s0] * SU;
    MHD_7UNROLL
    for (int i0 = 0; i0 < 3; i0++)
      buf[ou + oj[i0]] = K.Puj[KD][a * 3 + i0][0] * xs[i0][0] + K.Puj[KD][a * 3 + i0][1] * xs[i0][1] + K.Puj[KD][a * 3 + i0][2] * xs[i0][2];
  }
}
// one item (of 324): direction-major order, so that the direction switch is warp-uniform except at two boundaries.  The kernel
// calls this for item = tid and, for tid < 68, item = tid + 256 -- NOT through the loop below: inside a loop ptxas hoists the
// 81 constant-bank table values of the three cases into registers on every call (a 110-instruction prologue).
MHD_7HD void chunk_uj_item(Cell7& S, const Small7& K, const Params& P, double* buf, int it, bool JU) {
  const int k = it / 108, r = it - 108 * k, c = r / 36, r2 = r - 36 * c, p1 = r2 / 6, p2 = r2 - 6 * p1;
  switch (k) {
    case 0: chunk_uj_k<0>(S, K, P, buf, c, p1, p2, JU); break;
    case 1: chunk_uj_k<1>(S, K, P, buf, c, p1, p2, JU); break;
    default: chunk_uj_k<2>(S, K, P, buf, c, p1, p2, JU); break;
  }
}
MHD_7HD void chunk_uj(Cell7& S, const Small7& K, int tid, int nt, const Params& P, double* buf, bool JU) {
  for (int it = tid; it < 324; it += nt) chunk_uj_item(S, K, P, buf, it, JU);
}

// the rest: jj | j-phi | phi-j | up | pu
template <bool ZJ>
MHD_7HD void chunk_rest(Cell7& S, const SmallDyn& C, int tid, int nt, double* buf) {
  constexpr int N_JA = 48, N_JB = N_JA + 72, N_JF = N_JB + 48, N_UP = N_JF + 324;
  const double* D = cellD(S);
  for (int it = tid; it < N_UP; it += nt) {
    if (it < N_JA) {
      const int blk = it / 16, p1 = (it / 4) % 4, p2 = it % 4, k = blk;
      const double* x = S.r2 + T2_JA + (blk * 2) * 48 + p1 * 12 + p2 * 3;
      const int tm12 = 12 * k + 3 * ((p1 / 2) + 2 * (p2 / 2)), tn12 = 12 * k + 3 * ((p1 % 2) + 2 * (p2 % 2));
      MHD_7UNROLL
      for (int i = 0; i < 3; i++)
        MHD_7UNROLL
        for (int j = 0; j < 3; j++) {
          const double* ra = C.RV[k][0][i];
          const double* rb = C.RV[k][0][j];
          double v = ra[0] * rb[0] * x[0] + ra[1] * rb[1] * x[1] + ra[2] * rb[2] * x[2];
          if (ZJ) {
            const double* da = C.RD[k][i];
            const double* db = C.RD[k][j];
            v += da[0] * db[0] * x[48] + da[1] * db[1] * x[49] + da[2] * db[2] * x[50];
          }
          const int tm = tm12 + i, tn = tn12 + j;
          buf[R_JJ + S.slot_j[tm] * 36 + S.slot_j[tn]] = S.sgn[tm] * S.sgn[tn] * v;
        }
    } else if (it < N_JB) {
      const int r = it - N_JA, blk = r / 24, p1 = (r / 4) % 6, p2 = r % 4;
      int k, kp;
      jjb_pair(blk, &k, &kp);
      const int d0 = k, d1 = kp, d2 = 3 - d0 - d1;
      const double* x = S.r2 + T2_JB + (blk * 2) * 72 + p1 * 12 + p2 * 3;
      const int i1 = p1 / 3, j1 = p1 % 3, i2 = p2 / 2, j2 = p2 % 2;   // row classes (i) | col classes (j) in directions d1, d2
      // tensor ids: class of rotated direction r of a component sits at weight (1, 3, 6)[r]
      auto wgt = [](int r) { return r == 0 ? 1 : (r == 1 ? 3 : 6); };
      const int tm12 = 12 * k + i1 * wgt((d1 - k + 3) % 3) + i2 * wgt((d2 - k + 3) % 3);
      const int tn12 = 12 * kp + j1 /* its quadratic factor */ + j2 * wgt((d2 - kp + 3) % 3);
      const int wn0 = wgt((d0 - kp + 3) % 3);
      MHD_7UNROLL
      for (int i = 0; i < 3; i++)
        MHD_7UNROLL
        for (int j = 0; j < 2; j++) {
          const double* ra = C.RV[k][0][i];
          const double* rb = C.RV[kp][(d0 - kp + 3) % 3][j];
          double v = ra[0] * rb[0] * x[0] + ra[1] * rb[1] * x[1] + ra[2] * rb[2] * x[2];
          if (ZJ) {
            const double* da = C.RD[k][i];
            v += da[0] * rb[0] * x[72] + da[1] * rb[1] * x[73] + da[2] * rb[2] * x[74];
          }
          const int tm = tm12 + i, tn = tn12 + j * wn0;
          const double vv = S.sgn[tm] * S.sgn[tn] * v;
          buf[R_JJ + S.slot_j[tm] * 36 + S.slot_j[tn]] = vv;
          buf[R_JJ + S.slot_j[tn] * 36 + S.slot_j[tm]] = vv;
        }
    } else if (it < N_JF) {
      const int r = it - N_JB, k = r / 16, p1 = (r / 4) % 4, p2 = r % 4, d1 = (k + 1) % 3, d2 = (k + 2) % 3;
      const double* x = S.r2 + T2_JF + k * 48 + p1 * 12 + p2 * 3;
      const int tm12 = 12 * k + 3 * ((p1 / 2) + 2 * (p2 / 2));
      const int lt12 = (p1 % 2) * (1 << d1) + (p2 % 2) * (1 << d2);
      MHD_7UNROLL
      for (int i = 0; i < 3; i++)
        MHD_7UNROLL
        for (int l0 = 0; l0 < 2; l0++) {
          const double* da = C.RD[k][i];
          const double* xl = C.XV[k][l0];
          const int tm = tm12 + i, l = C.t_phi[lt12 + l0 * (1 << k)];
          const double v = S.sgn[tm] * (da[0] * xl[0] * x[0] + da[1] * xl[1] * x[1] + da[2] * xl[2] * x[2]);
          buf[R_JF + S.slot_j[tm] * 8 + l] = -S.sigma_cell * v;
          buf[R_FJ + l * 36 + S.slot_j[tm]] = S.phi_sign * v;
        }
    } else {
      const int r = it - N_JF, kp = r / 81, ct = r % 81, c = ct / 27, su = S.slot_u[ct % 27];
      const double v = -D[r];
      buf[R_UP + (c * 27 + su) * 4 + kp] = v;
      buf[R_PU + kp * 81 + 3 * su + c] = v;
    }
  }
}

// =================================================================== residual (res_fluid_h1_hdiv / res_solid_h1_hdiv,
// src/weakforms.jl:255-281, :314-325) by sum factorisation, interleaved with the Jacobian phases of the same cell: every
// res_* phase shares a barrier interval with a Jacobian phase (see hdiv_v7.cu).  radd(local row (permuted numbering), value)
// accumulates into the global vector.
//   r_u[(a,c)] = sum_q [ G^k_c d^_k N_a + H_c N_a ],  G^k_c = w sum_i invJ[k][i] (beta (grad u)[i][c] + delta_ic (zeta_u Pi_p - p)),
//                                                      H_c   = w (alpha ((u.grad)u)_c - gamma (j x B)_c - f_c)
//   r_p[k]     = - sum_q w pi_k div u
//   r_j[m]     = sgn_m sum_q [ V_k psi^_m,k + S div^ psi^_m ],  V = w/det J^T (j - sigma u x B - g),  S = w/det (zeta_j div j - sigma_c phi)
//   r_phi[l]   = phi_sign sum_q w chi_l div j          (phi_sign = -1 fluid, +1 solid)

// interval of phase_geom_a: reference current j^_k, its divergence by component, p and phi at the points (state only)
MHD_7HD void res_pv(Cell7& S, const SmallDyn& C, int tid, int nt) {
  for (int it = tid; it < 216; it += nt) {
    if (it < 162) {
      const bool dv = it >= 81;
      const int r = dv ? it - 81 : it, k = r / 27, q = r % 27;
      const int q0 = qdir(q, k), q1 = qdir(q, (k + 1) % 3), q2 = qdir(q, (k + 2) % 3);
      const double* jt = S.rs + RS_JT + 12 * k;
      double s = 0.0;
      MHD_7UNROLL
      for (int i2 = 0; i2 < 2; i2++)
        MHD_7UNROLL
        for (int i1 = 0; i1 < 2; i1++) {
          const double f12 = C.RV[k][1][i1][q1] * C.RV[k][2][i2][q2];
          MHD_7UNROLL
          for (int i0 = 0; i0 < 3; i0++) s += jt[i0 + 3 * (i1 + 2 * i2)] * (dv ? C.RD[k][i0][q0] : C.RV[k][0][i0][q0]) * f12;
        }
      S.rs[(dv ? RS_DJ : RS_JH) + r] = s;
    } else if (it < 189) {
      const int q = it - 162;
      const double* ps = S.rs + RS_PST;
      S.rs[RS_P + q] = C.pp[q][0] * ps[0] + C.pp[q][1] * ps[1] + C.pp[q][2] * ps[2] + C.pp[q][3] * ps[3];
    } else {
      const int q = it - 189, q0 = q % 3, q1 = (q / 3) % 3, q2 = q / 9;
      double s = 0.0;
      MHD_7UNROLL
      for (int l = 0; l < 8; l++)
        s += S.rs[RS_FST + C.t_phi[l]] * C.XV[0][l & 1][q0] * C.XV[1][(l >> 1) & 1][q1] * C.XV[2][l >> 2][q2];
      S.rs[RS_F + q] = s;
    }
  }
}

// interval of phase_fields: div u at the points
MHD_7HD void res_divu(Cell7& S, int tid, int nt) {
  for (int q = tid; q < 27; q += nt) {
    const double* I = S.invJ[q];
    const double* G = S.gur[q];
    double s = 0.0;
    MHD_7UNROLL
    for (int k = 0; k < 3; k++)
      MHD_7UNROLL
      for (int i = 0; i < 3; i++) s += I[k * 3 + i] * G[k * 3 + i];
    S.rs[RS_DIVU + q] = s;
  }
}

// interval of phase_stage1: d_k = sum_q w pi_k div u ; r_p = -d
template <class RAdd>
MHD_7HD void res_d(Cell7& S, const SmallDyn& C, int tid, int nt, RAdd& radd) {
  for (int k = tid; k < 4; k += nt) {
    double s = 0.0;
    for (int q = 0; q < 27; q++) s += S.W[q] * C.pp[q][k] * S.rs[RS_DIVU + q];
    S.rs[RS_D + k] = s;
    radd(OFF_P + k, -s);
  }
}

// interval of phase_stage2: the coefficient fields of the residual -> r3[RF_*]
template <int CONV, bool ZU, bool ZJ>
MHD_7HD void res_fields(Cell7& S, const SmallDyn& C, int tid, int nt, const Params& P) {
  double* F = S.r3;
  for (int it = tid; it < 162; it += nt) {
    const int q = it % 27, grp = it / 27;
    const double* I = S.invJ[q];
    const double* J = S.J[q];
    const double* G = S.gur[q];
    const double w = S.W[q], id = S.idet[q];
    if (grp < 3) {  // G^k_c, c = grp
      const int c = grp;
      double s = -S.rs[RS_P + q];
      if (ZU) {
        double pr = 0.0;
        MHD_7UNROLL
        for (int k = 0; k < 4; k++)
          pr += C.pp[q][k] * (S.Minv[k * 4 + 0] * S.rs[RS_D] + S.Minv[k * 4 + 1] * S.rs[RS_D + 1] + S.Minv[k * 4 + 2] * S.rs[RS_D + 2] +
                              S.Minv[k * 4 + 3] * S.rs[RS_D + 3]);
        s += P.zeta_u * pr;
      }
      double t[3];  // beta (grad u)[i][c] + delta_ic s
      MHD_7UNROLL
      for (int i = 0; i < 3; i++) t[i] = P.beta * (I[0 * 3 + i] * G[0 * 3 + c] + I[1 * 3 + i] * G[1 * 3 + c] + I[2 * 3 + i] * G[2 * 3 + c]) + (i == c ? s : 0.0);
      MHD_7UNROLL
      for (int k = 0; k < 3; k++) F[RF_G + (c * 3 + k) * 27 + q] = w * (I[k * 3 + 0] * t[0] + I[k * 3 + 1] * t[1] + I[k * 3 + 2] * t[2]);
    } else if (grp == 3 || grp == 4) {  // H_c | V_k: both need the physical current
      double jp[3];
      MHD_7UNROLL
      for (int i = 0; i < 3; i++)
        jp[i] = id * (J[i * 3 + 0] * S.rs[RS_JH + q] + J[i * 3 + 1] * S.rs[RS_JH + 27 + q] + J[i * 3 + 2] * S.rs[RS_JH + 54 + q]);
      const double* u = S.uq[q];
      if (grp == 3) {
        MHD_7UNROLL
        for (int c = 0; c < 3; c++) {
          const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
          double h = -P.gamma * (jp[c1] * P.B[c2] - jp[c2] * P.B[c1]) - P.f[c];
          if (CONV != 0) {
            double cv = 0.0;  // ((u . grad) u)_c = sum_i u_i d_i u_c
            MHD_7UNROLL
            for (int i = 0; i < 3; i++) cv += u[i] * (I[0 * 3 + i] * G[0 * 3 + c] + I[1 * 3 + i] * G[1 * 3 + c] + I[2 * 3 + i] * G[2 * 3 + c]);
            h += P.alpha * cv;
          }
          F[RF_H + c * 27 + q] = w * h;
        }
      } else {
        double v[3];
        MHD_7UNROLL
        for (int i = 0; i < 3; i++) {
          const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
          v[i] = jp[i] - P.sigma * (u[i1] * P.B[i2] - u[i2] * P.B[i1]) - P.g[i];
        }
        MHD_7UNROLL
        for (int k = 0; k < 3; k++) F[RF_V + k * 27 + q] = w * id * (J[0 * 3 + k] * v[0] + J[1 * 3 + k] * v[1] + J[2 * 3 + k] * v[2]);
      }
    } else {  // S and w div j
      const double dj = id * (S.rs[RS_DJ + q] + S.rs[RS_DJ + 27 + q] + S.rs[RS_DJ + 54 + q]);
      F[RF_S + q] = w * id * ((ZJ ? P.zeta_j * dj : 0.0) - S.sigma_cell * S.rs[RS_F + q]);
      F[RF_DJ + q] = w * dj;
    }
  }
}

// interval of phase_D: first contraction of the u rows | partial sums of the j rows | phi rows
template <class RAdd>
MHD_7HD void res_stage_a(Cell7& S, const SmallDyn& C, int tid, int nt, RAdd& radd) {
  double* F = S.r3;
  for (int it = tid; it < 168; it += nt) {
    if (it < 108) {
      const int c = it / 36, f = (it / 9) % 4, q01 = it % 9;
      const double* x = F + (f < 3 ? RF_G + (c * 3 + f) * 27 : RF_H + c * 27) + q01;
      const double x0 = x[0], x1 = x[9], x2 = x[18];
      const double* tb = f == 2 ? C.LD[2][0] : C.LV[2][0];
      double* o = F + RR_A + ((c * 4 + f) * 3) * 9 + q01;
      MHD_7UNROLL
      for (int a = 0; a < 3; a++) o[a * 9] = tb[a * 3] * x0 + tb[a * 3 + 1] * x1 + tb[a * 3 + 2] * x2;
    } else if (it < 144) {
      const int r = it - 108, k = r / 12, i1 = (r / 6) % 2, i2 = (r / 3) % 2, qk = r % 3;
      const int s0 = pow3(k), s1 = pow3((k + 1) % 3), s2 = pow3((k + 2) % 3);
      double pv = 0.0, ps = 0.0;
      MHD_7UNROLL
      for (int q2 = 0; q2 < 3; q2++)
        MHD_7UNROLL
        for (int q1 = 0; q1 < 3; q1++) {
          const int q = qk * s0 + q1 * s1 + q2 * s2;
          const double t = C.RV[k][1][i1][q1] * C.RV[k][2][i2][q2];
          pv += t * F[RF_V + k * 27 + q];
          ps += t * F[RF_S + q];
        }
      F[RR_JP + ((k * 4 + i1 + 2 * i2) * 3 + qk) * 2] = pv;
      F[RR_JP + ((k * 4 + i1 + 2 * i2) * 3 + qk) * 2 + 1] = ps;
    } else {
      // phi rows: 8 rows x 3 planes q2 (one 27-term sum per row kept a single warp busy while 7 waited at the barrier);
      // the three partial sums of a row are accumulated by radd
      const int r = it - 144, lt = r / 3, q2 = r - 3 * lt, l0 = lt & 1, l1 = (lt >> 1) & 1, l2 = lt >> 2;
      double s = 0.0;
      MHD_7UNROLL
      for (int q1 = 0; q1 < 3; q1++)
        MHD_7UNROLL
        for (int q0 = 0; q0 < 3; q0++) s += F[RF_DJ + q0 + 3 * q1 + 9 * q2] * C.XV[0][l0][q0] * C.XV[1][l1][q1];
      radd(OFF_F + C.t_phi[lt], S.phi_sign * C.XV[2][l2][q2] * s);
    }
  }
}

// next interval: second contraction of the u rows | j rows
template <class RAdd>
MHD_7HD void res_stage_b(Cell7& S, const SmallDyn& C, int tid, int nt, RAdd& radd) {
  double* F = S.r3;
  for (int it = tid; it < 144; it += nt) {
    if (it < 108) {
      const int cf = it / 9, f = cf % 4, a2 = (it / 3) % 3, q0 = it % 3;
      const double* x = F + RR_A + (cf * 3 + a2) * 9 + q0;
      const double x0 = x[0], x1 = x[3], x2 = x[6];
      const double* tb = f == 1 ? C.LD[1][0] : C.LV[1][0];
      double* o = F + RR_B + (cf * 9 + a2) * 3 + q0;  // [(cf * 3 + a1) * 3 + a2] * 3 + q0
      MHD_7UNROLL
      for (int a = 0; a < 3; a++) o[a * 9] = tb[a * 3] * x0 + tb[a * 3 + 1] * x1 + tb[a * 3 + 2] * x2;
    } else {
      const int r = it - 108, k = r / 12, i0 = r % 3, i12 = (r / 3) % 4;
      const double* jp = F + RR_JP + ((k * 4 + i12) * 3) * 2;
      double s = 0.0;
      MHD_7UNROLL
      for (int qk = 0; qk < 3; qk++) s += jp[qk * 2] * C.RV[k][0][i0][qk] + jp[qk * 2 + 1] * C.RD[k][i0][qk];
      const int tj = 12 * k + i0 + 3 * i12;
      radd(OFF_J + S.slot_j[tj], S.sgn[tj] * s);
    }
  }
}

// interval of chunk_uu(0): last contraction of the u rows
template <class RAdd>
MHD_7HD void res_stage_c(Cell7& S, const SmallDyn& C, int tid, int nt, RAdd& radd) {
  const double* F = S.r3;
  for (int it = tid; it < 81; it += nt) {
    const int c = it / 27, t = it % 27, a0 = t % 3, a12 = ((t / 3) % 3) * 3 + t / 9;  // a1 * 3 + a2: layout of res_stage_b
    double s = 0.0;
    MHD_7UNROLL
    for (int f = 0; f < 4; f++) {
      const double* x = F + RR_B + (((c * 4 + f) * 9) + a12) * 3;
      const double* tb = f == 0 ? C.LD[0][a0] : C.LV[0][a0];
      s += tb[0] * x[0] + tb[1] * x[1] + tb[2] * x[2];
    }
    radd(c * 27 + S.slot_u[t], s);
  }
}

// ------------------------------------------------------------------ sweeps: store(e, local row (permuted numbering), value)
template <class Store>
MHD_7HD void sweep_section(const double* buf, int e0, int n, int ncol, int row0, int tid, int nt, Store& store) {
  for (int i = tid; i < n; i += nt) store(e0 + i, row0 + i / ncol, buf[i]);
}
template <class Store>
MHD_7HD void sweep_uu(const double* buf, int c, int tid, int nt, Store& store) {
  sweep_section(buf, E_UU + c * CH_UU_PAD, CH_UU, 81, c * 27, tid, nt, store);
}
template <class Store>
MHD_7HD void sweep_uj(const double* buf, int tid, int nt, Store& store) { sweep_section(buf, E_UJ, CH_UJ, 36, 0, tid, nt, store); }
template <class Store>
MHD_7HD void sweep_ju(const double* buf, int tid, int nt, Store& store) { sweep_section(buf, E_JU, CH_UJ, 81, OFF_J, tid, nt, store); }
template <class Store>
MHD_7HD void sweep_rest(const double* buf, int tid, int nt, Store& store) {
  sweep_section(buf + R_JJ, E_REST + R_JJ, NJ * NJ, 36, OFF_J, tid, nt, store);
  sweep_section(buf + R_JF, E_REST + R_JF, NJ * NF, 8, OFF_J, tid, nt, store);
  sweep_section(buf + R_FJ, E_REST + R_FJ, NF * NJ, 36, OFF_F, tid, nt, store);
  sweep_section(buf + R_UP, E_REST + R_UP, NU * NP, 4, 0, tid, nt, store);
  sweep_section(buf + R_PU, E_REST + R_PU, NP * NU, 81, OFF_P, tid, nt, store);
}

}  // namespace h7
}  // namespace mhd
