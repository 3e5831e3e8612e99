// CPU execution of the v7 H1-HDiv Jacobian device code (gridapmhd.jl_b200/csrc/hdiv7_cell.h + hdiv7_tables.h): the phase
// functions run with a loop over thread ids (forward / reverse) on NaN-filled cell data, in exactly the phase sequence of
// the kernel (hdiv7.cu), two staging buffers included.  Output: dense 129 x 129 cell matrices in the REFERENCE local
// numbering, and a count of map entries not stored exactly once at the (row, col) the enumeration promises.
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../gridapmhd.jl_b200/csrc/hdiv7_cell.h"

using namespace mhd::h7;

namespace {
bool g_rev = false;
#define FOR_T for (int t_ = 0, t = g_rev ? nt - 1 : 0; t_ < nt; t_++, t += g_rev ? -1 : 1)

struct HostStore {
  double* K;             // [129][129] reference numbering
  const int* unperm;     // permuted local index -> reference local index
  long long* nbad;
  unsigned char* hit;
  void operator()(int e, int row, double v) {
    int li, lj;
    if (e < 0 || e >= NENT || !entry_rowcol(e, &li, &lj) || li != row) { (*nbad)++; return; }
    hit[e]++;
    K[unperm[li] * NLOC + unperm[lj]] += v;
  }
};

struct HostRAdd {
  double* R;          // [129] reference numbering, nullable
  const int* unperm;
  void operator()(int row, double v) { R[unperm[row]] += v; }
};

// the kernel's interval structure (hdiv_v7.cu): every FOR_T is one barrier interval
template <int CONV, bool ZU, bool ZJ, bool RES>
void cell(Cell7& S, const Small7& C, const mhd::h7::Tab7& T, int nt, const Params& P, HostStore& st, HostRAdd& ra) {
  constexpr int WU = (CONV != 0 || RES) ? 1 : 0;
  FOR_T { phase_geom_a<WU>(S, C, t, nt, &T.gg[0][0]); if (RES) res_pv(S, C, t, nt); }
  FOR_T phase_geom_b<WU>(S, C, t, nt, T.w);
  if (WU || ZU) FOR_T { if (WU) phase_points(S, C, t, nt); if (ZU) phase_Mp(S, C, t, nt); }
  FOR_T { phase_fields<CONV, ZJ>(S, C, t, nt, P); if (RES) res_divu(S, t, nt); if (ZU) phase_Minv(S, t, nt); }
  FOR_T { phase_stage1<ZJ>(S, C, C, t, nt); if (RES) res_d(S, C, t, nt, ra); }
  FOR_T { phase_stage2<CONV, ZJ>(S, C, C, t, nt); if (RES) res_fields<CONV, ZU, ZJ>(S, C, t, nt, P); }
  FOR_T { phase_D<ZU>(S, C, C, t, nt); if (RES) res_stage_a(S, C, t, nt, ra); }
  if (ZU || RES) FOR_T { if (ZU) phase_E(S, t, nt, P.zeta_u); if (RES) res_stage_b(S, C, t, nt, ra); }
  double* b0 = S.r1;
  double* b1 = S.r3;
  FOR_T { chunk_uu<CONV, ZU>(S, C, C, t, nt, 0, b0); if (RES) res_stage_c(S, C, t, nt, ra); }
  FOR_T { sweep_uu(b0, 0, t, nt, st); chunk_uu<CONV, ZU>(S, C, C, t, nt, 1, b1); }
  FOR_T { sweep_uu(b1, 1, t, nt, st); chunk_uu<CONV, ZU>(S, C, C, t, nt, 2, b0); }
  FOR_T { sweep_uu(b0, 2, t, nt, st); chunk_uj(S, C, t, nt, P, b1, false); }
  FOR_T { sweep_uj(b1, t, nt, st); chunk_uj(S, C, t, nt, P, b0, true); }
  FOR_T { sweep_ju(b0, t, nt, st); chunk_rest<ZJ>(S, C, t, nt, b1); }
  FOR_T sweep_rest(b1, t, nt, st);
}
}  // namespace

extern "C" {
// raw tables in the layouts of mhd_tables_t.  prm = {alpha, beta, gamma, sigma, zeta_u, zeta_j, B[3], f[3], g[3]}.
// returns -1 if the tables lack the tensor structure, else the number of mis-stored entries
long long emul_hdiv7_cells(long long ncells, const double* coords, const int* cell_nodes, const int* gids, const signed char* jsign,
                           const unsigned char* cell_solid, const double* cell_sigma, const double* dir, const double* x,
                           const double* w, const double* geo_grad, const double* u_val, const double* u_grad, const double* p_val,
                           const double* j_val, const double* j_div, const double* phi_val, const double* prm, int conv, int nt,
                           int reverse, double* K_out, double* R_out /* nullable: [ncells][129] cell residuals */) {
  g_rev = reverse != 0;
  mhd::h7::Tab7* T = new mhd::h7::Tab7;
  if (!build_tab7(w, geo_grad, u_val, u_grad, p_val, j_val, j_div, phi_val, T)) { delete T; return -1; }
  Small7* C = new Small7;
  for (int t = 0; t < nt; t++) small_from_tab(*C, *T, t, nt);
  Params P;
  P.alpha = prm[0]; P.beta = prm[1]; P.gamma = prm[2]; P.sigma = prm[3]; P.zeta_u = prm[4]; P.zeta_j = prm[5];
  for (int i = 0; i < 3; i++) { P.B[i] = prm[6 + i]; P.f[i] = prm[9 + i]; P.g[i] = prm[12 + i]; }
  const bool zu = P.zeta_u != 0.0, zj = P.zeta_j != 0.0;
  Cell7* S = new Cell7;
  std::vector<unsigned char> hit(NENT);
  long long nbad = 0;
  for (long long c = 0; c < ncells; c++) {
    memset(S, 0xFF, sizeof(Cell7));
    // the per-cell permutation of symbolic.cu (cell_permutation): every field sorted by global id, Dirichlet / absent last
    const int* g = gids + c * NLOC;
    unsigned char perm[64];
    int pg[NLOC], unperm[NLOC];
    auto keyof = [&](int id) { return id < 0 ? 0x7fffffff : id; };
    int ord[36];
    for (int i = 0; i < 27; i++) ord[i] = i;
    std::stable_sort(ord, ord + 27, [&](int a, int b) { return keyof(g[a]) < keyof(g[b]); });
    for (int s = 0; s < 27; s++) perm[s] = (unsigned char)ord[s];
    for (int i = 0; i < 36; i++) ord[i] = i;
    std::stable_sort(ord, ord + 36, [&](int a, int b) { return keyof(g[OFF_J + a]) < keyof(g[OFF_J + b]); });
    for (int s = 0; s < 36; s++) perm[27 + s] = (unsigned char)(ord[s] | (jsign[c * NJ + ord[s]] < 0 ? 0x80 : 0));
    perm[63] = 0;
    for (int i = 0; i < NLOC; i++) {
      int src = i;
      if (i < NU) src = (i / 27) * 27 + perm[i % 27];
      else if (i >= OFF_J && i < OFF_F) src = OFF_J + (perm[27 + i - OFF_J] & 0x7F);
      pg[i] = g[src];
      unperm[i] = src;
    }
    const bool solid = cell_solid && cell_solid[c];
    FOR_T phase_load(*S, *C, t, nt, coords, cell_nodes + c * 8, pg, nullptr, perm, dir, x, conv != 0 || R_out != nullptr, solid,
                     solid ? cell_sigma[c] : 0.0, P.sigma, nullptr, R_out != nullptr);
    std::fill(hit.begin(), hit.end(), 0);
    HostStore st{K_out + c * NLOC * NLOC, unperm, &nbad, hit.data()};
    HostRAdd ra{R_out ? R_out + c * NLOC : nullptr, unperm};
#define RUN(CV, U, J) do { if (R_out) cell<CV, U, J, true>(*S, *C, *T, nt, P, st, ra); else cell<CV, U, J, false>(*S, *C, *T, nt, P, st, ra); } while (0)
#define RUNJ(CV, U) do { if (zj) RUN(CV, U, true); else RUN(CV, U, false); } while (0)
#define RUNU(CV) do { if (zu) RUNJ(CV, true); else RUNJ(CV, false); } while (0)
    if (conv == 0) RUNU(0); else if (conv == 1) RUNU(1); else RUNU(2);
    for (int e = 0; e < NENT; e++) {
      int li, lj;
      nbad += hit[e] != (entry_rowcol(e, &li, &lj) ? 1 : 0);
    }
  }
  delete S; delete C; delete T;
  return nbad;
}
}
