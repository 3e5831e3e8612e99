// CPU execution of the v6 H1-HDiv Jacobian device code (gridapmhd.jl_b200/csrc/hdiv_cell.h + sumfac_uu.h): the phase
// functions run with a loop over thread ids (forward / reverse) on NaN-filled cell data; output = dense 129 x 129 cell
// matrices and a check that every entry of the enumeration is stored exactly once at the right (row, col).
#include <string.h>

#include <vector>

#include "../../gridapmhd.jl_b200/csrc/hdiv_cell.h"

using namespace mhd::h6;

namespace {
bool g_rev = false;
#define FOR_T for (int t_ = 0, t = g_rev ? nt - 1 : 0; t_ < nt; t_++, t += g_rev ? -1 : 1)
struct HostStore {
  double* K;
  long long* nbad;
  unsigned char* hit;
  void operator()(int e, int li, int lj, double v) {
    int ri, rj;
    entry_rowcol(e, &ri, &rj);
    if (e < 0 || e >= NENT || ri != li || rj != lj) (*nbad)++;
    else hit[e]++;
    K[li * NLOC + lj] += v;
  }
};

template <int CONV, bool ZU, bool ZJ>
void cell(Shared& S, const mhd::sf::Tables& T, int nt, const Params& P, const double* tab, HostStore& st) {
  FOR_T phase_geometry(S, t, nt, tab);
  FOR_T phase_mapped_bases<CONV>(S, t, nt, tab);
  FOR_T phase_coefficients<CONV, ZU>(S, t, nt, P, tab);
  if (ZU) FOR_T phase_projection(S, t, nt);
  FOR_T mhd::sf::phase_stage1(S.W, T, t, nt);
  FOR_T mhd::sf::phase_stage2(S.W, T, t, nt);
  FOR_T phase_entries<CONV, ZU, ZJ>(S, T, t, nt, P, tab, st);
}
}  // namespace

extern "C" {
// tab: packed tables in the T_* layout (h6::T_*, = common.h).  prm = {alpha, beta, gamma, sigma, zeta_u, zeta_j, B[3]}
// returns the number of mis-addressed / not-exactly-once entries; *tensor_dev = deviation of the tables from tensor products
long long emul_hdiv_cells(long long ncells, const double* coords, const int* cell_nodes, const int* gids, const signed char* jsign,
                          const unsigned char* cell_solid, const double* cell_sigma, const double* dir, const double* x,
                          const double* tab, const signed char* ijk, const double* prm, int conv, int nt, int reverse, double* K_out,
                          double* tensor_dev) {
  g_rev = reverse != 0;
  mhd::sf::Tables T;
  *tensor_dev = derive_tensor_tables(tab + T_NU, tab + T_DNU, (const int8_t*)ijk, &T);
  Params P;
  P.alpha = prm[0]; P.beta = prm[1]; P.gamma = prm[2]; P.sigma = prm[3]; P.zeta_u = prm[4]; P.zeta_j = prm[5];
  for (int i = 0; i < 3; i++) P.B[i] = prm[6 + i];
  const bool zu = P.zeta_u != 0.0, zj = P.zeta_j != 0.0;
  Shared* S = new Shared;
  std::vector<unsigned char> hit(NENT);
  long long nbad = 0;
  for (long long c = 0; c < ncells; c++) {
    memset(S, 0xFF, sizeof(Shared));
    const bool solid = cell_solid && cell_solid[c];
    FOR_T phase_load(*S, t, nt, coords, cell_nodes + c * 8, gids + c * NLOC, nullptr, (const int8_t*)jsign + c * NJ, dir, x, tab,
                     conv != 0, solid, solid ? cell_sigma[c] : 0.0, P.sigma);
    std::fill(hit.begin(), hit.end(), 0);
    HostStore st{K_out + c * NLOC * NLOC, &nbad, hit.data()};
#define RUN(C, U, J) cell<C, U, J>(*S, T, nt, P, tab, st)
#define RUNJ(C, U) do { if (zj) RUN(C, U, true); else RUN(C, U, false); } while (0)
#define RUNU(C) do { if (zu) RUNJ(C, true); else RUNJ(C, false); } while (0)
    if (conv == 0) RUNU(0); else if (conv == 1) RUNU(1); else RUNU(2);
    for (int e = 0; e < NENT; e++) nbad += hit[e] != 1;
  }
  delete S;
  return nbad;
}
}
