// CPU execution of the H1-H1 device code (test infrastructure): compiles gridapmhd.jl_b200/csrc/h1h1_cell.h -- the very
// phase functions the CUDA kernels of h1h1.cu run -- with g++ and replaces the CTA by a loop over the thread ids, phase by
// phase (a barrier separates phases on the device).  Output: dense cell matrices / vectors, compared with the oracle in
// tests/test_h1h1_host.py, and a check that every stored entry e addresses the (row, col) the symbolic phase assumes.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../gridapmhd.jl_b200/csrc/h1h1_cell.h"

using namespace mhd::h1;

namespace {
// thread order inside a phase: forward or reverse -- a phase that read what another thread of the SAME phase wrote (a
// race on the device) gives NaNs in one of the two orders because the cell data starts NaN-filled
bool g_reverse = false;
#define FOR_T for (int t_ = 0, t = g_reverse ? nt - 1 : 0; t_ < nt; t_++, t += g_reverse ? -1 : 1)
struct HostStore {
  double* K;         // [149*149]
  long long* nbad;   // entries whose (row, col) disagree with entry_rowcol
  unsigned char* hit;  // [NENT] how often entry e was stored
  void operator()(int e, int li, int lj, double v) {
    int ri, rj;
    entry_rowcol(e, &ri, &rj);
    if (ri != li || rj != lj || e < 0 || e >= NENT) (*nbad)++;
    else hit[e]++;
    K[li * NLOC + lj] += v;
  }
};
struct HostAdd {
  double* R;
  void operator()(int li, double v) { R[li] += v; }
};

// fused = the phase sequence of h1h1_jacobian_kernel<CONV, ZU, RES = true>: Jacobian phases, then the residual phases on the
// same cell data (no reload)
template <int CONV, bool ZU>
void jac_cell(Shared& S, int nt, const Params& P, HostStore& st, const double* tab, bool fused = false, HostAdd* add = nullptr) {
  FOR_T phase_geometry(S, t, nt, tab);
  FOR_T phase_gradients(S, t, nt, tab);
  if (CONV != 0 || fused)
    FOR_T phase_point_values(S, t, nt);
  FOR_T phase_jac_coefficients<CONV, ZU>(S, t, nt, P);
  if (ZU)
    FOR_T phase_jac_projection(S, t, nt);
  FOR_T phase_jac_entries<CONV, ZU>(S, t, nt, P, st);
  if (fused) {
    FOR_T phase_res_points<ZU>(S, t, nt);
    FOR_T phase_res_coefficients<(CONV != 0 ? 1 : 0), ZU>(S, t, nt, P);
    FOR_T phase_res_rows(S, t, nt, *add);
  }
}

template <int CONV, bool ZU>
void res_cell(Shared& S, int nt, const Params& P, HostAdd& add, const double* tab) {
  FOR_T phase_geometry(S, t, nt, tab);
  FOR_T phase_gradients(S, t, nt, tab);
  FOR_T phase_point_values(S, t, nt);
  FOR_T phase_res_points<ZU>(S, t, nt);
  FOR_T phase_res_coefficients<CONV, ZU>(S, t, nt, P);
  FOR_T phase_res_rows(S, t, nt, add);
}
}  // namespace

extern "C" {

// prm = {alpha, beta, gamma, zeta_u, B[3], f[3]}; gids as the device holds them (>=0 free id into x, <0: -(index into dir)-1)
// returns the number of mis-addressed or not-exactly-once-stored entries (0 = the enumeration is consistent)
long long emul_h1h1_cells(long long ncells, const double* coords, const int* cell_nodes, const int* gids, const double* dir,
                          const double* x, const double* w, const double* geo_grad, const double* u_val, const double* u_grad,
                          const double* p_val, const double* phi_grad, const double* prm, int conv, int nt, int reverse,
                          double* K_out, double* R_out) {
  const bool fused = (reverse & 2) != 0;  // bit 1 of `reverse`: fused residual + Jacobian sequence (needs K_out and R_out)
  g_reverse = (reverse & 1) != 0;
  std::vector<double> T(T_TOTAL);
  pack_tables(w, geo_grad, u_val, u_grad, p_val, phi_grad, T.data());
  Params P;
  P.alpha = prm[0]; P.beta = prm[1]; P.gamma = prm[2]; P.zeta_u = prm[3];
  for (int i = 0; i < 3; i++) { P.B[i] = prm[4 + i]; P.f[i] = prm[7 + i]; }
  const bool zu = P.zeta_u != 0.0;
  Shared* S = new Shared;
  std::vector<unsigned char> hit(NENT);
  long long nbad = 0;
  for (long long c = 0; c < ncells; c++) {
    memset(S, 0xFF, sizeof(Shared));  // NaN-fill: a phase reading data no earlier phase wrote shows up in the result
    FOR_T phase_load(*S, t, nt, coords, cell_nodes + c * 8, gids + c * NLOC, nullptr, dir, x, T.data());
    if (K_out) {
      std::fill(hit.begin(), hit.end(), 0);
      HostStore st{K_out + c * NLOC * NLOC, &nbad, hit.data()};
      HostAdd fadd{R_out ? R_out + c * NLOC : nullptr};
      const bool f = fused && R_out;
      if (conv == 0) { if (zu) jac_cell<0, true>(*S, nt, P, st, T.data(), f, &fadd); else jac_cell<0, false>(*S, nt, P, st, T.data(), f, &fadd); }
      else if (conv == 1) { if (zu) jac_cell<1, true>(*S, nt, P, st, T.data(), f, &fadd); else jac_cell<1, false>(*S, nt, P, st, T.data(), f, &fadd); }
      else { if (zu) jac_cell<2, true>(*S, nt, P, st, T.data(), f, &fadd); else jac_cell<2, false>(*S, nt, P, st, T.data(), f, &fadd); }
      for (int e = 0; e < NENT; e++) nbad += hit[e] != 1;
    }
    if (R_out && !(fused && K_out)) {
      if (K_out) {
        memset(S, 0xFF, sizeof(Shared));
        FOR_T phase_load(*S, t, nt, coords, cell_nodes + c * 8, gids + c * NLOC, nullptr, dir, x, T.data());
      }
      HostAdd add{R_out + c * NLOC};
      if (conv == 0) { if (zu) res_cell<0, true>(*S, nt, P, add, T.data()); else res_cell<0, false>(*S, nt, P, add, T.data()); }
      else { if (zu) res_cell<1, true>(*S, nt, P, add, T.data()); else res_cell<1, false>(*S, nt, P, add, T.data()); }
    }
  }
  delete S;
  return nbad;
}

int emul_h1h1_sizes(int* nent, int* nent_pad, int* shared_bytes) {
  *nent = NENT;
  *nent_pad = NENT_PAD;
  *shared_bytes = (int)sizeof(Shared);
  return 0;
}
}
