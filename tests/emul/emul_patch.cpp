// CPU execution of the patch-inversion device code (test infrastructure): gridapmhd.jl_b200/csrc/patch_cell.h compiled
// with g++, the CTA replaced by a loop over thread ids per phase (forward or reverse order, to expose intra-phase races).
#include <string.h>

#include "../../gridapmhd.jl_b200/csrc/patch_cell.h"

static bool g_rev = false;
#define HOST_PHASE(...)                                                                   \
  for (int t_ = 0, tid = g_rev ? nt - 1 : 0; t_ < nt; t_++, tid += g_rev ? -1 : 1) { \
    __VA_ARGS__;                                                                          \
  }

extern "C" int emul_patch_invert(double* A, int n, int nt, int reverse) {
  g_rev = reverse != 0;
  mhd::patch::Shared* S = new mhd::patch::Shared;
  memset(S, 0xFF, sizeof(*S));
  MHD_PATCH_INVERT(HOST_PHASE, *S, A, n);
  const int sing = S->singular;
  delete S;
  return sing;
}
