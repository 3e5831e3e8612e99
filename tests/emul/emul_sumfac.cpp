// CPU execution of gridapmhd.jl_b200/csrc/sumfac_uu.h (test infrastructure): the three contraction phases with the CTA
// replaced by a loop over thread ids (forward / reverse order over NaN-filled work arrays).
#include <string.h>

#include "../../gridapmhd.jl_b200/csrc/sumfac_uu.h"

using namespace mhd::sf;

namespace {
struct HostStore {
  double* K;  // [81][81], row c*27+a, col d*27+b
  void operator()(int a, int b, int c, int d, double v) { K[(c * 27 + a) * 81 + d * 27 + b] += v; }
};
}  // namespace

extern "C" int emul_sumfac_uu(const double* P, const signed char* ijk, const double* C, int nt, int reverse, double* K) {
  Tables T;
  memcpy(T.P, P, sizeof(T.P));
  memcpy(T.ijk, ijk, sizeof(T.ijk));
  Work* W = new Work;
  memset(W, 0xFF, sizeof(Work));
  memcpy(W->C, C, sizeof(W->C));
  HostStore st{K};
#define FOR_T for (int t_ = 0, t = reverse ? nt - 1 : 0; t_ < nt; t_++, t += reverse ? -1 : 1)
  FOR_T phase_stage1(*W, T, t, nt);
  FOR_T phase_stage2(*W, T, t, nt);
  FOR_T phase_stage3(*W, T, t, nt, st);
  delete W;
  return 0;
}
