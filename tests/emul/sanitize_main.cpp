// AddressSanitizer / UBSan driver for the device code compiled on the CPU (tests/test_h1h1_host.py::test_device_code_is_clean_under_address_sanitizer):
// one cell with arbitrary tables through every template variant of the H1-H1 phases, thread counts 256 / 96 / 33, both
// thread orders, and the patch inversion for sizes 1..256 -- an out-of-range shared-memory index would abort here.
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <math.h>
extern "C" long long emul_h1h1_cells(long long ncells, const double* coords, const int* cell_nodes, const int* gids, const double* dir,
                          const double* x, const double* w, const double* geo_grad, const double* u_val, const double* u_grad,
                          const double* p_val, const double* phi_grad, const double* prm, int conv, int nt, int reverse,
                          double* K_out, double* R_out);
extern "C" int emul_patch_invert(double* A, int n, int nt, int reverse);
int main() {
  // one unit cube cell, random tables (arithmetic content is irrelevant: this run is for the address sanitizer)
  std::vector<double> coords = {0,0,0, 1,0,0, 0,1,0, 1,1,0, 0,0,1, 1,0,1, 0,1,1, 1,1,1};
  std::vector<int> cn = {0,1,2,3,4,5,6,7};
  std::vector<int> gids(149); for (int i = 0; i < 149; i++) gids[i] = (i % 7 == 0) ? -(i % 5) - 1 : i;
  std::vector<double> dir(6, 0.5), x(149, 0.25), w(27, 1.0 / 27), gg(27 * 24), nu(27 * 27), dnu(27 * 81), pp(27 * 4), dphi(27 * 192);
  srand(1);
  auto rnd = [](std::vector<double>& v) { for (auto& a : v) a = rand() / (double)RAND_MAX - 0.5; };
  rnd(nu); rnd(dnu); rnd(pp); rnd(dphi);
  // trilinear geometry gradients of the unit cube at the cell centre for every point (constant Jacobian = identity)
  for (int q = 0; q < 27; q++) for (int v = 0; v < 8; v++) for (int k = 0; k < 3; k++) {
    int b[3] = {v & 1, (v >> 1) & 1, (v >> 2) & 1};
    double g = (b[k] ? 1.0 : -1.0);
    for (int d = 0; d < 3; d++) if (d != k) g *= 0.5;
    gg[(q * 8 + v) * 3 + k] = g;
  }
  std::vector<double> prm = {1.0, 0.5, 3.0, 2.0, 0.1, 1.0, -0.3, 0.0, 0.0, 1.0};
  std::vector<double> K(149 * 149), R(149);
  long long bad = 0;
  for (int conv = 0; conv < 3; conv++) for (int zu = 0; zu < 2; zu++) for (int nt : {256, 96, 33}) for (int rev = 0; rev < 2; rev++) {
    prm[3] = zu ? 2.0 : 0.0;
    std::fill(K.begin(), K.end(), 0.0); std::fill(R.begin(), R.end(), 0.0);
    bad += emul_h1h1_cells(1, coords.data(), cn.data(), gids.data(), dir.data(), x.data(), w.data(), gg.data(), nu.data(), dnu.data(),
                           pp.data(), dphi.data(), prm.data(), conv, nt, rev, K.data(), R.data());
    for (double v : K) if (!std::isfinite(v)) bad++;
  }
  for (int n : {1, 7, 16, 17, 100, 225, 256}) {
    std::vector<double> A(n * n); rnd(A); for (int i = 0; i < n; i++) A[i * n + i] += 3.0;
    bad += emul_patch_invert(A.data(), n, 256, 0);
  }
  printf("bad=%lld\n", bad);
  return bad != 0;
}
