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
extern "C" long long emul_hdiv_cells(long long ncells, const double* coords, const int* cell_nodes, const int* gids, const signed char* jsign,
                          const unsigned char* cell_solid, const double* cell_sigma, const double* dir, const double* x,
                          const double* tab, const signed char* ijk, const double* prm, int conv, int nt, int reverse, double* K_out,
                          double* tensor_dev);
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
  {
    // v6 H1-HDiv cell code (hdiv_cell.h): packed tables in the T_* layout, arbitrary content, both cell kinds
    const int TT = 27 + 648 + 729 + 2187 + 108 + 2916 + 972 + 216;
    std::vector<double> tab(TT);
    rnd(tab);
    for (int i = 0; i < 27; i++) tab[i] = 1.0 / 27;
    for (int i = 0; i < 648; i++) tab[27 + i] = gg[i];
    std::vector<int> g129(129);
    for (int i = 0; i < 129; i++) g129[i] = (i % 7 == 0) ? -(i % 5) - 1 : i;
    std::vector<signed char> js(36), ijk(81);
    for (int i = 0; i < 36; i++) js[i] = (i % 3) ? 1 : -1;
    for (int a = 0; a < 27; a++) { ijk[a * 3] = a % 3; ijk[a * 3 + 1] = (a / 3) % 3; ijk[a * 3 + 2] = a / 9; }
    std::vector<double> x129(129, 0.25), K6(129 * 129), p6 = {1.0, 0.5, 3.0, 0.7, 0.0, 0.0, 0.1, 1.0, -0.3};
    unsigned char solid[1] = {0};
    double sig[1] = {2.0}, dev = 0.0;
    for (int conv = 0; conv < 3; conv++) for (int z = 0; z < 2; z++) for (int nt : {256, 96, 33}) for (int sol = 0; sol < 2; sol++) {
      p6[4] = z ? 2.0 : 0.0; p6[5] = z ? 3.0 : 0.0; solid[0] = (unsigned char)sol;
      std::fill(K6.begin(), K6.end(), 0.0);
      bad += emul_hdiv_cells(1, coords.data(), cn.data(), g129.data(), js.data(), solid, sig, dir.data(), x129.data(), tab.data(),
                             ijk.data(), p6.data(), conv, nt, nt == 96, K6.data(), &dev);
      for (double v : K6) if (!std::isfinite(v)) bad++;
    }
  }
  printf("bad=%lld\n", bad);
  return bad != 0;
}
