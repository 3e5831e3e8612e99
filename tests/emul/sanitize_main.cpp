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
extern "C" long long emul_hdiv7_cells(long long ncells, const double* coords, const int* cell_nodes, const int* gids, const signed char* jsign,
                           const unsigned char* cell_solid, const double* cell_sigma, const double* dir, const double* x,
                           const double* w, const double* geo_grad, const double* u_val, const double* u_grad, const double* p_val,
                           const double* j_val, const double* j_div, const double* phi_val, const double* prm, int conv, int nt,
                           int reverse, double* K_out, double* R_out);
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
    // v7 H1-HDiv cell code (hdiv7_cell.h, Jacobian + fused residual): tables that ARE tensor products (the structure discovery
    // refuses anything else) -- quadratic / linear 1-D factors with arbitrary coefficients -- both cell kinds
    double tq[3] = {0.1127016653792583, 0.5, 0.8872983346207417};
    auto l2 = [&](int i, double t) { return i == 0 ? (2 * t - 1) * (t - 1) : (i == 1 ? 4 * t * (1 - t) : t * (2 * t - 1)); };
    auto d2 = [&](int i, double t) { return i == 0 ? 4 * t - 3 : (i == 1 ? 4 - 8 * t : 4 * t - 1); };
    auto l1 = [&](int i, double t) { return i == 0 ? 1 - t : t; };
    std::vector<double> uv(27 * 27), ug(27 * 81), jv(27 * 108, 0.0), jd(27 * 36), fv(27 * 8), p4(27 * 4), w27(27, 1.0 / 27);
    for (int q = 0; q < 27; q++) {
      const double t[3] = {tq[q % 3], tq[(q / 3) % 3], tq[q / 9]};
      for (int a = 0; a < 27; a++) {
        const int i[3] = {a % 3, (a / 3) % 3, a / 9};
        uv[q * 27 + a] = l2(i[0], t[0]) * l2(i[1], t[1]) * l2(i[2], t[2]);
        for (int k = 0; k < 3; k++) {
          double g = 1.0;
          for (int d = 0; d < 3; d++) g *= d == k ? d2(i[d], t[d]) : l2(i[d], t[d]);
          ug[(q * 27 + a) * 3 + k] = g;
        }
      }
      for (int m = 0; m < 36; m++) {  // component k = m / 12: quadratic in direction k, linear in the other two
        const int k = m / 12, r = m % 12, i0 = r % 3, i1 = (r / 3) % 2, i2 = r / 6;
        const double a1 = l1(i1, t[(k + 1) % 3]), a2 = l1(i2, t[(k + 2) % 3]);
        jv[(q * 36 + m) * 3 + k] = l2(i0, t[k]) * a1 * a2;
        jd[q * 36 + m] = d2(i0, t[k]) * a1 * a2;
      }
      for (int l = 0; l < 8; l++) fv[q * 8 + l] = l1(l & 1, t[0]) * l1((l >> 1) & 1, t[1]) * l1(l >> 2, t[2]);
      p4[q * 4] = 1.0; p4[q * 4 + 1] = t[0]; p4[q * 4 + 2] = t[1]; p4[q * 4 + 3] = t[2];
    }
    std::vector<int> g129(129);
    for (int i = 0; i < 129; i++) g129[i] = (i % 7 == 0) ? -(i % 5) - 1 : i;
    std::vector<signed char> js(36);
    for (int i = 0; i < 36; i++) js[i] = (i % 3) ? 1 : -1;
    std::vector<double> x129(129, 0.25), K7(129 * 129), R7(129), p7 = {1.0, 0.5, 3.0, 0.7, 0.0, 0.0, 0.1, 1.0, -0.3, 0.1, 0.2, 0.3, 0.3, 0.2, 0.1};
    unsigned char solid[1] = {0};
    double sig[1] = {2.0};
    for (int conv = 0; conv < 3; conv++) for (int z = 0; z < 2; z++) for (int nt : {256, 96, 33}) for (int sol = 0; sol < 2; sol++)
      for (int res = 0; res < 2; res++) {
        p7[4] = z ? 2.0 : 0.0; p7[5] = z ? 3.0 : 0.0; solid[0] = (unsigned char)sol;
        std::fill(K7.begin(), K7.end(), 0.0); std::fill(R7.begin(), R7.end(), 0.0);
        const long long rc = emul_hdiv7_cells(1, coords.data(), cn.data(), g129.data(), js.data(), solid, sig, dir.data(), x129.data(), w27.data(),
                                              gg.data(), uv.data(), ug.data(), p4.data(), jv.data(), jd.data(), fv.data(), p7.data(), conv, nt,
                                              nt == 96, K7.data(), res ? R7.data() : nullptr);
        bad += rc != 0;  // -1: structure not recognised; > 0: entries stored at the wrong place
        for (double v : K7) if (!std::isfinite(v)) bad++;
        for (double v : R7) if (!std::isfinite(v)) bad++;
      }
  }
  printf("bad=%lld\n", bad);
  return bad != 0;
}
