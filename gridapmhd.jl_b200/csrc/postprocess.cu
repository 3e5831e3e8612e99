// Post-processing of GridapMHD.hunt on the device (src/Applications/hunt.jl:239-260): discrete norms uh_l2, uh_h1,
// jh_l2 and the errors eu_l2, eu_h1, ej_l2 against the analytical Hunt solution (Fourier series analytical_hunt_u /
// analytical_hunt_j, hunt.jl:372-457) with the degree 2*(order+1) quadrature.  In the reference's published runs this
// section (`time_post_process`) rivals the Jacobian assembly because the series (nsums terms, several exp/cos each)
// is evaluated at every quadrature point on the host: 44 s at nc=32 ... 7277 s at nc=400.
//
// One CTA per cell, 4 threads per quadrature point: FE fields by the first of them, the series terms dealt round-robin
// to all four, per-cell partial sums to global memory, a second kernel adds them in a fixed order (deterministic).
#include "common.h"

namespace mhd {

constexpr int PP_MAXQ = 64;
constexpr int PP_PARTS = 4;

struct SeriesTerm {
  double al, r1, r2, c2, c3, d1, d2, cu;  // alpha_k, r1_k, r2_k, r2/N, r1/N, 1+exp(-2 r1), 1+exp(-2 r2), 2(-1)^k/(l alpha^3)
};

__global__ void hunt_series_terms(int n, double Ha, double l, SeriesTerm* t) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  const double al = (k + 0.5) * 3.14159265358979323846 / l;
  const double N = sqrt(Ha * Ha + 4.0 * al * al);
  const double r1 = 0.5 * (Ha + N), r2 = 0.5 * (-Ha + N);
  SeriesTerm s;
  s.al = al; s.r1 = r1; s.r2 = r2; s.c2 = r2 / N; s.c3 = r1 / N;
  s.d1 = 1.0 + exp(-2.0 * r1); s.d2 = 1.0 + exp(-2.0 * r2);
  s.cu = ((k & 1) ? -2.0 : 2.0) / (l * al * al * al);
  t[k] = s;
}

struct HuntPost {
  double a, mu, sigma, grad_pz, u0, jscale;
  int nsums;
};

__global__ void __launch_bounds__(PP_MAXQ* PP_PARTS)
hunt_norms_kernel(int64_t ncells, int nq, const double* __restrict__ w, const double* __restrict__ geo_val,
                  const double* __restrict__ geo_grad, const double* __restrict__ u_val, const double* __restrict__ u_grad,
                  const double* __restrict__ j_val, const double* __restrict__ coords, const int32_t* __restrict__ cell_nodes,
                  const int32_t* __restrict__ gids, const int8_t* __restrict__ jsign, const double* __restrict__ dirv,
                  const double* __restrict__ x, const SeriesTerm* __restrict__ terms, HuntPost P,
                  double* __restrict__ partial /* [ncells][6] */) {
  __shared__ double U[NLOC], X[24], S[PP_PARTS][PP_MAXQ][5], R[PP_MAXQ * PP_PARTS / 32][6];
  const int64_t cell = blockIdx.x;
  const int tid = threadIdx.x, q = tid % PP_MAXQ, part = tid / PP_MAXQ;
  for (int i = tid; i < NLOC; i += blockDim.x) {
    const int32_t g = gids[cell * NLOC + i];
    U[i] = g >= 0 ? x[g] : dirv[-g - 1];
  }
  if (tid < 24) X[tid] = coords[(int64_t)cell_nodes[cell * 8 + tid / 3] * 3 + tid % 3];
  __syncthreads();
  const bool active = q < nq;
  // physical point (trilinear map)
  double xq[3] = {0.0, 0.0, 0.0};
  if (active)
    for (int v = 0; v < 8; v++) {
      const double gv = geo_val[q * 8 + v];
      for (int i = 0; i < 3; i++) xq[i] = fma(gv, X[v * 3 + i], xq[i]);
    }
  // ---- this thread's share of the series (terms part, part + 4, ...)
  double s_u = 0.0, s_ux = 0.0, s_uy = 0.0, s_hdx = 0.0, s_hdy = 0.0;
  const double xi = xq[0] / P.a, eta = xq[1] / P.a;
  const bool inside = active && xi <= 1.0 && xi >= -1.0 && eta <= 1.0 && eta >= -1.0;
  if (inside) {
    for (int k = part; k <= P.nsums; k += PP_PARTS) {
      const SeriesTerm t = terms[k];
      const double e1m = exp(-t.r1 * (1.0 - eta)), e1p = exp(-t.r1 * (1.0 + eta));
      const double e2m = exp(-t.r2 * (1.0 - eta)), e2p = exp(-t.r2 * (1.0 + eta));
      double sk, ck;
      sincos(t.al * xi, &sk, &ck);
      const double V2 = t.c2 * (e1m + e1p) / t.d1, V3 = t.c3 * (e2m + e2p) / t.d2;
      const double V2e = t.c2 * t.r1 * (e1m - e1p) / t.d1, V3e = t.c3 * t.r2 * (e2m - e2p) / t.d2;
      s_u += t.cu * ck * (1.0 - V2 - V3);
      s_ux += t.cu * (-t.al * sk) * (1.0 - V2 - V3);
      s_uy += t.cu * ck * (-V2e - V3e);
      const double H2 = t.c2 * (e1m - e1p) / t.d1, H3 = t.c3 * (e2m - e2p) / t.d2;
      const double H2y = t.c2 * (t.r1 / P.a) * (e1m + e1p) / t.d1, H3y = t.c3 * (t.r2 / P.a) * (e2m + e2p) / t.d2;
      // -2(-1)^k sin / (a l alpha^2) = -cu alpha sin / a
      s_hdx += -t.cu * t.al / P.a * sk * (H2 - H3);
      s_hdy += t.cu * ck * (H2y - H3y);
    }
  }
  S[part][q][0] = s_u; S[part][q][1] = s_ux; S[part][q][2] = s_uy; S[part][q][3] = s_hdx; S[part][q][4] = s_hdy;
  __syncthreads();
  double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (part == 0 && active) {
    double ser[5];
    for (int i = 0; i < 5; i++) ser[i] = (S[0][q][i] + S[1][q][i]) + (S[2][q][i] + S[3][q][i]);
    const double su = (P.a * P.a / P.mu) * (-P.grad_pz);
    const double uz = su * ser[0], uz_x = su / P.a * ser[1], uz_y = su / P.a * ser[2];
    const double sj = P.a * P.a * sqrt(P.sigma) / sqrt(P.mu) * (-P.grad_pz);
    const double jx = sj * ser[4], jy = sj * (-ser[3]);
    // geometry
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int v = 0; v < 8; v++)
      for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) J[i][k] = __dadd_rn(__dmul_rn(X[v * 3 + i], geo_grad[(q * 8 + v) * 3 + k]), J[i][k]);  // see hdiv7_cell.h
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                 c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02, id = 1.0 / det;
    double inv[3][3];  // inv[k][i] = d xi_k / d x_i
    inv[0][0] = c00 * id; inv[1][0] = c01 * id; inv[2][0] = c02 * id;
    inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    const double wq = w[q] * fabs(det);
    // uh, grad uh (g[d][c] = d_d u_c), jh
    double uh[3] = {0, 0, 0}, g[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, jr[3] = {0, 0, 0};
    for (int a = 0; a < 27; a++) {
      const double n = u_val[q * 27 + a];
      double dn[3];
      for (int d = 0; d < 3; d++)
        dn[d] = u_grad[(q * 27 + a) * 3 + 0] * inv[0][d] + u_grad[(q * 27 + a) * 3 + 1] * inv[1][d] + u_grad[(q * 27 + a) * 3 + 2] * inv[2][d];
      for (int c = 0; c < 3; c++) {
        const double uc = U[c * 27 + a] * P.u0;
        uh[c] = fma(n, uc, uh[c]);
        for (int d = 0; d < 3; d++) g[d][c] = fma(dn[d], uc, g[d][c]);
      }
    }
    for (int m = 0; m < NJ; m++) {
      const double jm = U[OFF_J + m] * (double)jsign[cell * NJ + m] * P.jscale;
      for (int k = 0; k < 3; k++) jr[k] = fma(j_val[(q * NJ + m) * 3 + k], jm, jr[k]);
    }
    double jh[3];
    for (int i = 0; i < 3; i++) jh[i] = (J[i][0] * jr[0] + J[i][1] * jr[1] + J[i][2] * jr[2]) * id;
    double uu = 0.0, gg = 0.0, jj = 0.0;
    for (int c = 0; c < 3; c++) {
      uu = fma(uh[c], uh[c], uu);
      jj = fma(jh[c], jh[c], jj);
      for (int d = 0; d < 3; d++) gg = fma(g[d][c], g[d][c], gg);
    }
    const double eu[3] = {-uh[0], -uh[1], uz - uh[2]};
    const double ej[3] = {jx - jh[0], jy - jh[1], -jh[2]};
    double ge = 0.0;
    for (int d = 0; d < 3; d++)
      for (int c = 0; c < 3; c++) {
        double v = -g[d][c];
        if (c == 2 && d == 0) v += uz_x;
        if (c == 2 && d == 1) v += uz_y;
        ge = fma(v, v, ge);
      }
    const double el2 = eu[0] * eu[0] + eu[1] * eu[1] + eu[2] * eu[2];
    acc[0] = wq * el2;
    acc[1] = wq * (ge + el2);
    acc[2] = wq * (ej[0] * ej[0] + ej[1] * ej[1] + ej[2] * ej[2]);
    acc[3] = wq * uu;
    acc[4] = wq * (gg + uu);
    acc[5] = wq * jj;
  }
  // CTA reduction in a fixed order
  for (int i = 0; i < 6; i++)
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], o);
  if ((tid & 31) == 0)
    for (int i = 0; i < 6; i++) R[tid >> 5][i] = acc[i];
  __syncthreads();
  if (tid < 6) {
    double s = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x / 32); wv++) s += R[wv][tid];
    partial[cell * 6 + tid] = s;
  }
}

__global__ void __launch_bounds__(256) reduce_partials(int64_t ncells, const double* __restrict__ partial, double* __restrict__ out) {
  __shared__ double sh[256];
  const int comp = blockIdx.x;
  double s = 0.0;
  for (int64_t c = threadIdx.x; c < ncells; c += 256) s += partial[c * 6 + comp];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[comp] = sh[0];
}

int hunt_error_norms(mhd_operator* op, const double* d_x, const mhd_tables_t* t, const mhd_hunt_post_t* p, double* out6) {
  MHD_CHECK(t->nq >= 1 && t->nq <= PP_MAXQ, MHD_E_INVALID, "mhd_hunt_error_norms: nq=%d, at most %d quadrature points", t->nq, PP_MAXQ);
  MHD_CHECK(t->w && t->geo_grad && t->u_val && t->u_grad && t->j_val && t->phi_val, MHD_E_INVALID, "mhd_hunt_error_norms: null table");
  MHD_CHECK(p->nsums >= 0 && p->a > 0.0 && p->mu > 0.0 && p->sigma > 0.0, MHD_E_INVALID, "mhd_hunt_error_norms: invalid parameters");
  const int nq = t->nq;
  double *d_tab = nullptr, *d_part = nullptr, *d_out = nullptr;
  SeriesTerm* d_terms = nullptr;
  const int64_t o_w = 0, o_gv = o_w + nq, o_gg = o_gv + nq * 8, o_uv = o_gg + nq * 24, o_ug = o_uv + nq * 27,
                o_jv = o_ug + nq * 81, total = o_jv + nq * 108;
  std::vector<double> h(total);
  memcpy(&h[o_w], t->w, nq * sizeof(double));
  memcpy(&h[o_gv], t->phi_val, nq * 8 * sizeof(double));  // Q1 nodal basis = the vertex functions of the geometry map
  memcpy(&h[o_gg], t->geo_grad, nq * 24 * sizeof(double));
  memcpy(&h[o_uv], t->u_val, nq * 27 * sizeof(double));
  memcpy(&h[o_ug], t->u_grad, nq * 81 * sizeof(double));
  memcpy(&h[o_jv], t->j_val, nq * 108 * sizeof(double));
  int rc = 0;
  auto cleanup = [&]() { cudaFree(d_tab); cudaFree(d_part); cudaFree(d_out); cudaFree(d_terms); };
#define PR(x) do { if (!rc) rc = (x); } while (0)
  PR(dev_alloc(&d_tab, total));
  PR(dev_alloc(&d_part, op->ncells * 6));
  PR(dev_alloc(&d_out, 6));
  PR(dev_alloc(&d_terms, p->nsums + 1));
  PR(h2d(d_tab, h.data(), total));
  if (!rc) {
    hunt_series_terms<<<(p->nsums + 1 + 127) / 128, 128, 0, g_stream>>>(p->nsums, p->Ha, 1.0 /* l = b/a, square duct */, d_terms);
    HuntPost P{p->a, p->mu, p->sigma, p->grad_pz, p->u0, p->jscale, p->nsums};
    hunt_norms_kernel<<<(unsigned)op->ncells, PP_MAXQ * PP_PARTS, 0, g_stream>>>(
        op->ncells, nq, d_tab + o_w, d_tab + o_gv, d_tab + o_gg, d_tab + o_uv, d_tab + o_ug, d_tab + o_jv, op->d_coords,
        op->d_cell_nodes, op->d_gids, op->d_jsign, op->d_dir, d_x, d_terms, P, d_part);
    reduce_partials<<<6, 256, 0, g_stream>>>(op->ncells, d_part, d_out);
    g_launches += 3;
    if (cudaPeekAtLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "hunt_norms_kernel", __FILE__, __LINE__);
  }
  double hout[6];
  PR(d2h(hout, d_out, 6));
  if (!rc && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "sync", __FILE__, __LINE__);
#undef PR
  cleanup();
  if (rc) return rc;
  for (int i = 0; i < 6; i++) out6[i] = sqrt(hout[i]);
  return 0;
}

}  // namespace mhd
