// In-place inversion of one dense patch matrix (n <= 256, row-major, leading dimension n) by a cooperative thread array:
// blocked Gauss-Jordan with partial (row) pivoting.  Used by the vertex-patch block-Jacobi smoother of the (u,j) block
// (patch.cu; reference: PatchBasedSmoothers.BlockJacobiSolver with collected factorisations, src/Solvers/gmg.jl:62-81).
//
// Why blocked: an unblocked sweep touches the whole matrix once per pivot (n^3 x 16 B of L2 traffic, 180 MB for n = 225);
// here PB pivots are first eliminated inside an n x PB panel held in shared memory, then applied to the rest of the
// matrix as one rank-PB update (traffic / PB).  Partial pivoting is needed: the (u,j) patch matrices are positive real
// but carry a skew part (Lorentz coupling, gamma = Ha^2) that is orders of magnitude above the symmetric part.
//
// Like h1h1_cell.h the algorithm is written as barrier-separated PHASES, functions of (thread id, thread count), so that
// tests/emul/emul_patch.cpp runs the very same code on the CPU against numpy.linalg.inv.  Requires thread count >= n.
//
// Algebra of one block step on the pivot columns K = [k0, k0+b): with the row swaps applied to whole rows (multipliers
// included, as in LAPACK getrf) the panel ends as  P[K,:] = inv(A[K,K]),  P[O,:] = -A[O,K] inv(A[K,K])  (O = other rows), and
//   new A[i, j] = (i in K ? 0 : A[i, j]) + sum_s P[i, s] A[k0+s, j]     for the columns j outside K,
//   new A[:, K] = P.
// Row swaps permute the implicit identity columns of the augmented system; the column swaps of the final phase undo them
// in reverse order (Gauss-Jordan in place, cf. Numerical Recipes `gaussj`).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MHD_PHD __host__ __device__ __forceinline__
#define MHD_PUNROLL _Pragma("unroll")
#else
#define MHD_PHD inline
#define MHD_PUNROLL
#endif

namespace mhd {
namespace patch {

constexpr int NMAX = 256;  // largest patch (threads per CTA = columns)
constexpr int PB = 16;     // panel width
constexpr int NLEAD = 16;  // leaders of the two-level pivot search (each scans NMAX / NLEAD rows)

struct Shared {
  double P[PB][NMAX + 1];  // panel, TRANSPOSED: P[s][i] = panel entry (row i, column s); thread i walks its row with
                           // consecutive lanes on consecutive words (a row-major panel gave 32-way bank conflicts:
                           // 112 ms -> see DESIGN 4.6), the odd leading dimension keeps the panel load conflict-free
  double prow[PB];      // scaled pivot row of the current sub-step
  double lead_val[NLEAD];
  int lead_idx[NLEAD];
  int piv[NMAX];        // pivot row chosen for column k
  int singular;         // a zero pivot was met (the patch keeps going with pivot 1 to stay finite; reported)
};

MHD_PHD double dabs(double x) { return x < 0.0 ? -x : x; }

// phase: load the panel columns [k0, k0+b)
MHD_PHD void phase_load_panel(Shared& S, int tid, int nt, const double* A, int n, int k0, int b) {
  for (int idx = tid; idx < n * b; idx += nt) S.P[idx % b][idx / b] = A[(int64_t)(idx / b) * n + k0 + idx % b];
}
// sub-step s, phase a: leaders scan their rows for the largest |P[i][s]|, i >= k0+s
MHD_PHD void phase_pivot_leaders(Shared& S, int tid, int nt, int n, int k0, int s) {
  if (tid < NLEAD) {
    const int chunk = NMAX / NLEAD;
    double best = -1.0;
    int bi = -1;
    for (int i = tid * chunk; i < (tid + 1) * chunk && i < n; i++)
      if (i >= k0 + s) {
        const double v = dabs(S.P[s][i]);
        if (v > best) { best = v; bi = i; }
      }
    S.lead_val[tid] = best;
    S.lead_idx[tid] = bi;
  }
}
// phase b: pick the pivot row
MHD_PHD void phase_pivot_pick(Shared& S, int tid, int nt, int k0, int s) {
  if (tid == 0) {
    double best = -1.0;
    int bi = k0 + s;
    for (int l = 0; l < NLEAD; l++)
      if (S.lead_val[l] > best) { best = S.lead_val[l]; bi = S.lead_idx[l]; }
    S.piv[k0 + s] = bi;
  }
}
// phase c: swap the two panel rows (whole rows: multipliers of earlier sub-steps included)
MHD_PHD void phase_swap_panel_rows(Shared& S, int tid, int nt, int k0, int s, int b) {
  const int r = S.piv[k0 + s];
  if (tid < b && r != k0 + s) {
    const double t = S.P[tid][k0 + s];
    S.P[tid][k0 + s] = S.P[tid][r];
    S.P[tid][r] = t;
  }
}
// phase d: scaled pivot row into prow
MHD_PHD void phase_scale_pivot_row(Shared& S, int tid, int nt, int k0, int s, int b) {
  if (tid < b) {
    double p = S.P[s][k0 + s];
    if (p == 0.0) {
      p = 1.0;
      if (tid == 0) S.singular = 1;
    }
    S.prow[tid] = (tid == s ? 1.0 : S.P[tid][k0 + s]) / p;
  }
}
// phase e: eliminate inside the panel (thread i owns row i)
MHD_PHD void phase_eliminate_panel(Shared& S, int tid, int nt, int n, int k0, int s, int b) {
  for (int i = tid; i < n; i += nt) {
    if (i == k0 + s) {
      for (int t = 0; t < b; t++) S.P[t][i] = S.prow[t];
    } else {
      const double f = S.P[s][i];
      for (int t = 0; t < b; t++) S.P[t][i] = (t == s ? 0.0 : S.P[t][i]) - f * S.prow[t];
    }
  }
}
// phase: rank-b update of the columns outside the panel, panel columns written back (thread j owns column j)
MHD_PHD void phase_update(const Shared& S, int tid, int nt, double* A, int n, int k0, int b) {
  for (int j = tid; j < n; j += nt) {
    if (j >= k0 && j < k0 + b) {
      for (int i = 0; i < n; i++) A[(int64_t)i * n + j] = S.P[j - k0][i];
      continue;
    }
    double rb[PB];
    MHD_PUNROLL
    for (int s = 0; s < PB; s++) {
      rb[s] = 0.0;
      if (s < b) {
        const int r = S.piv[k0 + s];
        if (r != k0 + s) {
          const double t = A[(int64_t)(k0 + s) * n + j];
          A[(int64_t)(k0 + s) * n + j] = A[(int64_t)r * n + j];
          A[(int64_t)r * n + j] = t;
        }
      }
    }
    MHD_PUNROLL
    for (int s = 0; s < PB; s++)
      if (s < b) rb[s] = A[(int64_t)(k0 + s) * n + j];
    // rows in chunks of RU: the loads of a chunk are issued together (the stores of the previous rows alias them as far as
    // the compiler knows, so a row-at-a-time loop would pay the full memory latency per row)
    constexpr int RU = 8;
    for (int i0 = 0; i0 < n; i0 += RU) {
      double v[RU];
      MHD_PUNROLL
      for (int u = 0; u < RU; u++) {
        const int i = i0 + u;
        v[u] = (i < n && !(i >= k0 && i < k0 + b)) ? A[(int64_t)i * n + j] : 0.0;
      }
      MHD_PUNROLL
      for (int u = 0; u < RU; u++) {
        const int i = i0 + u < n ? i0 + u : n - 1;
        MHD_PUNROLL
        for (int s = 0; s < PB; s++) v[u] += S.P[s][i] * rb[s];  // P[i][s] = 0 for s >= b (phase_clear_panel_tail)
      }
      MHD_PUNROLL
      for (int u = 0; u < RU; u++)
        if (i0 + u < n) A[(int64_t)(i0 + u) * n + j] = v[u];
    }
  }
}
// phase: zero the unused panel columns of a short last block
MHD_PHD void phase_clear_panel_tail(Shared& S, int tid, int nt, int n, int b) {
  if (b < PB)
    for (int idx = tid; idx < n * (PB - b); idx += nt) S.P[b + idx % (PB - b)][idx / (PB - b)] = 0.0;
}
// final phase: undo the row permutation on the columns, in reverse order (thread i owns row i)
MHD_PHD void phase_unscramble(const Shared& S, int tid, int nt, double* A, int n) {
  for (int i = tid; i < n; i += nt)
    for (int k = n - 1; k >= 0; k--) {
      const int r = S.piv[k];
      if (r != k) {
        const double t = A[(int64_t)i * n + k];
        A[(int64_t)i * n + k] = A[(int64_t)i * n + r];
        A[(int64_t)i * n + r] = t;
      }
    }
}

// The whole inversion as a sequence of phases.  PHASE(...) runs its statements for every thread id `tid` of `nt` and ends
// with a barrier: `__VA_ARGS__; __syncthreads();` on the device (patch.cu), a loop over tid on the host (emulation).
#define MHD_PATCH_INVERT(PHASE, S, A, n)                                                            \
  do {                                                                                              \
    PHASE(if (tid == 0) (S).singular = 0);                                                          \
    for (int k0_ = 0; k0_ < (n); k0_ += mhd::patch::PB) {                                           \
      const int b_ = (n) - k0_ < mhd::patch::PB ? (n) - k0_ : mhd::patch::PB;                       \
      PHASE(mhd::patch::phase_load_panel((S), tid, nt, (A), (n), k0_, b_);                          \
            mhd::patch::phase_clear_panel_tail((S), tid, nt, (n), b_));                             \
      for (int s_ = 0; s_ < b_; s_++) {                                                             \
        PHASE(mhd::patch::phase_pivot_leaders((S), tid, nt, (n), k0_, s_));                         \
        PHASE(mhd::patch::phase_pivot_pick((S), tid, nt, k0_, s_));                                 \
        PHASE(mhd::patch::phase_swap_panel_rows((S), tid, nt, k0_, s_, b_));                        \
        PHASE(mhd::patch::phase_scale_pivot_row((S), tid, nt, k0_, s_, b_));                        \
        PHASE(mhd::patch::phase_eliminate_panel((S), tid, nt, (n), k0_, s_, b_));                   \
      }                                                                                             \
      PHASE(mhd::patch::phase_update((S), tid, nt, (A), (n), k0_, b_));                             \
    }                                                                                               \
    PHASE(mhd::patch::phase_unscramble((S), tid, nt, (A), (n)));                                    \
  } while (0)

}  // namespace patch
}  // namespace mhd
