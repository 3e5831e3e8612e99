// Device-resident flexible GMRES + block-triangular preconditioner.
//
// Stands behind the Gridap LinearSolver seam (symbolic_setup / numerical_setup! / solve!) that GridapMHD fills
// at src/main.jl:181-190 and src/Solvers/badia2024.jl:2-48:
//   FGMRESSolver(m,P;maxiter=m,rtol,atol)           -> Fgmres below: the whole restart cycle is enqueued on the
//                                                      stream; Hessenberg/Givens/convergence live in device memory
//                                                      (one tiny kernel per Arnoldi step), no host round trip.
//   BlockTriangularSolver([uj p phi], :upper)       -> apply_block_tri: exact cell-block inverses of the
//                                                      discontinuous p / phi mass matrices (badia2024.jl:11-15,23-24)
//                                                      and an inner Jacobi-preconditioned GMRES on the (u,j) block
//                                                      (the reference uses LU/MUMPS there, badia2024.jl:22).
// Orthogonalisation is classical Gram-Schmidt with one re-orthogonalisation pass (2 fused multi-dots per step
// instead of the j+1 dependent dots of the reference's modified Gram-Schmidt: SURVEY.md 5.8).
#include <cusolverDn.h>
#include <dlfcn.h>

#include <functional>
#include <map>
#include <tuple>

#include "common.h"
#include "h1h1_cell.h"

namespace mhd {

constexpr int MAXM = 64;

struct KrylovScalars {  // lives in device memory
  double H[(MAXM + 1) * MAXM];
  double cs[MAXM], sn[MAXM], g[MAXM + 1], y[MAXM];
  double h[MAXM + 2];   // fresh Gram-Schmidt coefficients (pass 1)
  double h2[MAXM + 2];  // pass 2
  double hn;            // ||w||^2 after orthogonalisation
  double beta;          // current residual norm
  double beta2;         // scratch: squared norm
  double tol;
  double inv;           // scaling for the next basis vector
  double hist[4 * MAXM + 2];
  int reorth;           // this Arnoldi step needs the second Gram-Schmidt pass (DGKS criterion), see k_dgks_scalars
  int done;             // converged (or breakdown): later steps of the cycle become no-ops
  int k;                // Arnoldi steps taken in this cycle
  int iters;            // total iterations
};

__global__ void k_cycle_begin(KrylovScalars* S, double rtol, double atol, int first, int total_hist) {
  // S->beta2 holds ||r||^2
  const double beta = sqrt(S->beta2);
  S->beta = beta;
  if (first) {
    S->tol = fmax(atol, rtol * beta);
    S->iters = 0;
    S->hist[0] = beta;
  }
  S->k = 0;
  S->done = (beta <= S->tol) ? 1 : 0;
  S->g[0] = beta;
  for (int i = 1; i <= MAXM; i++) S->g[i] = 0.0;
  for (int i = 0; i < MAXM; i++) S->y[i] = 0.0;
  S->inv = beta > 0.0 ? 1.0 / beta : 0.0;
}

// V = done ? 0 : inv * w
__global__ void __launch_bounds__(256) k_scale_to(int64_t n, const KrylovScalars* __restrict__ S, const double* __restrict__ w,
                                                   double* __restrict__ v) {
  const double a = S->done ? 0.0 : S->inv;
  const bool dead = S->done != 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) v[i] = dead ? 0.0 : a * w[i];
}

__global__ void k_add_h(KrylovScalars* S, int j) {
  const int i = threadIdx.x;
  if (i <= j) S->h[i] += S->h2[i];
}

// one Arnoldi step's scalar work: new Hessenberg column, Givens rotations, residual estimate.  Called by ALL threads of a
// SCALAR_T-thread block (uniform control flow): the column and the rotations are staged in shared memory in parallel and thread 0
// walks the dependent chain there -- as a single thread on global memory the chain of L2 round trips cost 5.6 us per step.
constexpr int SCALAR_T = 96;  // >= MAXM + 2
__device__ void arnoldi_scalars(KrylovScalars* S, int j, int m, int maxiter) {
  __shared__ double col[MAXM + 2], cs[MAXM], sn[MAXM];
  if (S->done) return;
  const int tid = threadIdx.x;
  if (tid <= j) col[tid] = S->h[tid];
  if (tid < j) {
    cs[tid] = S->cs[tid];
    sn[tid] = S->sn[tid];
  }
  __syncthreads();
  if (tid == 0) {
    const double hn = sqrt(fmax(S->hn, 0.0));
    col[j + 1] = hn;
    for (int i = 0; i < j; i++) {
      const double t = cs[i] * col[i] + sn[i] * col[i + 1];
      col[i + 1] = -sn[i] * col[i] + cs[i] * col[i + 1];
      col[i] = t;
    }
    const double a = col[j], b = col[j + 1];
    const double d = hypot(a, b);
    if (d == 0.0) {  // exact breakdown: stop
      S->done = 1;
    } else {
      const double c = a / d, sj = b / d;
      S->cs[j] = c;
      S->sn[j] = sj;
      col[j] = d;
      col[j + 1] = 0.0;
      const double gj = S->g[j];
      S->g[j + 1] = -sj * gj;
      S->g[j] = c * gj;
      const double beta = fabs(sj * gj);
      S->beta = beta;
      S->k = j + 1;
      const int it = S->iters + 1;
      S->iters = it;
      if (it < 4 * MAXM + 2) S->hist[it] = beta;
      S->inv = hn > 0.0 ? 1.0 / hn : 0.0;
      if (beta <= S->tol || hn == 0.0 || it >= maxiter) S->done = 1;
    }
  }
  __syncthreads();
  if (tid <= j + 1) S->H[tid * MAXM + j] = col[tid];
}
__global__ void __launch_bounds__(SCALAR_T) k_arnoldi_scalars(KrylovScalars* S, int j, int m, int maxiter) { arnoldi_scalars(S, j, m, maxiter); }

// Pass 1 of the fused Gram-Schmidt step is in: h = V'w and h[j+1] = ||w||^2.  ||w1||^2 = ||w||^2 - ||h||^2 (Pythagoras, V
// orthonormal).  Daniel-Gragg-Kaufman-Stewart criterion: the second pass is needed only when the projection removed more than half
// of w (||w1||^2 < ||w||^2 / 2) -- otherwise w1 is orthogonal to V to working precision, the Pythagoras norm has lost at most one
// bit, and the step is finished here: Hessenberg column, rotations, scaling of v_{j+1}.  (always = 1: classical CGS2, the default)
__global__ void __launch_bounds__(SCALAR_T) k_dgks_scalars(KrylovScalars* S, int j, int m, int maxiter, int always) {
  __shared__ int go;
  if (threadIdx.x == 0) {
    go = 0;
    int reorth = 0;
    if (!S->done) {
      if (always) {
        reorth = 1;
      } else {
        double s1 = 0.0;
        for (int i = 0; i <= j; i++) s1 += S->h[i] * S->h[i];
        const double n0 = S->h[j + 1], n1 = n0 - s1;
        if (!(n1 >= 0.5 * n0)) {
          reorth = 1;
        } else {
          S->hn = n1;
          go = 1;
        }
      }
    }
    S->reorth = reorth;
  }
  __syncthreads();
  if (go) arnoldi_scalars(S, j, m, maxiter);
}
// second pass (only when reorth): h += h2, ||w2||^2 by Pythagoras from ||w1||^2 (computed, not estimated, in pass 2), then the step
__global__ void __launch_bounds__(SCALAR_T) k_cgs2_finish(KrylovScalars* S, int j, int m, int maxiter) {
  __shared__ double sq[MAXM + 2];
  if (!S->reorth || S->done) return;
  const int tid = threadIdx.x;
  if (tid <= j) {
    const double h2 = S->h2[tid];
    sq[tid] = h2 * h2;
    S->h[tid] += h2;
  }
  __syncthreads();
  if (tid == 0) {
    double s2 = 0.0;
    for (int i = 0; i <= j; i++) s2 += sq[i];
    S->hn = S->h2[j + 1] - s2;
  }
  __syncthreads();  // also makes this block's writes to S->h / S->hn visible to its own later reads
  arnoldi_scalars(S, j, m, maxiter);
}

__global__ void k_back_substitute(KrylovScalars* S) {
  const int k = S->k;
  for (int i = 0; i < MAXM; i++) S->y[i] = 0.0;
  for (int i = k - 1; i >= 0; i--) {
    double s = S->g[i];
    for (int l = i + 1; l < k; l++) s -= S->H[i * MAXM + l] * S->y[l];
    S->y[i] = s / S->H[i * MAXM + i];
  }
}

__global__ void __launch_bounds__(256) k_sub(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) out[i] = a[i] - b[i];
}
__global__ void __launch_bounds__(256) k_mul(int64_t n, const double* __restrict__ d, const double* __restrict__ v, double* __restrict__ z) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) z[i] = d[i] * v[i];
}

static unsigned vgrid(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)device_sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

using VecOp = std::function<int(const double*, double*)>;

// FGMRES on the rows [0,n) of an operator given as a matvec closure. Vectors handed to matvec have `ld` entries
// (owned rows + ghost section) so that the halo exchange can fill the ghosts.
struct Fgmres {
  mhd_operator* op = nullptr;
  int64_t n = 0, ld = 0;
  int m = 15;
  double *V = nullptr, *Z = nullptr, *w = nullptr, *t = nullptr;
  KrylovScalars* S = nullptr;
  bool flexible = true;

  int init(mhd_operator* op_, int64_t n_, int64_t ld_, int m_, bool flexible_) {
    op = op_; n = n_; ld = ld_; m = m_; flexible = flexible_;
    MHD_CHECK(m >= 1 && m <= MAXM, MHD_E_INVALID, "restart length %d outside [1,%d]", m, MAXM);
    MHD_TRY(dev_alloc(&V, (int64_t)(m + 1) * ld));
    MHD_TRY(dev_alloc(&Z, (int64_t)(flexible ? m : 1) * ld));
    MHD_TRY(dev_alloc(&w, ld));
    MHD_TRY(dev_alloc(&t, ld));
    MHD_TRY(dev_alloc(&S, 1));
    MHD_CUDA(cudaMemsetAsync(V, 0, (size_t)(m + 1) * ld * 8, g_stream));
    MHD_CUDA(cudaMemsetAsync(Z, 0, (size_t)(flexible ? m : 1) * ld * 8, g_stream));
    MHD_CUDA(cudaMemsetAsync(w, 0, (size_t)ld * 8, g_stream));
    MHD_CUDA(cudaMemsetAsync(t, 0, (size_t)ld * 8, g_stream));
    MHD_CUDA(cudaMemsetAsync(S, 0, sizeof(KrylovScalars), g_stream));
    return 0;
  }
  void release() {
    drop_graphs();
    cudaFree(V); cudaFree(Z); cudaFree(w); cudaFree(t); cudaFree(S);
    V = Z = w = t = nullptr; S = nullptr;
  }

  // squared norm of v[0..n) into *d_out (device), all-reduced
  int norm2(const double* v, double* d_out) {
    MHD_TRY(launch_multi_dot(op, n, 1, v, ld, v, d_out));
    return allreduce_sum(d_out, 1);
  }

  // ---- CUDA graph of a restart cycle: the cycle is a fixed sequence of launches on fixed buffers, so it is
  // captured once per (b, x, first, zero_guess) and replayed -- ~10 launches per Arnoldi step stop costing host time and
  // inter-kernel gaps.  Not used while another capture is active (inner solvers are captured as part of the outer cycle), with
  // more than one rank (the fused SpMV + halo kernel takes its round counter as a launch argument) or when MHD_KRYLOV_GRAPH=0.
  struct GraphKey {
    const double* b; double* x; bool first, zero;
    bool operator<(const GraphKey& o) const { return std::tie(b, x, first, zero) < std::tie(o.b, o.x, o.first, o.zero); }
  };
  std::map<GraphKey, cudaGraphExec_t> graphs;
  bool graphs_ok = true;
  void drop_graphs() {
    for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
    graphs.clear();
  }

  int cycle(const VecOp& matvec, const VecOp& precond, const double* b, double* x, double rtol, double atol, int maxiter,
            bool first, bool zero_guess) {
    static int want = -1;
    if (want < 0) {
      const char* e = getenv("MHD_KRYLOV_GRAPH");
      want = e ? atoi(e) : 1;
    }
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(g_stream, &st);
    if (!want || !graphs_ok || g_nranks > 1 || g_prof_on || st != cudaStreamCaptureStatusNone)
      return enqueue_cycle(matvec, precond, b, x, rtol, atol, maxiter, first, zero_guess);
    const GraphKey key{b, x, first, zero_guess};
    auto it = graphs.find(key);
    if (it == graphs.end()) {
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        graphs_ok = false;
        return enqueue_cycle(matvec, precond, b, x, rtol, atol, maxiter, first, zero_guess);
      }
      const int rc = enqueue_cycle(matvec, precond, b, x, rtol, atol, maxiter, first, zero_guess);
      const cudaError_t ec = cudaStreamEndCapture(g_stream, &g);
      cudaGraphExec_t ge = nullptr;
      if (rc != 0 || ec != cudaSuccess || g == nullptr || cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) {
        cudaGetLastError();  // something in the cycle cannot be captured: run it the plain way from now on
        if (g) cudaGraphDestroy(g);
        graphs_ok = false;
        if (rc != 0) return rc;
        return enqueue_cycle(matvec, precond, b, x, rtol, atol, maxiter, first, zero_guess);
      }
      cudaGraphDestroy(g);
      it = graphs.emplace(key, ge).first;
    }
    MHD_CUDA(cudaGraphLaunch(it->second, g_stream));
    g_launches += launches_per_cycle;
    return 0;
  }
  int64_t launches_per_cycle = 0;

  // one restart cycle, fully enqueued. x (ld entries) is updated in place.
  int enqueue_cycle(const VecOp& matvec, const VecOp& precond, const double* b, double* x, double rtol, double atol, int maxiter,
                    bool first, bool zero_guess) {
    const int64_t l0 = g_launches;
    // r = b - A x
    if (zero_guess && first) {
      MHD_CUDA(cudaMemcpyAsync(w, b, (size_t)n * 8, cudaMemcpyDeviceToDevice, g_stream));
    } else {
      MHD_TRY(matvec(x, t));
      k_sub<<<vgrid(n), 256, 0, g_stream>>>(n, b, t, w);
      MHD_LAUNCH_CHECK();
    }
    MHD_TRY(norm2(w, &S->beta2));
    k_cycle_begin<<<1, 1, 0, g_stream>>>(S, rtol, atol, first ? 1 : 0, 0);
    MHD_LAUNCH_CHECK();
    k_scale_to<<<vgrid(n), 256, 0, g_stream>>>(n, S, w, V);
    MHD_LAUNCH_CHECK();
    const bool fused = gs_fused_ok(m);
    for (int j = 0; j < m; j++) {
      double* Vj = V + (int64_t)j * ld;
      double* Zj = flexible ? Z + (int64_t)j * ld : Z;
      MHD_TRY(precond(Vj, Zj));
      MHD_TRY(matvec(Zj, w));
      if (fused) {
        // CGS2 in 4 launches + the scalar step: h = V'w | w1 = w - V h | (h2, ||w1||^2) = (V'w1, w1'w1) | v_{j+1} = (w1 - V h2)/||.||
        // Pass 1: (h, ||w||^2) = (V'w, w'w) in one sweep; the scalar kernel decides whether a second pass is needed (always, unless
        // MHD_KRYLOV_DGKS=1).  If not, the first update already writes v_{j+1} = (w - V h) / ||.|| and the three launches of the
        // second pass return at once (they read S->reorth); with more than one rank its all-reduce still runs, on unused data.
        static int always = -1;
        if (always < 0) {
          // default: always re-orthogonalise (classical CGS2).  MHD_KRYLOV_DGKS=1 makes the second pass conditional: measured on
          // cfg2 it almost never saves the pass (a preconditioned operator is close to the identity, so the projection on v_j
          // removes most of w and the criterion fires) -- 0.424 against 0.427 ms per iteration, 100 against 96 outer iterations.
          const char* e = getenv("MHD_KRYLOV_DGKS");
          always = (e && atoi(e) != 0) ? 0 : 1;
        }
        double* Vn = V + (int64_t)(j + 1) * ld;
        MHD_TRY(launch_gs_dots(op, n, j + 1, true, V, ld, w, S->h, nullptr));
        MHD_TRY(allreduce_sum(S->h, j + 2));
        k_dgks_scalars<<<1, SCALAR_T, 0, g_stream>>>(S, j, m, maxiter, always);
        MHD_LAUNCH_CHECK();
        MHD_TRY(launch_gs_update(n, j + 1, V, ld, S->h, w, &S->inv, &S->done, Vn, &S->reorth, 1, w));
        MHD_TRY(launch_gs_dots(op, n, j + 1, true, V, ld, w, S->h2, &S->reorth));
        MHD_TRY(allreduce_sum(S->h2, j + 2));
        k_cgs2_finish<<<1, SCALAR_T, 0, g_stream>>>(S, j, m, maxiter);
        MHD_LAUNCH_CHECK();
        MHD_TRY(launch_gs_update(n, j + 1, V, ld, S->h2, w, &S->inv, &S->done, Vn, &S->reorth, 2, nullptr));
        continue;
      }
      // CGS2: h = V^T w ; w -= V h ; h2 = V^T w ; w -= V h2 ; h += h2
      MHD_TRY(launch_multi_dot(op, n, j + 1, V, ld, w, S->h));
      MHD_TRY(allreduce_sum(S->h, j + 1));
      MHD_TRY(launch_multi_axpy(n, j + 1, V, ld, S->h, -1.0, w));
      MHD_TRY(launch_multi_dot(op, n, j + 1, V, ld, w, S->h2));
      MHD_TRY(allreduce_sum(S->h2, j + 1));
      MHD_TRY(launch_multi_axpy(n, j + 1, V, ld, S->h2, -1.0, w));
      k_add_h<<<1, 64, 0, g_stream>>>(S, j);
      MHD_LAUNCH_CHECK();
      MHD_TRY(norm2(w, &S->hn));
      k_arnoldi_scalars<<<1, SCALAR_T, 0, g_stream>>>(S, j, m, maxiter);
      MHD_LAUNCH_CHECK();
      k_scale_to<<<vgrid(n), 256, 0, g_stream>>>(n, S, w, V + (int64_t)(j + 1) * ld);
      MHD_LAUNCH_CHECK();
    }
    k_back_substitute<<<1, 1, 0, g_stream>>>(S);
    MHD_LAUNCH_CHECK();
    if (flexible) {
      MHD_TRY(launch_multi_axpy(n, m, Z, ld, S->y, 1.0, x));
    } else {
      // x += P^{-1} (V y)
      MHD_CUDA(cudaMemsetAsync(w, 0, (size_t)n * 8, g_stream));
      MHD_TRY(launch_multi_axpy(n, m, V, ld, S->y, 1.0, w));
      MHD_TRY(precond(w, t));
      MHD_TRY(launch_axpy(n, 1.0, t, x));
    }
    launches_per_cycle = g_launches - l0;
    return 0;
  }
};

}  // namespace mhd

using namespace mhd;

// ---- cuSOLVER (dense LU of the (u,j) block), bound lazily like NCCL so the library has no link-time dependency
namespace mhd {
struct CusolverApi {
  void* lib = nullptr;
  cusolverStatus_t (*Create)(cusolverDnHandle_t*) = nullptr;
  cusolverStatus_t (*Destroy)(cusolverDnHandle_t) = nullptr;
  cusolverStatus_t (*SetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
  cusolverStatus_t (*DgetrfBufferSize)(cusolverDnHandle_t, int, int, double*, int, int*) = nullptr;
  cusolverStatus_t (*Dgetrf)(cusolverDnHandle_t, int, int, double*, int, double*, int*, int*) = nullptr;
  cusolverStatus_t (*Dgetrs)(cusolverDnHandle_t, cublasOperation_t, int, int, const double*, int, const int*, double*, int, int*) = nullptr;
};
static CusolverApi g_cs;
static int load_cusolver() {
  if (g_cs.lib) return 0;
  const char* names[] = {"libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  MHD_CHECK(lib != nullptr, MHD_E_INVALID, "cannot dlopen libcusolver.so.11 (needed for MHD_UJ_DENSE_LU): %s", dlerror());
#define SYM(field, name)                                                              \
  do {                                                                                \
    *(void**)(&g_cs.field) = dlsym(lib, name);                                        \
    MHD_CHECK(g_cs.field != nullptr, MHD_E_INVALID, "libcusolver lacks symbol %s", name); \
  } while (0)
  SYM(Create, "cusolverDnCreate");
  SYM(Destroy, "cusolverDnDestroy");
  SYM(SetStream, "cusolverDnSetStream");
  SYM(DgetrfBufferSize, "cusolverDnDgetrf_bufferSize");
  SYM(Dgetrf, "cusolverDnDgetrf");
  SYM(Dgetrs, "cusolverDnDgetrs");
#undef SYM
  g_cs.lib = lib;
  return 0;
}
#define MHD_CUSOLVER(call)                                             \
  do {                                                                 \
    cusolverStatus_t _s = (call);                                      \
    if (_s != CUSOLVER_STATUS_SUCCESS) {                               \
      set_error("cuSOLVER error %d in %s", (int)_s, #call);            \
      return MHD_E_CUDA;                                               \
    }                                                                  \
  } while (0)

// dense column-major copy of the leading n x n block of the CSR matrix
__global__ void __launch_bounds__(256)
k_csr_block_to_dense(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
                     const double* __restrict__ nzval, double* __restrict__ D) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  for (int64_t p = rowptr[row] + lane; p < rowptr[row + 1]; p += 32) {
    const int32_t c = colval[p];
    if (c < n) D[(int64_t)c * n + row] = nzval[p];
  }
}
}  // namespace mhd

struct mhd_solver {
  mhd_operator* op = nullptr;
  mhd_solver_opts_t opts;
  Fgmres outer, inner;
  double* d_dinv = nullptr;     // [nrows] Jacobi
  double* d_minv_p = nullptr;   // [ncells*16] inverse P1disc mass blocks
  double* d_minv_f = nullptr;   // [ncells*64] inverse Q1disc mass blocks
  double *d_b = nullptr, *d_x = nullptr, *d_t1 = nullptr, *d_t2 = nullptr, *d_t3 = nullptr;
  int64_t n_uj = 0;
  bool setup_done = false;
  // dense LU of the (u,j) block
  cusolverDnHandle_t cs = nullptr;
  double* d_dense = nullptr;
  double* d_lu_work = nullptr;
  int* d_ipiv = nullptr;
  int* d_info = nullptr;
  int lu_lwork = 0;
  // vertex-patch smoother of the (u,j) block (patch.cu)
  mhd::PatchData* patches = nullptr;
  double *d_t4 = nullptr, *d_t5 = nullptr;
  // H1-H1 block preconditioner: u block = `inner` + `patches`, phi block = `inner2` + `patches_phi`, p = d_minv_p
  Fgmres inner2;
  mhd::PatchData* patches_phi = nullptr;
  int64_t off_p = 0, off_phi = 0;
};

namespace mhd {

__global__ void __launch_bounds__(256)
k_extract_dinv(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
               const double* __restrict__ nzval, double* __restrict__ dinv) {
  for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * 256) {
    double d = 0.0;
    int64_t lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while (lo <= hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int32_t c = colval[mid];
      if (c == r) { d = nzval[mid]; break; }
      if (c < r) lo = mid + 1; else hi = mid - 1;
    }
    dinv[r] = d != 0.0 ? 1.0 / d : 1.0;
  }
}

// inverse cell mass blocks of the discontinuous p (4x4) and phi (8x8) spaces: M[k][l] = sum_q w |det J| b_k b_l
constexpr int TM_W = T_W, TM_GG = T_GG, TM_PP = T_PP, TM_CHI = T_CHI;

template <int NB>
__device__ void invert_spd(double* a /* NB x 2NB */) {
  for (int p = 0; p < NB; p++) {
    const double ip = 1.0 / a[p * 2 * NB + p];
    for (int j = 0; j < 2 * NB; j++) a[p * 2 * NB + j] *= ip;
    for (int i = 0; i < NB; i++)
      if (i != p) {
        const double f = a[i * 2 * NB + p];
        if (f != 0.0)
          for (int j = 0; j < 2 * NB; j++) a[i * 2 * NB + j] -= f * a[p * 2 * NB + j];
      }
  }
}

__global__ void __launch_bounds__(64)
k_mass_inverses(int64_t ncells, const double* __restrict__ tab, const double* __restrict__ coords,
                const int32_t* __restrict__ cell_nodes, double* __restrict__ minv_p, double* __restrict__ minv_f) {
  const int64_t cell = (int64_t)blockIdx.x * 64 + threadIdx.x;
  if (cell >= ncells) return;
  double X[8][3];
  for (int v = 0; v < 8; v++)
    for (int i = 0; i < 3; i++) X[v][i] = coords[(int64_t)cell_nodes[cell * 8 + v] * 3 + i];
  double Mp[4][8], Mf[8][16];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) Mp[i][j] = (j - 4 == i) ? 1.0 : 0.0;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 16; j++) Mf[i][j] = (j - 8 == i) ? 1.0 : 0.0;
  for (int q = 0; q < 27; q++) {
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int v = 0; v < 8; v++)
      for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) J[i][k] += X[v][i] * tab[TM_GG + (q * 8 + v) * 3 + k];
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                       J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    const double w = tab[TM_W + q] * fabs(det);
    for (int k = 0; k < 4; k++)
      for (int l = 0; l < 4; l++) Mp[k][l] += w * tab[TM_PP + q * 4 + k] * tab[TM_PP + q * 4 + l];
    for (int k = 0; k < 8; k++)
      for (int l = 0; l < 8; l++) Mf[k][l] += w * tab[TM_CHI + q * 8 + k] * tab[TM_CHI + q * 8 + l];
  }
  invert_spd<4>(&Mp[0][0]);
  invert_spd<8>(&Mf[0][0]);
  for (int k = 0; k < 4; k++) for (int l = 0; l < 4; l++) minv_p[cell * 16 + k * 4 + l] = Mp[k][4 + l];
  for (int k = 0; k < 8; k++) for (int l = 0; l < 8; l++) minv_f[cell * 64 + k * 8 + l] = Mf[k][8 + l];
}

// z[rows of the cell's p and phi dofs] = (1/alpha) M^{-1} v
__global__ void __launch_bounds__(128)
k_apply_mass_inverses(int64_t ncells, int64_t nrows, const int32_t* __restrict__ gids, const double* __restrict__ minv_p,
                      const double* __restrict__ minv_f, double inv_alpha_p, double inv_alpha_f,
                      const double* __restrict__ v, double* __restrict__ z) {
  const int64_t t = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int64_t cell = t / 12;
  const int r = (int)(t - cell * 12);
  if (cell >= ncells) return;
  const int32_t* g = gids + cell * NLOC;
  if (r < 4) {
    const int32_t row = g[OFF_P + r];
    if (row < 0 || row >= nrows) return;
    double s = 0.0;
    for (int l = 0; l < 4; l++) s = fma(minv_p[cell * 16 + r * 4 + l], v[g[OFF_P + l]], s);
    z[row] = inv_alpha_p * s;
  } else {
    const int k = r - 4;
    const int32_t row = g[OFF_F + k];
    if (row < 0 || row >= nrows) return;
    double s = 0.0;
    for (int l = 0; l < 8; l++) s = fma(minv_f[cell * 64 + k * 8 + l], v[g[OFF_F + l]], s);
    z[row] = inv_alpha_f * s;
  }
}

// ---- H1-H1 block preconditioner (src/Solvers/h1h1blocks.jl:2-43): inverse P1disc cell mass blocks from the H1-H1 tables
__global__ void __launch_bounds__(64)
k_h1h1_mass_inverse_p(int64_t ncells, const double* __restrict__ tab, const double* __restrict__ coords,
                      const int32_t* __restrict__ cell_nodes, double* __restrict__ minv_p) {
  const int64_t cell = (int64_t)blockIdx.x * 64 + threadIdx.x;
  if (cell >= ncells) return;
  double X[8][3];
  for (int v = 0; v < 8; v++)
    for (int i = 0; i < 3; i++) X[v][i] = coords[(int64_t)cell_nodes[cell * 8 + v] * 3 + i];
  double Mp[4][8];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) Mp[i][j] = (j - 4 == i) ? 1.0 : 0.0;
  for (int q = 0; q < 27; q++) {
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int v = 0; v < 8; v++)
      for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) J[i][k] = __dadd_rn(__dmul_rn(X[v][i], tab[h1::T_GG + (q * 8 + v) * 3 + k]), J[i][k]);  // see hdiv7_cell.h
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                       J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    const double w = tab[h1::T_W + q] * fabs(det);
    for (int k = 0; k < 4; k++)
      for (int l = 0; l < 4; l++) Mp[k][l] += w * tab[h1::T_PP + q * 4 + k] * tab[h1::T_PP + q * 4 + l];
  }
  invert_spd<4>(&Mp[0][0]);
  for (int k = 0; k < 4; k++) for (int l = 0; l < 4; l++) minv_p[cell * 16 + k * 4 + l] = Mp[k][4 + l];
}

// z[p rows of the cell] = (1/alpha_p) M_p^{-1} v
__global__ void __launch_bounds__(128)
k_h1h1_apply_mass_p(int64_t ncells, int64_t nrows, const int32_t* __restrict__ gids, const double* __restrict__ minv_p,
                    double inv_alpha_p, const double* __restrict__ v, double* __restrict__ z) {
  const int64_t t = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int64_t cell = t >> 2;
  const int r = (int)(t & 3);
  if (cell >= ncells) return;
  const int32_t* g = gids + cell * h1::NLOC + h1::OFF_P;
  const int32_t row = g[r];
  if (row < 0 || row >= nrows) return;
  double s = 0.0;
  for (int l = 0; l < 4; l++) s = fma(minv_p[cell * 16 + r * 4 + l], v[g[l]], s);
  z[row] = inv_alpha_p * s;
}

}  // namespace mhd

extern "C" {

int mhd_solver_default_opts(mhd_solver_opts_t* o) {
  MHD_CHECK(o != nullptr, MHD_E_INVALID, "null opts");
  o->m = 15;             // niter_ls (src/parameters.jl:268)
  o->maxiter = 15;       // FGMRESSolver(m,P;maxiter=m) (badia2024.jl:40)
  o->rtol = 1e-7;        // nl_rtol/10 (badia2024.jl:37) with nl_rtol = 1e-6 (parameters.jl:269)
  o->atol = 1e-8;        // parameters.jl:270
  o->precond = MHD_PC_BLOCK_TRI;
  o->uj_inner_its = 30;
  o->uj_inner_restart = 30;
  o->alpha_p = -1.0;     // -1/(beta+zeta_u); the host overrides with the actual fluid parameters
  o->alpha_phi = -1.0;   // -1/(1+zeta_j)
  o->uj_solver = MHD_UJ_GMRES_JACOBI;
  o->patch_its = 1;
  o->patch_omega = 1.0;
  return MHD_OK;
}

int mhd_solver_create(mhd_operator_t* op, const mhd_solver_opts_t* opts, mhd_solver_t** out) {
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CHECK(op && opts && out, MHD_E_INVALID, "mhd_solver_create: null argument");
  MHD_CHECK(op->has_symbolic, MHD_E_STATE, "mhd_solver_create: call mhd_operator_symbolic first");
  MHD_CHECK(opts->precond >= 0 && opts->precond <= 3, MHD_E_INVALID, "unknown preconditioner %d", opts->precond);
  MHD_CHECK(opts->m >= 1 && opts->m <= MAXM && opts->maxiter >= 1, MHD_E_INVALID, "bad m/maxiter");
  MHD_CUDA(cudaSetDevice(g_device));
  if (opts->precond == MHD_PC_H1H1_BLOCKS) {
    MHD_CHECK(op->formulation == FORM_H1H1, MHD_E_INVALID, "MHD_PC_H1H1_BLOCKS needs an operator from mhd_h1h1_operator_create");
    MHD_CHECK(op->field_order[0] == MHD_FIELD_U && op->field_order[1] == MHD_FIELD_P && op->field_order[2] == MHD_FIELD_PHI,
              MHD_E_INVALID, "MHD_PC_H1H1_BLOCKS needs field_order (u,p,phi) (_multi_field_style(::Val{:h1h1blocks}))");
    MHD_CHECK(g_nranks == 1, MHD_E_INVALID, "MHD_PC_H1H1_BLOCKS runs on one GPU");
    MHD_CHECK(opts->alpha_p != 0.0, MHD_E_INVALID, "alpha_p must be nonzero");
    MHD_CHECK(opts->uj_inner_its >= 1 && opts->uj_inner_restart >= 1 && opts->uj_inner_restart <= MAXM, MHD_E_INVALID,
              "bad inner iteration counts");
    MHD_CHECK(opts->uj_solver == MHD_UJ_GMRES_JACOBI || opts->uj_solver == MHD_UJ_GMRES_PATCH, MHD_E_INVALID,
              "MHD_PC_H1H1_BLOCKS: block solvers are MHD_UJ_GMRES_JACOBI or MHD_UJ_GMRES_PATCH");
    if (opts->uj_solver == MHD_UJ_GMRES_PATCH)
      MHD_CHECK(opts->patch_its >= 1 && opts->patch_its <= 100 && opts->patch_omega > 0.0, MHD_E_INVALID, "bad patch_its / patch_omega");
  }
  if (opts->precond == MHD_PC_BLOCK_TRI) {
    MHD_CHECK(op->formulation == FORM_HDIV, MHD_E_INVALID,
              "mhd_solver_create: the block-triangular preconditioner is built for H1-HDiv operators (use MHD_PC_JACOBI)");
    MHD_CHECK(op->field_order[0] == MHD_FIELD_U && op->field_order[1] == MHD_FIELD_J && op->field_order[2] == MHD_FIELD_P &&
                  op->field_order[3] == MHD_FIELD_PHI,
              MHD_E_INVALID, "block-triangular preconditioner needs field_order (u,j,p,phi) = ([u,j],p,phi) blocks");
    MHD_CHECK(opts->alpha_p != 0.0 && opts->alpha_phi != 0.0, MHD_E_INVALID, "alpha_p/alpha_phi must be nonzero");
    MHD_CHECK(opts->uj_inner_its >= 1 && opts->uj_inner_restart >= 1 && opts->uj_inner_restart <= MAXM, MHD_E_INVALID,
              "bad inner iteration counts");
    MHD_CHECK(opts->uj_solver == MHD_UJ_GMRES_JACOBI || opts->uj_solver == MHD_UJ_DENSE_LU ||
                  opts->uj_solver == MHD_UJ_GMRES_PATCH, MHD_E_INVALID, "unknown uj_solver");
    if (opts->uj_solver == MHD_UJ_GMRES_PATCH) {
      MHD_CHECK(g_nranks == 1, MHD_E_INVALID, "the patch smoother runs on one GPU (patches across ranks are not exchanged yet)");
      MHD_CHECK(opts->patch_its >= 1 && opts->patch_its <= 100 && opts->patch_omega > 0.0, MHD_E_INVALID,
                "bad patch_its / patch_omega");
    }
    if (opts->uj_solver == MHD_UJ_DENSE_LU) {
      MHD_CHECK(g_nranks == 1, MHD_E_INVALID, "MHD_UJ_DENSE_LU is a single-GPU option (the (u,j) block is distributed)");
      MHD_CHECK(op->nowned[MHD_FIELD_U] + op->nowned[MHD_FIELD_J] <= 24576, MHD_E_CAPACITY,
                "MHD_UJ_DENSE_LU: (u,j) block of %lld rows exceeds the dense limit 24576",
                (long long)(op->nowned[MHD_FIELD_U] + op->nowned[MHD_FIELD_J]));
      MHD_TRY(load_cusolver());
    }
  }
  mhd_solver* s = new mhd_solver();
  s->op = op;
  s->opts = *opts;
  s->n_uj = op->nowned[MHD_FIELD_U] + op->nowned[MHD_FIELD_J];
  int rc = 0;
#define CR(x) if (!rc) rc = (x)
  CR(s->outer.init(op, op->nrows, op->ncols, opts->m, true));
  CR(dev_alloc(&s->d_dinv, op->nrows));
  CR(dev_alloc(&s->d_b, op->ncols));
  CR(dev_alloc(&s->d_x, op->ncols));
  CR(dev_alloc(&s->d_t1, op->ncols));
  CR(dev_alloc(&s->d_t2, op->ncols));
  CR(dev_alloc(&s->d_t3, op->ncols));
  if (!rc) {
    cudaMemsetAsync(s->d_x, 0, op->ncols * 8, g_stream);
    cudaMemsetAsync(s->d_b, 0, op->ncols * 8, g_stream);
    cudaMemsetAsync(s->d_t1, 0, op->ncols * 8, g_stream);
    cudaMemsetAsync(s->d_t2, 0, op->ncols * 8, g_stream);
    cudaMemsetAsync(s->d_t3, 0, op->ncols * 8, g_stream);
  }
  if (opts->precond == MHD_PC_H1H1_BLOCKS) {
    const int mi = opts->uj_inner_restart < opts->uj_inner_its ? opts->uj_inner_restart : opts->uj_inner_its;
    s->off_p = op->nowned[MHD_FIELD_U];
    s->off_phi = s->off_p + op->nowned[MHD_FIELD_P];
    CR(s->inner.init(op, s->n_uj, op->ncols, mi, false));     // u block: leading rows (n_uj = n_u, there is no j field)
    CR(s->inner2.init(op, op->nrows, op->ncols, mi, false));  // phi block: full-length vectors, zero outside phi
    CR(dev_alloc(&s->d_t4, op->ncols));
    CR(dev_alloc(&s->d_t5, op->ncols));
    CR(dev_alloc(&s->d_minv_p, op->ncells * 16));
    if (!rc) {
      k_h1h1_mass_inverse_p<<<(unsigned)((op->ncells + 63) / 64), 64, 0, g_stream>>>(op->ncells, op->d_tables, op->d_coords,
                                                                                     op->d_cell_nodes, s->d_minv_p);
      g_launches++;
      if (cudaPeekAtLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "k_h1h1_mass_inverse_p", __FILE__, __LINE__);
    }
  }
  if (opts->precond == MHD_PC_BLOCK_TRI) {
    const int mi = opts->uj_inner_restart < opts->uj_inner_its ? opts->uj_inner_restart : opts->uj_inner_its;
    CR(s->inner.init(op, s->n_uj, op->ncols, mi, false));
    if (opts->uj_solver == MHD_UJ_GMRES_PATCH) {
      CR(dev_alloc(&s->d_t4, op->ncols));
      CR(dev_alloc(&s->d_t5, op->ncols));
    }
    if (opts->uj_solver == MHD_UJ_DENSE_LU) {
      CR(dev_alloc(&s->d_dense, s->n_uj * s->n_uj));
      CR(dev_alloc(&s->d_ipiv, s->n_uj));
      CR(dev_alloc(&s->d_info, 1));
      if (!rc && g_cs.Create(&s->cs) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnCreate failed"); rc = MHD_E_CUDA; }
      if (!rc && g_cs.DgetrfBufferSize(s->cs, (int)s->n_uj, (int)s->n_uj, s->d_dense, (int)s->n_uj, &s->lu_lwork) != CUSOLVER_STATUS_SUCCESS) {
        set_error("cusolverDnDgetrf_bufferSize failed"); rc = MHD_E_CUDA;
      }
      CR(dev_alloc(&s->d_lu_work, s->lu_lwork));
    }
    CR(dev_alloc(&s->d_minv_p, op->ncells * 16));
    CR(dev_alloc(&s->d_minv_f, op->ncells * 64));
    if (!rc) {
      k_mass_inverses<<<(unsigned)((op->ncells + 63) / 64), 64, 0, g_stream>>>(op->ncells, op->d_tables, op->d_coords,
                                                                               op->d_cell_nodes, s->d_minv_p, s->d_minv_f);
      g_launches++;
      if (cudaPeekAtLastError() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "k_mass_inverses", __FILE__, __LINE__);
    }
  }
#undef CR
  if (!rc && cudaStreamSynchronize(g_stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "sync", __FILE__, __LINE__);
  if (rc) {
    mhd_solver_destroy(s);
    return rc;
  }
  *out = s;
  return MHD_OK;
}

int mhd_solver_destroy(mhd_solver_t* s) {
  if (!s) return MHD_OK;
  if (g_device >= 0) cudaSetDevice(g_device);
  cudaStreamSynchronize(g_stream);
  s->outer.release();
  s->inner.release();
  s->inner2.release();
  patch_destroy(s->patches_phi);
  if (s->cs) g_cs.Destroy(s->cs);
  cudaFree(s->d_dense); cudaFree(s->d_lu_work); cudaFree(s->d_ipiv); cudaFree(s->d_info);
  cudaFree(s->d_dinv); cudaFree(s->d_minv_p); cudaFree(s->d_minv_f);
  cudaFree(s->d_b); cudaFree(s->d_x); cudaFree(s->d_t1); cudaFree(s->d_t2); cudaFree(s->d_t3);
  cudaFree(s->d_t4); cudaFree(s->d_t5);
  patch_destroy(s->patches);
  delete s;
  return MHD_OK;
}

int mhd_solver_set_patches(mhd_solver_t* s, int64_t npatch, const int64_t* patch_ptr, const int32_t* patch_dofs) {
  MHD_CHECK(s != nullptr, MHD_E_INVALID, "null solver");
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CHECK((s->opts.precond == MHD_PC_BLOCK_TRI || s->opts.precond == MHD_PC_H1H1_BLOCKS) &&
                s->opts.uj_solver == MHD_UJ_GMRES_PATCH, MHD_E_STATE,
            "mhd_solver_set_patches: the solver was not created with uj_solver = MHD_UJ_GMRES_PATCH");
  MHD_CUDA(cudaSetDevice(g_device));
  patch_destroy(s->patches);
  s->patches = nullptr;
  s->setup_done = false;
  return patch_create(&s->patches, s->n_uj, npatch, patch_ptr, patch_dofs);
}

int mhd_solver_set_phi_patches(mhd_solver_t* s, int64_t npatch, const int64_t* patch_ptr, const int32_t* patch_dofs) {
  MHD_CHECK(s != nullptr, MHD_E_INVALID, "null solver");
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CHECK(s->opts.precond == MHD_PC_H1H1_BLOCKS && s->opts.uj_solver == MHD_UJ_GMRES_PATCH, MHD_E_STATE,
            "mhd_solver_set_phi_patches: needs precond = MHD_PC_H1H1_BLOCKS and uj_solver = MHD_UJ_GMRES_PATCH");
  MHD_CUDA(cudaSetDevice(g_device));
  patch_destroy(s->patches_phi);
  s->patches_phi = nullptr;
  s->setup_done = false;
  MHD_CHECK(npatch > 0 && patch_ptr != nullptr, MHD_E_INVALID, "mhd_solver_set_phi_patches: bad arguments");
  const int64_t ndofs = patch_ptr[npatch];
  MHD_CHECK(ndofs == 0 || patch_dofs != nullptr, MHD_E_INVALID, "mhd_solver_set_phi_patches: null dof list");
  for (int64_t i = 0; i < ndofs; i++)
    MHD_CHECK(patch_dofs[i] >= s->off_phi, MHD_E_INVALID, "phi patches: dof %d is not a phi row (rows start at %lld)", patch_dofs[i],
              (long long)s->off_phi);
  return patch_create(&s->patches_phi, s->op->nrows, npatch, patch_ptr, patch_dofs);
}

int mhd_solver_setup(mhd_solver_t* s) {
  MHD_CHECK(s != nullptr, MHD_E_INVALID, "null solver");
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CUDA(cudaSetDevice(g_device));
  mhd_operator* op = s->op;
  // the captured cycles hold pointers into preconditioner data that a new setup may reallocate
  s->outer.drop_graphs();
  s->inner.drop_graphs();
  s->inner2.drop_graphs();
  k_extract_dinv<<<vgrid(op->nrows), 256, 0, g_stream>>>(op->nrows, op->d_rowptr, op->d_colval, op->d_nzval, s->d_dinv);
  MHD_LAUNCH_CHECK();
  if (s->opts.precond == MHD_PC_BLOCK_TRI && s->opts.uj_solver == MHD_UJ_DENSE_LU) {
    const int64_t n = s->n_uj;
    MHD_CUDA(cudaMemsetAsync(s->d_dense, 0, (size_t)n * n * 8, g_stream));
    k_csr_block_to_dense<<<(unsigned)((n + 7) / 8), 256, 0, g_stream>>>(n, op->d_rowptr, op->d_colval, op->d_nzval, s->d_dense);
    MHD_LAUNCH_CHECK();
    MHD_CUSOLVER(g_cs.SetStream(s->cs, g_stream));
    MHD_CUSOLVER(g_cs.Dgetrf(s->cs, (int)n, (int)n, s->d_dense, (int)n, s->d_lu_work, s->d_ipiv, s->d_info));
    int info = 0;
    MHD_TRY(d2h(&info, s->d_info, 1));
    MHD_CUDA(cudaStreamSynchronize(g_stream));
    MHD_CHECK(info == 0, MHD_E_INVALID, "dense LU of the (u,j) block failed: getrf info = %d", info);
  }
  if ((s->opts.precond == MHD_PC_BLOCK_TRI || s->opts.precond == MHD_PC_H1H1_BLOCKS) && s->opts.uj_solver == MHD_UJ_GMRES_PATCH) {
    MHD_CHECK(s->patches != nullptr, MHD_E_STATE, "mhd_solver_setup: call mhd_solver_set_patches first (uj_solver = MHD_UJ_GMRES_PATCH)");
    MHD_TRY(patch_setup(s->patches, op));
    if (s->opts.precond == MHD_PC_H1H1_BLOCKS) {
      MHD_CHECK(s->patches_phi != nullptr, MHD_E_STATE, "mhd_solver_setup: call mhd_solver_set_phi_patches first");
      MHD_TRY(patch_setup(s->patches_phi, op));
    }
  }
  s->setup_done = true;
  return MHD_OK;
}

int mhd_solver_patch_apply(mhd_solver_t* s, const double* r, double* z, double omega) {
  MHD_CHECK(s && r && z, MHD_E_INVALID, "mhd_solver_patch_apply: null argument");
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CHECK(s->patches != nullptr && s->setup_done, MHD_E_STATE,
            "mhd_solver_patch_apply: call mhd_solver_set_patches and mhd_solver_setup first");
  MHD_CUDA(cudaSetDevice(g_device));
  const int64_t nuj = s->n_uj;
  const bool rdev = is_device_ptr(r), zdev = is_device_ptr(z);
  const double* dr = r;
  if (!rdev) {
    MHD_TRY(h2d(s->d_t4, r, nuj));
    dr = s->d_t4;
  }
  double* dz = zdev ? z : s->d_t5;
  MHD_TRY(patch_apply(s->patches, dr, dz, omega, false));
  if (!zdev) {
    MHD_TRY(d2h(z, dz, nuj));
    MHD_CUDA(cudaStreamSynchronize(g_stream));
  }
  return MHD_OK;
}

static int solve_impl(mhd_solver_t* s, const double* b, double* x, int32_t* iters, double* resnorm, double* res_history);

// The restart cycles are replayed as CUDA graphs, and the legacy default stream cannot be captured: when the caller has not
// set a stream (mhd_set_stream), the solve runs on a library-owned BLOCKING stream -- created with cudaStreamCreate, so it
// is ordered against everything the caller enqueued on the legacy stream before and enqueues after (legacy-stream semantics).
int mhd_solve(mhd_solver_t* s, const double* b, double* x, int32_t* iters, double* resnorm, double* res_history) {
  MHD_CHECK(s && b && x, MHD_E_INVALID, "mhd_solve: null argument");
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CHECK(s->setup_done, MHD_E_STATE, "mhd_solve: call mhd_solver_setup after mhd_jacobian");
  MHD_CUDA(cudaSetDevice(g_device));
  static cudaStream_t own = nullptr;
  static int own_device = -1;
  cudaStream_t saved = g_stream;
  if (g_stream == 0 && g_nranks == 1) {
    if (own_device != g_device) {
      MHD_CUDA(cudaStreamCreate(&own));
      own_device = g_device;
    }
    g_stream = own;
  }
  // Keep the Krylov basis of the outer FGMRES in L2 across the SpMVs: every product streams the whole matrix (1.8 GB at cfg2)
  // through the 126 MB L2 and evicts the basis, which the two Gram-Schmidt passes of the next step then fetch from HBM again.
  // A persisting access-policy window over V (hit ratio scaled to the set-aside the device allows) keeps it resident; the
  // window travels into the captured cycle graphs as a kernel-node attribute.  OFF by default (MHD_KRYLOV_L2=1 enables): measured
  // on cfg2 it LOSES -- 0.430 against 0.405 ms per iteration, the cfg2 solve 6.2 s against 4.6 s: the 94 MB set-aside takes the
  // L2 away from the x gathers of the SpMV and from the patch smoother's inverses.
  static int l2_want = -1;
  if (l2_want < 0) {
    const char* e = getenv("MHD_KRYLOV_L2");
    l2_want = e ? atoi(e) : 0;
  }
  bool window = false;
  if (l2_want && s->outer.V != nullptr) {
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, g_device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, g_device);
    const size_t vbytes = (size_t)(s->outer.m + 1) * (size_t)s->outer.ld * sizeof(double);
    if (max_persist > 0 && max_window > 0) {
      const size_t setaside = vbytes < (size_t)max_persist ? vbytes : (size_t)max_persist;
      const size_t wbytes = vbytes < (size_t)max_window ? vbytes : (size_t)max_window;
      cudaStreamAttrValue a;
      memset(&a, 0, sizeof(a));
      a.accessPolicyWindow.base_ptr = s->outer.V;
      a.accessPolicyWindow.num_bytes = wbytes;
      a.accessPolicyWindow.hitRatio = setaside >= wbytes ? 1.0f : (float)((double)setaside / (double)wbytes);
      a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setaside) == cudaSuccess &&
          cudaStreamSetAttribute(g_stream, cudaStreamAttributeAccessPolicyWindow, &a) == cudaSuccess)
        window = true;
      else
        cudaGetLastError();  // optional optimisation: carry on without it
    }
  }
  const int rc = solve_impl(s, b, x, iters, resnorm, res_history);
  if (window) {
    cudaStreamAttrValue a;
    memset(&a, 0, sizeof(a));
    cudaStreamSetAttribute(g_stream, cudaStreamAttributeAccessPolicyWindow, &a);
    cudaCtxResetPersistingL2Cache();
    cudaGetLastError();
  }
  g_stream = saved;
  return rc;
}

static int solve_impl(mhd_solver_t* s, const double* b, double* x, int32_t* iters, double* resnorm, double* res_history) {
  mhd_operator* op = s->op;
  const int64_t n = op->nrows;
  const mhd_solver_opts_t& o = s->opts;
  const bool bdev = is_device_ptr(b), xdev = is_device_ptr(x);
  // the solve works on library-owned vectors with a ghost section
  MHD_CUDA(cudaMemcpyAsync(s->d_b, b, n * 8, bdev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, g_stream));
  MHD_CUDA(cudaMemcpyAsync(s->d_x, x, n * 8, xdev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, g_stream));

  VecOp matvec = [op](const double* v, double* y) -> int {
    return spmv_with_halo(op, op->nrows, const_cast<double*>(v), y);
  };
  // (u,j)-block operator: rows [0,n_uj), p/phi entries of the input are kept at zero by construction
  const int64_t nuj = s->n_uj;
  VecOp matvec_uj = [op, nuj](const double* v, double* y) -> int {
    return spmv_with_halo(op, nuj, const_cast<double*>(v), y);
  };
  double* dinv = s->d_dinv;
  VecOp jacobi_uj = [dinv, nuj](const double* v, double* z) -> int {
    k_mul<<<vgrid(nuj), 256, 0, g_stream>>>(nuj, dinv, v, z);
    MHD_LAUNCH_CHECK();
    return 0;
  };
  // vertex-patch smoother: z = Richardson(patch_its, patch_omega) on A_uj with the additive patch solver (gmg.jl:62-81)
  VecOp patch_uj = [s, nuj, &matvec_uj](const double* v, double* z) -> int {
    const mhd_solver_opts_t& oo = s->opts;
    MHD_TRY(patch_apply(s->patches, v, z, oo.patch_omega, false));
    for (int it = 1; it < oo.patch_its; it++) {
      MHD_TRY(matvec_uj(z, s->d_t4));
      k_sub<<<vgrid(nuj), 256, 0, g_stream>>>(nuj, v, s->d_t4, s->d_t5);
      MHD_LAUNCH_CHECK();
      MHD_TRY(patch_apply(s->patches, s->d_t5, z, oo.patch_omega, true));
    }
    return 0;
  };
  VecOp precond;
  if (o.precond == MHD_PC_NONE) {
    precond = [n](const double* v, double* z) -> int {
      MHD_CUDA(cudaMemcpyAsync(z, v, n * 8, cudaMemcpyDeviceToDevice, g_stream));
      return 0;
    };
  } else if (o.precond == MHD_PC_JACOBI) {
    precond = [dinv, n](const double* v, double* z) -> int {
      k_mul<<<vgrid(n), 256, 0, g_stream>>>(n, dinv, v, z);
      MHD_LAUNCH_CHECK();
      return 0;
    };
  } else if (o.precond == MHD_PC_H1H1_BLOCKS) {
    // H1H1BlockSolver (src/Solvers/h1h1blocks.jl:17-26): BlockTriangularSolver(:upper) over (u, p, phi) with coefficients
    // [1 1 1; 0 1 0; 0 0 1]: z_phi = S_phi v_phi ; z_p = (alpha_p M_p)^{-1} v_p ; z_u = S_u (v_u - A_up z_p - A_{u phi} z_phi),
    // S_u / S_phi = inner GMRES on the assembled diagonal blocks, preconditioned by their vertex-patch solvers (or Jacobi)
    precond = [s, op, n, nuj, &matvec_uj, &jacobi_uj, &patch_uj](const double* v, double* z) -> int {
      const mhd_solver_opts_t& oo = s->opts;
      const int64_t off = s->off_phi, nphi = n - off;
      const bool use_patch = oo.uj_solver == MHD_UJ_GMRES_PATCH;
      VecOp matvec_phi = [op, off, n](const double* x, double* y) -> int {
        MHD_CUDA(cudaMemsetAsync(y, 0, (size_t)off * 8, g_stream));
        return spmv_row_range(op, off, n, x, y);
      };
      double* dinv = s->d_dinv;
      VecOp jacobi_phi = [dinv, off, n](const double* x, double* y) -> int {
        MHD_CUDA(cudaMemsetAsync(y, 0, (size_t)off * 8, g_stream));
        k_mul<<<vgrid(n - off), 256, 0, g_stream>>>(n - off, dinv + off, x + off, y + off);
        MHD_LAUNCH_CHECK();
        return 0;
      };
      VecOp patch_phi = [s, n, &matvec_phi](const double* x, double* y) -> int {
        const mhd_solver_opts_t& o2 = s->opts;
        MHD_TRY(patch_apply(s->patches_phi, x, y, o2.patch_omega, false));
        for (int it = 1; it < o2.patch_its; it++) {
          MHD_TRY(matvec_phi(y, s->d_t4));
          k_sub<<<vgrid(n), 256, 0, g_stream>>>(n, x, s->d_t4, s->d_t5);
          MHD_LAUNCH_CHECK();
          MHD_TRY(patch_apply(s->patches_phi, s->d_t5, y, o2.patch_omega, true));
        }
        return 0;
      };
      MHD_CUDA(cudaMemsetAsync(z, 0, (size_t)op->ncols * 8, g_stream));
      // phi block
      if (nphi > 0) {
        MHD_CUDA(cudaMemsetAsync(s->d_t2, 0, (size_t)op->ncols * 8, g_stream));
        MHD_CUDA(cudaMemcpyAsync(s->d_t2 + off, v + off, (size_t)nphi * 8, cudaMemcpyDeviceToDevice, g_stream));
        MHD_CUDA(cudaMemsetAsync(s->d_t3, 0, (size_t)op->ncols * 8, g_stream));
        int done_its = 0;
        bool first = true;
        while (done_its < oo.uj_inner_its) {
          MHD_TRY(s->inner2.cycle(matvec_phi, use_patch ? patch_phi : jacobi_phi, s->d_t2, s->d_t3, 1e-2, 0.0, oo.uj_inner_its, first, true));
          done_its += s->inner2.m;
          first = false;
        }
        MHD_CUDA(cudaMemcpyAsync(z + off, s->d_t3 + off, (size_t)nphi * 8, cudaMemcpyDeviceToDevice, g_stream));
      }
      // p block
      k_h1h1_apply_mass_p<<<(unsigned)((op->ncells * 4 + 127) / 128), 128, 0, g_stream>>>(op->ncells, n, op->d_gids, s->d_minv_p,
                                                                                          1.0 / oo.alpha_p, v, z);
      MHD_LAUNCH_CHECK();
      // u block with the coupling to p and phi
      MHD_TRY(spmv_with_halo(op, nuj, z, s->d_t1));
      k_sub<<<vgrid(nuj), 256, 0, g_stream>>>(nuj, v, s->d_t1, s->d_t2);
      MHD_LAUNCH_CHECK();
      MHD_CUDA(cudaMemsetAsync(s->d_t3, 0, (size_t)op->ncols * 8, g_stream));
      int done_its = 0;
      bool first = true;
      while (done_its < oo.uj_inner_its) {
        MHD_TRY(s->inner.cycle(matvec_uj, use_patch ? patch_uj : jacobi_uj, s->d_t2, s->d_t3, 1e-2, 0.0, oo.uj_inner_its, first, true));
        done_its += s->inner.m;
        first = false;
      }
      MHD_CUDA(cudaMemcpyAsync(z, s->d_t3, (size_t)nuj * 8, cudaMemcpyDeviceToDevice, g_stream));
      return 0;
    };
  } else {
    // upper block-triangular solve (badia2024.jl:25-31): phi, p first, then the (u,j) block with the coupling
    precond = [s, op, n, nuj, &matvec_uj, &jacobi_uj, &patch_uj](const double* v, double* z) -> int {
      const mhd_solver_opts_t& oo = s->opts;
      MHD_CUDA(cudaMemsetAsync(z, 0, (size_t)op->ncols * 8, g_stream));
      k_apply_mass_inverses<<<(unsigned)((op->ncells * 12 + 127) / 128), 128, 0, g_stream>>>(
          op->ncells, n, op->d_gids, s->d_minv_p, s->d_minv_f, 1.0 / oo.alpha_p, 1.0 / oo.alpha_phi, v, z);
      MHD_LAUNCH_CHECK();
      // t1 = A [0; z_p; z_phi] restricted to the (u,j) rows ; rhs = v_uj - t1
      MHD_TRY(spmv_with_halo(op, nuj, z, s->d_t1));
      k_sub<<<vgrid(nuj), 256, 0, g_stream>>>(nuj, v, s->d_t1, s->d_t2);
      MHD_LAUNCH_CHECK();
      if (oo.uj_solver == MHD_UJ_DENSE_LU) {
        // exact solve with the LU factors (in place on the right-hand side)
        MHD_CUSOLVER(g_cs.SetStream(s->cs, g_stream));
        MHD_CUSOLVER(g_cs.Dgetrs(s->cs, CUBLAS_OP_N, (int)nuj, 1, s->d_dense, (int)nuj, s->d_ipiv, s->d_t2, (int)nuj, s->d_info));
        MHD_CUDA(cudaMemcpyAsync(z, s->d_t2, (size_t)nuj * 8, cudaMemcpyDeviceToDevice, g_stream));
        return 0;
      }
      // inner GMRES on A_uj with Jacobi (zero initial guess); d_t3 = solution with zero p/phi/ghost entries
      MHD_CUDA(cudaMemsetAsync(s->d_t3, 0, (size_t)op->ncols * 8, g_stream));
      int done_its = 0;
      bool first = true;
      while (done_its < oo.uj_inner_its) {
        MHD_TRY(s->inner.cycle(matvec_uj, oo.uj_solver == MHD_UJ_GMRES_PATCH ? patch_uj : jacobi_uj, s->d_t2, s->d_t3, 1e-2, 0.0,
                               oo.uj_inner_its, first, true));
        done_its += s->inner.m;
        first = false;
      }
      MHD_CUDA(cudaMemcpyAsync(z, s->d_t3, (size_t)nuj * 8, cudaMemcpyDeviceToDevice, g_stream));
      return 0;
    };
  }

  KrylovScalars hs;
  int total = 0;
  bool first = true;
  int rc = 0;
  while (true) {
    rc = s->outer.cycle(matvec, precond, s->d_b, s->d_x, o.rtol, o.atol, o.maxiter, first, false);
    if (rc) return rc;
    first = false;
    MHD_CUDA(cudaMemcpyAsync(&hs, s->outer.S, sizeof(KrylovScalars), cudaMemcpyDeviceToHost, g_stream));
    MHD_CUDA(cudaStreamSynchronize(g_stream));
    total = hs.iters;
    if (hs.beta <= hs.tol || total >= o.maxiter || hs.k == 0) break;
  }
  if (iters) *iters = total;
  if (resnorm) *resnorm = hs.beta;
  if (res_history) {
    const int nh = total + 1 < 4 * MAXM + 2 ? total + 1 : 4 * MAXM + 2;
    for (int i = 0; i < nh; i++) res_history[i] = hs.hist[i];
  }
  MHD_CUDA(cudaMemcpyAsync(x, s->d_x, n * 8, xdev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, g_stream));
  MHD_CUDA(cudaStreamSynchronize(g_stream));
  MHD_TRY(halo_check(s->op));  // a ghost exchange that timed out inside the cycles invalidates the solve
  if (hs.beta > hs.tol) {
    set_error("FGMRES stopped at %d iterations with residual %.3e > tol %.3e", total, hs.beta, hs.tol);
    return MHD_E_NOTCONV;
  }
  return MHD_OK;
}

}  // extern "C"
