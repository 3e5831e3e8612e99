// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// Replaces the MPI traffic PartitionedArrays generates for this path:
//   consistent!(x)  (owner -> ghost dof values before mul!)      -> halo_exchange: pack, grouped ncclSend/ncclRecv, unpack
//   dot / norm      (MPI_Allreduce of 1..m+1 doubles)            -> allreduce_sum on a device buffer, in-stream
// NCCL is bound lazily with dlopen so that a single-GPU run never needs it and the copy torch already loaded
// (same SONAME) is reused when the host process is Python.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include "common.h"

namespace mhd {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;

static int load_nccl() {
  if (g_nccl.lib) return 0;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  MHD_CHECK(lib != nullptr, MHD_E_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                         \
  do {                                                                           \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                                 \
    MHD_CHECK(g_nccl.field != nullptr, MHD_E_COMM, "libnccl lacks symbol %s", name); \
  } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.lib = lib;
  return 0;
}

#define MHD_NCCL(call)                                                                   \
  do {                                                                                   \
    ncclResult_t _r = (call);                                                            \
    if (_r != ncclSuccess) {                                                             \
      set_error("NCCL error %d (%s) in %s", (int)_r, g_nccl.GetErrorString(_r), #call);  \
      return MHD_E_COMM;                                                                 \
    }                                                                                    \
  } while (0)

__global__ void pack_kernel(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ x, double* __restrict__ buf) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) buf[i] = x[idx[i]];
}
__global__ void unpack_kernel(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ buf, double* __restrict__ x) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[idx[i]] = buf[i];
}

int halo_exchange(mhd_operator* op, double* d_x) {
  Halo& h = op->halo;
  if (g_nranks <= 1 || h.nneigh == 0) return 0;
  MHD_CHECK(g_comm != nullptr, MHD_E_STATE, "halo exchange requested before mhd_comm_init");
  if (h.nsend > 0) {
    pack_kernel<<<(unsigned)((h.nsend + 255) / 256), 256, 0, g_stream>>>(h.nsend, h.d_send_idx, d_x, h.d_send_buf);
    MHD_LAUNCH_CHECK();
  }
  MHD_NCCL(g_nccl.GroupStart());
  for (int k = 0; k < h.nneigh; k++) {
    const int64_t ns = h.send_ptr[k + 1] - h.send_ptr[k], nr = h.recv_ptr[k + 1] - h.recv_ptr[k];
    if (ns > 0) MHD_NCCL(g_nccl.Send(h.d_send_buf + h.send_ptr[k], (size_t)ns, ncclFloat64, h.ranks[k], g_comm, g_stream));
    if (nr > 0) MHD_NCCL(g_nccl.Recv(h.d_recv_buf + h.recv_ptr[k], (size_t)nr, ncclFloat64, h.ranks[k], g_comm, g_stream));
  }
  MHD_NCCL(g_nccl.GroupEnd());
  if (h.nrecv > 0) {
    unpack_kernel<<<(unsigned)((h.nrecv + 255) / 256), 256, 0, g_stream>>>(h.nrecv, h.d_recv_idx, h.d_recv_buf, d_x);
    MHD_LAUNCH_CHECK();
  }
  return 0;
}

int allreduce_sum(double* d_buf, int n) {
  if (g_nranks <= 1) return 0;
  MHD_CHECK(g_comm != nullptr, MHD_E_STATE, "all-reduce requested before mhd_comm_init");
  MHD_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclFloat64, ncclSum, g_comm, g_stream));
  return 0;
}

int halo_check(mhd_operator* op) {
  if (!op->halo.fused || op->halo.d_err == nullptr) return 0;
  int e = 0;
  MHD_CUDA(cudaMemcpyAsync(&e, op->halo.d_err, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  MHD_CUDA(cudaStreamSynchronize(g_stream));
  MHD_CHECK(e == 0, MHD_E_COMM,
            "fused SpMV + ghost exchange: a neighbour's values did not arrive within the spin limit (stalled or dead peer); "
            "results computed since are invalid");
  return 0;
}

}  // namespace mhd

using namespace mhd;

extern "C" {

int mhd_comm_get_unique_id(void* id128) {
  MHD_CHECK(id128 != nullptr, MHD_E_INVALID, "null id buffer");
  MHD_TRY(load_nccl());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  MHD_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return MHD_OK;
}

int mhd_comm_init(int rank, int nranks, const void* id128) {
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, MHD_E_INVALID, "bad rank %d / nranks %d", rank, nranks);
  if (nranks == 1) {
    g_nranks = 1;
    g_rank = 0;
    return MHD_OK;
  }
  MHD_CHECK(id128 != nullptr, MHD_E_INVALID, "null NCCL id");
  MHD_TRY(load_nccl());
  MHD_CUDA(cudaSetDevice(g_device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  MHD_NCCL(g_nccl.CommInitRank(&g_comm, nranks, id, rank));
  g_nranks = nranks;
  g_rank = rank;
  return MHD_OK;
}

int mhd_comm_finalize(void) {
  if (g_comm) {
    cudaStreamSynchronize(g_stream);
    g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
  }
  g_nranks = 1;
  g_rank = 0;
  return MHD_OK;
}

// ghost_lo[r] = position of the first ghost column of row r (columns are sorted and ghost ids come last: the ghost entries are
// the tail [ghost_lo, rowptr[r+1]) of the row; == rowptr[r+1] for rows without ghosts)
__global__ void tag_ghost_rows(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colval,
                               long long* __restrict__ ghost_lo) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) {
    if (r == nrows) ghost_lo[r] = rowptr[r];
    return;
  }
  int64_t a = rowptr[r], b = rowptr[r + 1];
  while (a < b) {  // first position with colval >= nrows
    const int64_t m = (a + b) >> 1;
    if (colval[m] >= nrows) b = m;
    else a = m + 1;
  }
  ghost_lo[r] = a;
}

// ---- fused peer-memory halo: export this rank's inbox, connect to the neighbours' inboxes
int mhd_operator_halo_ipc_export(mhd_operator_t* op, void* handle64) {
  MHD_CHECK(op != nullptr && handle64 != nullptr, MHD_E_INVALID, "mhd_operator_halo_ipc_export: null argument");
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CUDA(cudaSetDevice(g_device));
  Halo& h = op->halo;
  MHD_CHECK(h.nneigh > 0 && h.nneigh <= HALO_MAX_NEIGH, MHD_E_INVALID, "halo plan missing or more than %d neighbours", HALO_MAX_NEIGH);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  if (!h.ipc_mem) {
    const int64_t nghost = op->ncols - op->nrows;
    const size_t bytes = HALO_FLAG_BYTES + 2 * (size_t)(nghost > 0 ? nghost : 1) * sizeof(double);
    MHD_CUDA(cudaMalloc(&h.ipc_mem, bytes));
    MHD_CUDA(cudaMemset(h.ipc_mem, 0, bytes));
  }
  cudaIpcMemHandle_t hd;
  MHD_CUDA(cudaIpcGetMemHandle(&hd, h.ipc_mem));
  memcpy(handle64, &hd, 64);
  return MHD_OK;
}

int mhd_operator_halo_ipc_connect(mhd_operator_t* op, const void* handles, const int32_t* send_dst,
                                  const int32_t* peer_slot, const int64_t* peer_nghost) {
  MHD_CHECK(op && handles && send_dst && peer_slot && peer_nghost, MHD_E_INVALID, "mhd_operator_halo_ipc_connect: null argument");
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CUDA(cudaSetDevice(g_device));
  Halo& h = op->halo;
  MHD_CHECK(h.ipc_mem != nullptr, MHD_E_STATE, "call mhd_operator_halo_ipc_export first");
  HaloDev hd;
  memset(&hd, 0, sizeof(hd));
  hd.nneigh = h.nneigh;
  hd.nrecv = h.nrecv;
  h.peer_mem.assign(h.nneigh, nullptr);
  std::vector<int> push_neigh;
  std::vector<int64_t> push_begin;
  for (int k = 0; k < h.nneigh; k++) {
    cudaIpcMemHandle_t hk;
    memcpy(&hk, (const char*)handles + 64 * k, 64);
    MHD_CUDA(cudaIpcOpenMemHandle(&h.peer_mem[k], hk, cudaIpcMemLazyEnablePeerAccess));
    MHD_CHECK(peer_slot[k] >= 0 && peer_slot[k] < HALO_MAX_NEIGH, MHD_E_INVALID, "bad peer slot");
    char* base = (char*)h.peer_mem[k];
    for (int p = 0; p < 2; p++) {
      hd.peer_inbox[p][k] = (double*)(base + HALO_FLAG_BYTES) + (int64_t)p * (peer_nghost[k] > 0 ? peer_nghost[k] : 1);
      hd.peer_flag[p][k] = (unsigned*)base + p * HALO_MAX_NEIGH + peer_slot[k];
    }
    const int64_t nr = h.recv_ptr[k + 1] - h.recv_ptr[k];
    hd.expected[k] = (unsigned)((nr + HALO_CHUNK - 1) / HALO_CHUNK);
    hd.send_begin[k] = h.send_ptr[k];
    for (int64_t b = h.send_ptr[k]; b < h.send_ptr[k + 1]; b += HALO_CHUNK) {
      push_neigh.push_back(k);
      push_begin.push_back(b);
    }
  }
  hd.send_begin[h.nneigh] = h.send_ptr[h.nneigh];
  hd.npush = (int)push_neigh.size();
  for (int64_t i = 0; i < h.nsend; i++) {
    int k = 0;
    while (i >= h.send_ptr[k + 1]) k++;
    MHD_CHECK(send_dst[i] >= 0 && send_dst[i] < peer_nghost[k], MHD_E_INVALID, "send_dst[%lld]=%d outside the neighbour's ghost range", (long long)i, send_dst[i]);
  }
  int* d_push_neigh = nullptr;
  int64_t* d_push_begin = nullptr;
  MHD_TRY(dev_alloc(&d_push_neigh, (int64_t)push_neigh.size()));
  MHD_TRY(dev_alloc(&d_push_begin, (int64_t)push_begin.size()));
  MHD_TRY(dev_alloc(&h.d_ghost_src, h.nsend));
  MHD_TRY(dev_alloc(&h.d_err, 8));  // [0] time-out flag, [1..4] MHD_HALO_DEBUG statistics
  MHD_TRY(h2d(d_push_neigh, push_neigh.data(), (int64_t)push_neigh.size()));
  MHD_TRY(h2d(d_push_begin, push_begin.data(), (int64_t)push_begin.size()));
  MHD_TRY(h2d(h.d_ghost_src, send_dst, h.nsend));
  MHD_CUDA(cudaMemsetAsync(h.d_err, 0, 8 * sizeof(int), g_stream));
  hd.send_idx = h.d_send_idx;
  hd.push_neigh = d_push_neigh;
  hd.push_begin = d_push_begin;
  hd.my_flags = (const unsigned*)h.ipc_mem;
  hd.my_inbox[0] = (const double*)((char*)h.ipc_mem + HALO_FLAG_BYTES);
  hd.my_inbox[1] = hd.my_inbox[0] + (op->ncols - op->nrows > 0 ? op->ncols - op->nrows : 1);
  MHD_CHECK(op->has_symbolic, MHD_E_STATE, "mhd_operator_halo_ipc_connect: call mhd_operator_symbolic first");
  MHD_TRY(dev_alloc(&h.d_row_bits, op->nrows + 1));
  tag_ghost_rows<<<(unsigned)((op->nrows + 256) / 256), 256, 0, g_stream>>>(op->nrows, op->d_rowptr, op->d_colval, h.d_row_bits);
  MHD_LAUNCH_CHECK();
  {  // interface rows = rows with a non-empty ghost tail (setup-time host pass)
    std::vector<long long> gl((size_t)op->nrows + 1), rp((size_t)op->nrows + 1);
    MHD_CUDA(cudaMemcpyAsync(gl.data(), h.d_row_bits, gl.size() * sizeof(long long), cudaMemcpyDeviceToHost, g_stream));
    MHD_CUDA(cudaMemcpyAsync(rp.data(), op->d_rowptr, rp.size() * sizeof(long long), cudaMemcpyDeviceToHost, g_stream));
    MHD_CUDA(cudaStreamSynchronize(g_stream));
    std::vector<int32_t> ifr;
    for (int64_t r = 0; r < op->nrows; r++)
      if (gl[(size_t)r] < rp[(size_t)r + 1]) ifr.push_back((int32_t)r);
    cudaFree(h.d_if_rows);
    h.d_if_rows = nullptr;
    h.n_if_rows = (int64_t)ifr.size();
    if (h.n_if_rows > 0) {
      MHD_TRY(dev_alloc(&h.d_if_rows, h.n_if_rows));
      MHD_TRY(h2d(h.d_if_rows, ifr.data(), h.n_if_rows));
    }
  }
  hd.send_dst = h.d_ghost_src;
  hd.rowptr_tagged = h.d_row_bits;
  h.inbox[0] = hd.my_inbox[0];
  h.inbox[1] = hd.my_inbox[1];
  h.flags = hd.my_flags;
  hd.err = h.d_err;
  MHD_TRY(dev_alloc(&h.d_dev, 1));
  MHD_CUDA(cudaMemcpyAsync(h.d_dev, &hd, sizeof(HaloDev), cudaMemcpyHostToDevice, g_stream));
  MHD_CUDA(cudaStreamSynchronize(g_stream));
  h.epoch = 0;
  h.fused = getenv("MHD_HALO_NCCL") == nullptr;  // MHD_HALO_NCCL=1 keeps the NCCL send/recv path (A/B comparison)
  return MHD_OK;
}

int mhd_operator_halo_status(mhd_operator_t* op, int32_t* fused, int32_t* timed_out) {
  MHD_CHECK(op != nullptr, MHD_E_INVALID, "null operator");
  if (fused) *fused = op->halo.fused ? 1 : 0;
  if (timed_out) {
    *timed_out = 0;
    if (op->halo.d_err) {
      int e[8];
      MHD_CUDA(cudaMemcpy(e, op->halo.d_err, sizeof(e), cudaMemcpyDeviceToHost));
      *timed_out = e[0];
      if (getenv("MHD_HALO_DEBUG"))
        fprintf(stderr, "[mhd halo rank %d] products %u  max spin clocks %d  waiting CTAs (all products) %d  interface rows %lld  max push-CTA clocks %d\n", g_rank,
                op->halo.epoch, e[1], e[2], (long long)op->halo.n_if_rows, e[5]);
    }
  }
  return MHD_OK;
}

int mhd_operator_set_halo(mhd_operator_t* op, int32_t nneigh, const int32_t* neigh_ranks, const int64_t* send_ptr,
                          const int32_t* send_idx, const int64_t* recv_ptr, const int32_t* recv_idx) {
  MHD_CHECK(op != nullptr && nneigh >= 0, MHD_E_INVALID, "mhd_operator_set_halo: bad arguments");
  MHD_CHECK(g_device >= 0, MHD_E_STATE, "mhd_init has not been called");
  MHD_CUDA(cudaSetDevice(g_device));
  Halo& h = op->halo;
  cudaFree(h.d_send_idx); cudaFree(h.d_recv_idx); cudaFree(h.d_send_buf); cudaFree(h.d_recv_buf);
  h = Halo();
  h.nneigh = nneigh;
  if (nneigh == 0) return MHD_OK;
  MHD_CHECK(neigh_ranks && send_ptr && recv_ptr, MHD_E_INVALID, "mhd_operator_set_halo: null plan arrays");
  h.ranks.assign(neigh_ranks, neigh_ranks + nneigh);
  h.send_ptr.assign(send_ptr, send_ptr + nneigh + 1);
  h.recv_ptr.assign(recv_ptr, recv_ptr + nneigh + 1);
  h.nsend = send_ptr[nneigh];
  h.nrecv = recv_ptr[nneigh];
  for (int64_t i = 0; i < h.nsend; i++)
    MHD_CHECK(send_idx[i] >= 0 && send_idx[i] < op->nrows, MHD_E_INVALID, "send_idx[%lld]=%d is not an owned id", (long long)i, send_idx[i]);
  for (int64_t i = 0; i < h.nrecv; i++)
    MHD_CHECK(recv_idx[i] >= op->nrows && recv_idx[i] < op->ncols, MHD_E_INVALID, "recv_idx[%lld]=%d is not a ghost id", (long long)i, recv_idx[i]);
  MHD_TRY(dev_alloc(&h.d_send_idx, h.nsend));
  MHD_TRY(dev_alloc(&h.d_recv_idx, h.nrecv));
  MHD_TRY(dev_alloc(&h.d_send_buf, h.nsend));
  MHD_TRY(dev_alloc(&h.d_recv_buf, h.nrecv));
  MHD_TRY(h2d(h.d_send_idx, send_idx, h.nsend));
  MHD_TRY(h2d(h.d_recv_idx, recv_idx, h.nrecv));
  MHD_CUDA(cudaStreamSynchronize(g_stream));
  return MHD_OK;
}

}  // extern "C"
