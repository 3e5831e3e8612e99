// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// Replaces the MPI traffic PartitionedArrays generates for this path:
//   consistent!(x)  (owner -> ghost dof values before mul!)      -> halo_exchange: pack, grouped ncclSend/ncclRecv, unpack
//   dot / norm      (MPI_Allreduce of 1..m+1 doubles)            -> allreduce_sum on a device buffer, in-stream
// NCCL is bound lazily with dlopen so that a single-GPU run never needs it and the copy torch already loaded
// (same SONAME) is reused when the host process is Python.
#include <dlfcn.h>
#include <nccl.h>

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
