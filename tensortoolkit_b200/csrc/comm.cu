// comm.cu -- multi-GPU plumbing behind the C ABI (qlb200_comm_*): symmetric buffers over NVLink / NVSwitch, their multicast
// mapping, and a device-side barrier.  No torch, no NCCL: the CUDA virtual-memory-management driver API (cuMemCreate /
// cuMemExportToShareableHandle / cuMulticast*), file descriptors passed between the ranks of one node over abstract unix
// sockets (SCM_RIGHTS), and a caller-supplied all-gather callback for the few bytes of bootstrap (MPI_Allgather in a
// TensorToolkit program -- the reference distributes its DMRG mat-vec over MPI ranks, tensor_manipulation/dmrg/
// contract_1sector.h:181-228 -- or torch.distributed in the Python harness).  Ranks may be processes or threads.
//
// The reference has no multi-GPU path at all (README "CUDA Multi-Card Support" is an unchecked to-do); what this replaces is
// its MPI exchange of partial results: here every rank's output tiles are stored straight into a symmetric result buffer of
// every GPU from inside the GEMM epilogue (qlb200_execute_bcast / _mcast) and only a barrier remains.
#include <cuda.h>
#include <cuda_runtime.h>

#include <fcntl.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "plan.h"
#include "qlb200.h"

namespace qlb200 {
int FailWith(int code, const std::string &msg);   // capi.cu: sets qlb200_last_error
}
using qlb200::FailWith;

namespace {

// ---- driver API through the runtime's entry-point query (the library links cudart statically and never libcuda) ----
struct Driver {
  bool ok = false;
  std::string why;
#define QL_DRV(name) decltype(&name) p_##name = nullptr;
  QL_DRV(cuMemCreate) QL_DRV(cuMemRelease) QL_DRV(cuMemAddressReserve) QL_DRV(cuMemAddressFree) QL_DRV(cuMemMap) QL_DRV(cuMemUnmap)
  QL_DRV(cuMemSetAccess) QL_DRV(cuMemExportToShareableHandle) QL_DRV(cuMemImportFromShareableHandle)
  QL_DRV(cuMemGetAllocationGranularity) QL_DRV(cuMulticastCreate) QL_DRV(cuMulticastAddDevice) QL_DRV(cuMulticastBindMem)
  QL_DRV(cuMulticastGetGranularity) QL_DRV(cuDeviceGet) QL_DRV(cuDeviceGetAttribute) QL_DRV(cuGetErrorString)
#undef QL_DRV
  Driver() {
    auto get = [&](const char *sym, void **fp) {
      cudaDriverEntryPointQueryResult st;
      cudaError_t e = cudaGetDriverEntryPoint(sym, fp, cudaEnableDefault, &st);
      if (e != cudaSuccess || st != cudaDriverEntryPointSuccess || *fp == nullptr) { if (why.empty()) why = std::string("driver entry point missing: ") + sym; return false; }
      return true;
    };
    ok = true;
#define QL_GET(name) ok = get(#name, reinterpret_cast<void **>(&p_##name)) && ok;
    QL_GET(cuMemCreate) QL_GET(cuMemRelease) QL_GET(cuMemAddressReserve) QL_GET(cuMemAddressFree) QL_GET(cuMemMap) QL_GET(cuMemUnmap)
    QL_GET(cuMemSetAccess) QL_GET(cuMemExportToShareableHandle) QL_GET(cuMemImportFromShareableHandle)
    QL_GET(cuMemGetAllocationGranularity) QL_GET(cuMulticastCreate) QL_GET(cuMulticastAddDevice) QL_GET(cuMulticastBindMem)
    QL_GET(cuMulticastGetGranularity) QL_GET(cuDeviceGet) QL_GET(cuDeviceGetAttribute) QL_GET(cuGetErrorString)
#undef QL_GET
  }
  std::string Err(CUresult r) const {
    const char *s = nullptr;
    if (p_cuGetErrorString && p_cuGetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "CUDA driver error " + std::to_string(int(r));
  }
};
Driver &Drv() { static Driver d; return d; }

#define QL_DRVCALL(call)                                                                     \
  do {                                                                                       \
    CUresult r_ = Drv().p_##call;                                                            \
    if (r_ != CUDA_SUCCESS) return FailWith(QLB200_ERR_CUDA, std::string(#call) + ": " + Drv().Err(r_)); \
  } while (0)
#define QL_RT(call)                                                                          \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) return FailWith(QLB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// ---- file descriptors between ranks: abstract unix sockets + SCM_RIGHTS ----
sockaddr_un AbstractAddr(unsigned long long nonce, int rank, socklen_t *len) {
  sockaddr_un a;
  std::memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  char name[64];
  const int n = std::snprintf(name, sizeof(name), "qlb200-%016llx-%d", nonce, rank);
  std::memcpy(a.sun_path + 1, name, size_t(n));          // leading NUL: abstract namespace, nothing to unlink
  *len = socklen_t(offsetof(sockaddr_un, sun_path) + 1 + n);
  return a;
}

int SendFd(int sock, int tag, int fd) {
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  iovec io = {&tag, sizeof(tag)};
  msg.msg_iov = &io; msg.msg_iovlen = 1;
  char ctrl[CMSG_SPACE(sizeof(int))];
  std::memset(ctrl, 0, sizeof(ctrl));
  msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
  cmsghdr *c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
  std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
  return sendmsg(sock, &msg, 0) == ssize_t(sizeof(tag)) ? 0 : -1;
}

int RecvFd(int sock, int *tag, int *fd) {
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  iovec io = {tag, sizeof(*tag)};
  msg.msg_iov = &io; msg.msg_iovlen = 1;
  char ctrl[CMSG_SPACE(sizeof(int))];
  msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
  if (recvmsg(sock, &msg, MSG_WAITALL) != ssize_t(sizeof(*tag))) return -1;
  cmsghdr *c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return -1;
  std::memcpy(fd, CMSG_DATA(c), sizeof(int));
  return 0;
}

struct SymmBuf {
  size_t bytes = 0;                             // mapped size (granularity multiple)
  CUmemGenericAllocationHandle mine = 0;
  std::vector<CUmemGenericAllocationHandle> imported;   // peers' handles, [world] (mine at [rank])
  std::vector<CUdeviceptr> va;                  // [world] unicast mappings in this process
  CUmemGenericAllocationHandle mc = 0;
  CUdeviceptr mc_va = 0;
};

// barrier kernel: thread p publishes this rank's epoch into slot `rank` of peer p's pad, then waits for peer p's epoch in its
// own pad.  Pads live in a symmetric buffer; system-scope release / acquire orders the GEMM epilogue's peer stores before it.
// The epoch lives in device memory and is advanced by the kernel itself, so a captured CUDA graph replays correctly.
__global__ void CommBarrierKernel(unsigned long long *const *pads, int world, int rank, unsigned long long *epoch_ctr) {
  __shared__ unsigned long long s_epoch;
  const int p = threadIdx.x;
  if (p == 0) s_epoch = ++(*epoch_ctr);
  __syncthreads();
  if (p >= world) return;
  const unsigned long long epoch = s_epoch;
  __threadfence_system();
  unsigned long long *theirs = pads[p] + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
  const unsigned long long *mine = pads[rank] + p;
  unsigned long long v;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
  } while (v < epoch);
}

}  // namespace

struct qlb200_comm {
  qlb200_ctx *ctx = nullptr;
  int world = 1, rank = 0;
  qlb200_allgather_fn allgather = nullptr;
  void *user = nullptr;
  unsigned long long nonce = 0;
  int listen_fd = -1;
  bool multicast_ok = false;
  size_t gran = 0;
  std::vector<SymmBuf *> bufs;
  SymmBuf *pad = nullptr;                        // barrier signal pads
  unsigned long long **d_pads = nullptr;        // device array of the `world` pad pointers
  unsigned long long *d_epoch = nullptr;        // this rank's barrier count (device memory: graph-replay safe)
};

namespace {

int Gather(qlb200_comm *c, const void *send, void *recv, size_t bytes) {
  if (c->world == 1) { std::memcpy(recv, send, bytes); return QLB200_OK; }
  if (c->allgather(c->user, send, recv, bytes) != 0) return FailWith(QLB200_ERR_ARG, "the all-gather callback failed");
  return QLB200_OK;
}

// every rank sends `fd` (tagged with its rank) to every other rank; returns the received fds by rank (own entry = fd itself)
int ExchangeFds(qlb200_comm *c, int fd, std::vector<int> *out) {
  out->assign(c->world, -1);
  (*out)[c->rank] = fd;
  if (c->world == 1) return QLB200_OK;
  char token = 0, tokens[64];
  int rc = Gather(c, &token, tokens, 1);                  // everybody is listening
  if (rc != QLB200_OK) return rc;
  std::vector<int> socks;
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) continue;
    int s = socket(AF_UNIX, SOCK_STREAM, 0);
    socklen_t len;
    sockaddr_un a = AbstractAddr(c->nonce, p, &len);
    if (s < 0 || connect(s, reinterpret_cast<sockaddr *>(&a), len) != 0 || SendFd(s, c->rank, fd) != 0) {
      if (s >= 0) close(s);
      return FailWith(QLB200_ERR_CUDA, "could not pass a memory handle to a peer rank over the unix socket");
    }
    socks.push_back(s);
  }
  for (int i = 0; i + 1 < c->world; ++i) {
    int s = accept(c->listen_fd, nullptr, nullptr);
    int tag = -1, got = -1;
    if (s < 0 || RecvFd(s, &tag, &got) != 0 || tag < 0 || tag >= c->world) { if (s >= 0) close(s); return FailWith(QLB200_ERR_CUDA, "could not receive a peer's memory handle"); }
    (*out)[tag] = got;
    close(s);
  }
  for (int s : socks) close(s);
  return Gather(c, &token, tokens, 1);                    // nobody closes handles before everybody has imported
}

int AllocSymm(qlb200_comm *c, size_t bytes, bool want_mc, SymmBuf **out) {
  Driver &d = Drv();
  SymmBuf *b = new (std::nothrow) SymmBuf();
  if (!b) return FailWith(QLB200_ERR_NOMEM, "out of memory");
  QL_RT(cudaSetDevice(c->ctx->device));
  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = c->ctx->device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = c->gran;
  b->bytes = (bytes + gran - 1) / gran * gran;
  QL_DRVCALL(cuMemCreate(&b->mine, b->bytes, &prop, 0));
  int fd = -1;
  QL_DRVCALL(cuMemExportToShareableHandle(&fd, b->mine, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  std::vector<int> fds;
  int rc = ExchangeFds(c, fd, &fds);
  if (rc != QLB200_OK) return rc;
  b->imported.assign(c->world, 0);
  b->va.assign(c->world, 0);
  CUmemAccessDesc acc;
  std::memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = c->ctx->device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) b->imported[p] = b->mine;
    else QL_DRVCALL(cuMemImportFromShareableHandle(&b->imported[p], reinterpret_cast<void *>(static_cast<uintptr_t>(fds[p])), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    QL_DRVCALL(cuMemAddressReserve(&b->va[p], b->bytes, gran, 0, 0));
    QL_DRVCALL(cuMemMap(b->va[p], b->bytes, 0, b->imported[p], 0));
    QL_DRVCALL(cuMemSetAccess(b->va[p], b->bytes, &acc, 1));
    if (p != c->rank) close(fds[p]);
  }
  close(fd);
  if (want_mc && c->multicast_ok && c->world > 1) {
    // rank 0 creates the multicast object, everybody adds its device and binds its allocation, then maps the object
    CUmulticastObjectProp mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.numDevices = unsigned(c->world);
    mp.size = b->bytes;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    int mfd = -1;
    if (c->rank == 0) {
      QL_DRVCALL(cuMulticastCreate(&b->mc, &mp));
      QL_DRVCALL(cuMemExportToShareableHandle(&mfd, b->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    } else {
      mfd = open("/dev/null", 0);                         // placeholder so that every rank takes part in the exchange
    }
    std::vector<int> mfds;
    rc = ExchangeFds(c, mfd, &mfds);
    if (rc != QLB200_OK) return rc;
    if (c->rank != 0)
      QL_DRVCALL(cuMemImportFromShareableHandle(&b->mc, reinterpret_cast<void *>(static_cast<uintptr_t>(mfds[0])), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    for (int p = 0; p < c->world; ++p) if (mfds[p] >= 0) close(mfds[p]);
    CUdevice dev;
    QL_DRVCALL(cuDeviceGet(&dev, c->ctx->device));
    QL_DRVCALL(cuMulticastAddDevice(b->mc, dev));
    char token = 0, tokens[64];
    rc = Gather(c, &token, tokens, 1);                    // all devices added before anybody binds
    if (rc != QLB200_OK) return rc;
    QL_DRVCALL(cuMulticastBindMem(b->mc, 0, b->mine, 0, b->bytes, 0));
    QL_DRVCALL(cuMemAddressReserve(&b->mc_va, b->bytes, gran, 0, 0));
    QL_DRVCALL(cuMemMap(b->mc_va, b->bytes, 0, b->mc, 0));
    QL_DRVCALL(cuMemSetAccess(b->mc_va, b->bytes, &acc, 1));
    rc = Gather(c, &token, tokens, 1);
    if (rc != QLB200_OK) return rc;
  }
  QL_RT(cudaMemsetAsync(reinterpret_cast<void *>(b->va[c->rank]), 0, b->bytes, c->ctx->stream));
  QL_RT(cudaStreamSynchronize(c->ctx->stream));
  {
    char token = 0, tokens[64];
    rc = Gather(c, &token, tokens, 1);                    // zeroed everywhere before anybody stores into a peer
    if (rc != QLB200_OK) return rc;
  }
  *out = b;
  return QLB200_OK;
}

void FreeSymm(qlb200_comm *c, SymmBuf *b) {
  Driver &d = Drv();
  if (!b) return;
  if (b->mc_va) { d.p_cuMemUnmap(b->mc_va, b->bytes); d.p_cuMemAddressFree(b->mc_va, b->bytes); }
  if (b->mc) d.p_cuMemRelease(b->mc);
  for (int p = 0; p < int(b->va.size()); ++p) {
    if (b->va[p]) { d.p_cuMemUnmap(b->va[p], b->bytes); d.p_cuMemAddressFree(b->va[p], b->bytes); }
    if (b->imported[p] && p != c->rank) d.p_cuMemRelease(b->imported[p]);
  }
  if (b->mine) d.p_cuMemRelease(b->mine);
  delete b;
}

}  // namespace

extern "C" {

int qlb200_comm_create(qlb200_ctx *ctx, int32_t world, int32_t rank, qlb200_allgather_fn allgather, void *user, qlb200_comm **out) {
  if (!ctx || !out || world < 1 || world > 8 || rank < 0 || rank >= world) return FailWith(QLB200_ERR_ARG, "bad communicator arguments (1 <= world <= 8)");
  if (world > 1 && !allgather) return FailWith(QLB200_ERR_ARG, "an all-gather callback is needed for world > 1");
  Driver &d = Drv();
  if (!d.ok) return FailWith(QLB200_ERR_UNSUPPORTED, d.why);
  QL_RT(cudaSetDevice(ctx->device));
  QL_RT(cudaFree(nullptr));
  qlb200_comm *c = new (std::nothrow) qlb200_comm();
  if (!c) return FailWith(QLB200_ERR_NOMEM, "out of memory");
  c->ctx = ctx; c->world = world; c->rank = rank; c->allgather = allgather; c->user = user;
  CUdevice dev;
  QL_DRVCALL(cuDeviceGet(&dev, ctx->device));
  int vmm = 0, posix = 0, mcast = 0;
  QL_DRVCALL(cuDeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev));
  QL_DRVCALL(cuDeviceGetAttribute(&posix, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev));
  QL_DRVCALL(cuDeviceGetAttribute(&mcast, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev));
  if (!vmm || !posix) { delete c; return FailWith(QLB200_ERR_UNSUPPORTED, "device lacks virtual memory management / POSIX handle export"); }
  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = ctx->device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t g1 = 0, g2 = 0;
  QL_DRVCALL(cuMemGetAllocationGranularity(&g1, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (mcast && world > 1) {
    CUmulticastObjectProp mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.numDevices = unsigned(world); mp.size = g1; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (d.p_cuMulticastGetGranularity(&g2, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) { mcast = 0; g2 = 0; }
  }
  c->gran = g1 > g2 ? g1 : g2;
  // every rank must agree on multicast support and on the nonce that names the sockets
  struct Boot { unsigned long long nonce; int mcast; int pad; } mine = {0, mcast, 0}, all[8];
  if (rank == 0) {
    FILE *f = std::fopen("/dev/urandom", "rb");
    if (!f || std::fread(&mine.nonce, sizeof(mine.nonce), 1, f) != 1) mine.nonce = (unsigned long long) getpid() * 2654435761ull + (unsigned long long) (uintptr_t) c;
    if (f) std::fclose(f);
  }
  int rc = Gather(c, &mine, all, sizeof(Boot));
  if (rc != QLB200_OK) { delete c; return rc; }
  c->nonce = all[0].nonce;
  c->multicast_ok = world > 1;
  for (int p = 0; p < world; ++p) c->multicast_ok = c->multicast_ok && all[p].mcast != 0;
  if (world > 1) {
    c->listen_fd = socket(AF_UNIX, SOCK_STREAM, 0);
    socklen_t len;
    sockaddr_un a = AbstractAddr(c->nonce, rank, &len);
    if (c->listen_fd < 0 || bind(c->listen_fd, reinterpret_cast<sockaddr *>(&a), len) != 0 || listen(c->listen_fd, 16) != 0) {
      if (c->listen_fd >= 0) close(c->listen_fd);
      delete c;
      return FailWith(QLB200_ERR_CUDA, "could not open the bootstrap unix socket");
    }
  }
  // barrier signal pads: one symmetric buffer, a device table of the world pad pointers
  rc = AllocSymm(c, 4096, false, &c->pad);
  if (rc != QLB200_OK) { qlb200_comm_destroy(c); return rc; }
  std::vector<unsigned long long *> pads(world);
  for (int p = 0; p < world; ++p) pads[p] = reinterpret_cast<unsigned long long *>(c->pad->va[p]);
  QL_RT(cudaMalloc(reinterpret_cast<void **>(&c->d_pads), world * sizeof(void *)));
  QL_RT(cudaMemcpy(c->d_pads, pads.data(), world * sizeof(void *), cudaMemcpyHostToDevice));
  QL_RT(cudaMalloc(reinterpret_cast<void **>(&c->d_epoch), sizeof(unsigned long long)));
  QL_RT(cudaMemset(c->d_epoch, 0, sizeof(unsigned long long)));
  *out = c;
  return QLB200_OK;
}

void qlb200_comm_destroy(qlb200_comm *c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  if (c->world > 1 && c->allgather) { char t = 0, ts[64]; c->allgather(c->user, &t, ts, 1); }   // nobody unmaps while a peer may still store
  for (SymmBuf *b : c->bufs) FreeSymm(c, b);
  FreeSymm(c, c->pad);
  if (c->d_pads) cudaFree(c->d_pads);
  if (c->d_epoch) cudaFree(c->d_epoch);
  if (c->listen_fd >= 0) close(c->listen_fd);
  delete c;
}

int qlb200_comm_has_multicast(const qlb200_comm *c) { return c && c->multicast_ok ? 1 : 0; }

int qlb200_comm_alloc(qlb200_comm *c, size_t bytes, void **local, void **peers, void **multicast) {
  if (!c || !local || bytes == 0) return FailWith(QLB200_ERR_ARG, "bad argument");
  SymmBuf *b = nullptr;
  int rc = AllocSymm(c, bytes, multicast != nullptr, &b);
  if (rc != QLB200_OK) return rc;
  c->bufs.push_back(b);
  *local = reinterpret_cast<void *>(b->va[c->rank]);
  if (peers) for (int p = 0; p < c->world; ++p) peers[p] = reinterpret_cast<void *>(b->va[p]);
  if (multicast) *multicast = reinterpret_cast<void *>(b->mc_va);
  return QLB200_OK;
}

int qlb200_comm_free(qlb200_comm *c, void *local) {
  if (!c || !local) return FailWith(QLB200_ERR_ARG, "null argument");
  for (size_t i = 0; i < c->bufs.size(); ++i)
    if (reinterpret_cast<void *>(c->bufs[i]->va[c->rank]) == local) {
      cudaSetDevice(c->ctx->device);
      cudaStreamSynchronize(c->ctx->stream);
      if (c->world > 1) { char t = 0, ts[64]; int rc = Gather(c, &t, ts, 1); if (rc != QLB200_OK) return rc; }
      FreeSymm(c, c->bufs[i]);
      c->bufs.erase(c->bufs.begin() + i);
      return QLB200_OK;
    }
  return FailWith(QLB200_ERR_ARG, "not a buffer of this communicator");
}

int qlb200_comm_barrier(qlb200_comm *c) {
  if (!c) return FailWith(QLB200_ERR_ARG, "null argument");
  QL_RT(cudaSetDevice(c->ctx->device));
  CommBarrierKernel<<<1, 32, 0, c->ctx->stream>>>(c->d_pads, c->world, c->rank, c->d_epoch);
  QL_RT(cudaGetLastError());
  c->ctx->launches = 1; ++c->ctx->total_launches;
  return QLB200_OK;
}

}  // extern "C"
