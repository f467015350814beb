// symm_mem.cpp — symmetric peer-mapped memory over the CUDA VMM API, with an
// NVLS multicast view, for the fused all-reduce kernels.  Driver entry points are
// resolved through cudaGetDriverEntryPoint so the library loads on machines
// without libcuda (CPU-only CI) and never links the driver stub.
#include "comm.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>
#include <poll.h>

#include <cerrno>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace cpd {
namespace comm {

namespace {

template <class Fn>
Fn drv(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || fn == nullptr) {
    cudaGetLastError();
    throw std::runtime_error(std::string("cpd_b200: driver entry point not available: ") + name);
  }
  return reinterpret_cast<Fn>(fn);
}

void ck(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) {
    std::ostringstream os;
    os << "cpd_b200 symm_mem: " << what << " failed with CUresult " << (int)r;
    throw std::runtime_error(os.str());
  }
}
void ckrt(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    std::ostringstream os;
    os << "cpd_b200 symm_mem: " << what << ": " << cudaGetErrorString(e);
    cudaGetLastError();
    throw std::runtime_error(os.str());
  }
}

using PFN_cuMemCreate = CUresult (*)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
using PFN_cuMemRelease = CUresult (*)(CUmemGenericAllocationHandle);
using PFN_cuMemExport = CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
using PFN_cuMemImport = CUresult (*)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
using PFN_cuMemAddressReserve = CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
using PFN_cuMemAddressFree = CUresult (*)(CUdeviceptr, size_t);
using PFN_cuMemMap = CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
using PFN_cuMemUnmap = CUresult (*)(CUdeviceptr, size_t);
using PFN_cuMemSetAccess = CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
using PFN_cuMemGetGranularity = CUresult (*)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
using PFN_cuMulticastCreate = CUresult (*)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
using PFN_cuMulticastAddDevice = CUresult (*)(CUmemGenericAllocationHandle, CUdevice);
using PFN_cuMulticastBindMem = CUresult (*)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
using PFN_cuMulticastUnbind = CUresult (*)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
using PFN_cuMulticastGetGranularity = CUresult (*)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
using PFN_cuDeviceGet = CUresult (*)(CUdevice*, int);
using PFN_cuDeviceGetAttribute = CUresult (*)(int*, CUdevice_attribute, CUdevice);

std::string sock_name(const std::string& session, int rank) {
  std::ostringstream os;
  os << "cpd_b200." << session << "." << rank;
  return os.str();
}

sockaddr_un abstract_addr(const std::string& name, socklen_t* len) {
  sockaddr_un a;
  std::memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  // abstract namespace: leading NUL, no filesystem entry to clean up
  const size_t n = std::min(name.size(), sizeof(a.sun_path) - 2);
  std::memcpy(a.sun_path + 1, name.data(), n);
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
  return a;
}

void send_fd(int sock, int fd, int sender_rank, int kind) {
  int payload[2] = {sender_rank, kind};
  iovec iov{payload, sizeof(payload)};
  char ctrl[CMSG_SPACE(sizeof(int))];
  std::memset(ctrl, 0, sizeof(ctrl));
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
  if (sendmsg(sock, &msg, 0) < 0) throw std::runtime_error(std::string("cpd_b200 symm_mem: sendmsg: ") + strerror(errno));
}

bool recv_fd(int sock, int* fd, int* sender_rank, int* kind) {
  int payload[2] = {-1, -1};
  iovec iov{payload, sizeof(payload)};
  char ctrl[CMSG_SPACE(sizeof(int))];
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  ssize_t n = recvmsg(sock, &msg, 0);
  if (n <= 0) return false;
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_type != SCM_RIGHTS) return false;
  std::memcpy(fd, CMSG_DATA(c), sizeof(int));
  *sender_rank = payload[0];
  *kind = payload[1];
  return true;
}

}  // namespace

SymmRegion::SymmRegion(size_t bytes, int rank, int world, int device, const std::string& session, bool want_multicast)
    : bytes_(bytes), rank_(rank), world_(world), device_(device), session_(session), want_mc_(want_multicast),
      ptrs_(world, nullptr), handles_(world, 0) {}

void SymmRegion::open() {
  ckrt(cudaSetDevice(device_), "cudaSetDevice");
  ckrt(cudaFree(0), "context init");
  auto cuMemCreate_ = drv<PFN_cuMemCreate>("cuMemCreate");
  auto cuMemExport_ = drv<PFN_cuMemExport>("cuMemExportToShareableHandle");
  auto cuMemGetGran_ = drv<PFN_cuMemGetGranularity>("cuMemGetAllocationGranularity");
  auto cuMemAddressReserve_ = drv<PFN_cuMemAddressReserve>("cuMemAddressReserve");
  auto cuMemMap_ = drv<PFN_cuMemMap>("cuMemMap");
  auto cuMemSetAccess_ = drv<PFN_cuMemSetAccess>("cuMemSetAccess");
  auto cuDeviceGet_ = drv<PFN_cuDeviceGet>("cuDeviceGet");
  auto cuDeviceGetAttribute_ = drv<PFN_cuDeviceGetAttribute>("cuDeviceGetAttribute");

  CUdevice dev;
  ck(cuDeviceGet_(&dev, device_), "cuDeviceGet");
  int mc_supported = 0;
  cuDeviceGetAttribute_(&mc_supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
  mc_ok_ = want_mc_ && mc_supported && world_ > 1;

  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device_;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  ck(cuMemGetGran_(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
  if (mc_ok_) {
    auto cuMulticastGetGran_ = drv<PFN_cuMulticastGetGranularity>("cuMulticastGetGranularity");
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)world_;
    mp.size = bytes_;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (cuMulticastGetGran_(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > gran) gran = mg;
  }
  bytes_ = (bytes_ + gran - 1) / gran * gran;

  CUmemGenericAllocationHandle h;
  ck(cuMemCreate_(&h, bytes_, &prop, 0), "cuMemCreate");
  handles_[rank_] = h;
  ck(cuMemExport_(&local_fd_, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");

  CUdeviceptr va = 0;
  ck(cuMemAddressReserve_(&va, bytes_, gran, 0, 0), "cuMemAddressReserve");
  ck(cuMemMap_(va, bytes_, 0, h, 0), "cuMemMap(local)");
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  ck(cuMemSetAccess_(va, bytes_, &acc, 1), "cuMemSetAccess(local)");
  ptrs_[rank_] = reinterpret_cast<void*>(va);
  ckrt(cudaMemset(ptrs_[rank_], 0, bytes_), "cudaMemset(symmetric region)");
  ckrt(cudaDeviceSynchronize(), "cudaDeviceSynchronize");

  if (mc_ok_ && rank_ == 0) {
    auto cuMulticastCreate_ = drv<PFN_cuMulticastCreate>("cuMulticastCreate");
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)world_;
    mp.size = bytes_;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle mh;
    CUresult r = cuMulticastCreate_(&mh, &mp);
    if (r == CUDA_SUCCESS && cuMemExport_(&mc_fd_, mh, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS) {
      mc_handle_ = mh;
    } else {
      log_ += "multicast create/export failed (CUresult " + std::to_string((int)r) + "); ";
      mc_fd_ = -1;
    }
  }

  if (world_ > 1) {
    listen_fd_ = socket(AF_UNIX, SOCK_STREAM, 0);
    if (listen_fd_ < 0) throw std::runtime_error("cpd_b200 symm_mem: socket() failed");
    socklen_t len;
    sockaddr_un a = abstract_addr(sock_name(session_, rank_), &len);
    if (bind(listen_fd_, reinterpret_cast<sockaddr*>(&a), len) < 0)
      throw std::runtime_error(std::string("cpd_b200 symm_mem: bind: ") + strerror(errno));
    if (listen(listen_fd_, 64) < 0) throw std::runtime_error("cpd_b200 symm_mem: listen failed");
  }
}

void SymmRegion::exchange() {
  if (world_ == 1) return;
  auto cuMemImport_ = drv<PFN_cuMemImport>("cuMemImportFromShareableHandle");
  auto cuMemAddressReserve_ = drv<PFN_cuMemAddressReserve>("cuMemAddressReserve");
  auto cuMemMap_ = drv<PFN_cuMemMap>("cuMemMap");
  auto cuMemSetAccess_ = drv<PFN_cuMemSetAccess>("cuMemSetAccess");
  ckrt(cudaSetDevice(device_), "cudaSetDevice");

  // 1. push our fd (and rank 0's multicast fd, or kind 2 = "no multicast") to every peer
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    int s = socket(AF_UNIX, SOCK_STREAM, 0);
    socklen_t len;
    sockaddr_un a = abstract_addr(sock_name(session_, r), &len);
    int tries = 0;
    while (connect(s, reinterpret_cast<sockaddr*>(&a), len) < 0) {
      if (++tries > 2000) throw std::runtime_error(std::string("cpd_b200 symm_mem: connect: ") + strerror(errno));
      usleep(5000);
    }
    send_fd(s, local_fd_, rank_, 0);
    if (rank_ == 0 && mc_ok_) send_fd(s, mc_fd_ >= 0 ? mc_fd_ : local_fd_, 0, mc_fd_ >= 0 ? 1 : 2);
    close(s);
  }
  // 2. collect W-1 memory fds (+ the multicast message from rank 0)
  int expect = (world_ - 1) + ((mc_ok_ && rank_ != 0) ? 1 : 0);
  std::vector<int> peer_fd(world_, -1);
  int mc_in = -1;
  bool mc_denied = false;
  while (expect > 0) {
    pollfd pf{listen_fd_, POLLIN, 0};
    if (poll(&pf, 1, 60000) <= 0) throw std::runtime_error("cpd_b200 symm_mem: timed out waiting for peer handles");
    int c = accept(listen_fd_, nullptr, nullptr);
    if (c < 0) continue;
    int fd, sr, kind;
    while (expect > 0 && recv_fd(c, &fd, &sr, &kind)) {
      if (kind == 0 && sr >= 0 && sr < world_) peer_fd[sr] = fd;
      else if (kind == 1) mc_in = fd;
      else { mc_denied = true; close(fd); }
      --expect;
    }
    close(c);
  }
  close(listen_fd_);
  listen_fd_ = -1;

  // 3. import + map every peer allocation
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    if (peer_fd[r] < 0) throw std::runtime_error("cpd_b200 symm_mem: missing handle from rank " + std::to_string(r));
    CUmemGenericAllocationHandle h;
    ck(cuMemImport_(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(peer_fd[r])), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
       "cuMemImportFromShareableHandle(peer)");
    close(peer_fd[r]);
    handles_[r] = h;
    CUdeviceptr va = 0;
    ck(cuMemAddressReserve_(&va, bytes_, 0, 0, 0), "cuMemAddressReserve(peer)");
    ck(cuMemMap_(va, bytes_, 0, h, 0), "cuMemMap(peer)");
    ck(cuMemSetAccess_(va, bytes_, &acc, 1), "cuMemSetAccess(peer)");
    ptrs_[r] = reinterpret_cast<void*>(va);
  }

  // 4. multicast: import (non-zero ranks) and add this device
  if (mc_ok_) {
    if (rank_ != 0) {
      if (mc_in >= 0 && !mc_denied) {
        CUmemGenericAllocationHandle mh;
        CUresult r = cuMemImport_(&mh, reinterpret_cast<void*>(static_cast<uintptr_t>(mc_in)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        close(mc_in);
        if (r == CUDA_SUCCESS) mc_handle_ = mh;
        else log_ += "multicast import failed; ";
      }
    }
    if (mc_handle_) {
      auto cuMulticastAddDevice_ = drv<PFN_cuMulticastAddDevice>("cuMulticastAddDevice");
      auto cuDeviceGet_ = drv<PFN_cuDeviceGet>("cuDeviceGet");
      CUdevice dev;
      ck(cuDeviceGet_(&dev, device_), "cuDeviceGet");
      CUresult r = cuMulticastAddDevice_(mc_handle_, dev);
      if (r != CUDA_SUCCESS) { log_ += "cuMulticastAddDevice failed (" + std::to_string((int)r) + "); "; mc_handle_ = 0; }
    }
  }
}

void SymmRegion::bind_multicast() {
  if (!mc_ok_ || !mc_handle_) return;
  auto cuMulticastBindMem_ = drv<PFN_cuMulticastBindMem>("cuMulticastBindMem");
  auto cuMemAddressReserve_ = drv<PFN_cuMemAddressReserve>("cuMemAddressReserve");
  auto cuMemMap_ = drv<PFN_cuMemMap>("cuMemMap");
  auto cuMemSetAccess_ = drv<PFN_cuMemSetAccess>("cuMemSetAccess");
  ckrt(cudaSetDevice(device_), "cudaSetDevice");
  CUresult r = cuMulticastBindMem_(mc_handle_, 0, handles_[rank_], 0, bytes_, 0);
  if (r != CUDA_SUCCESS) { log_ += "cuMulticastBindMem failed (" + std::to_string((int)r) + "); "; return; }
  CUdeviceptr va = 0;
  if (cuMemAddressReserve_(&va, bytes_, 0, 0, 0) != CUDA_SUCCESS) { log_ += "mc reserve failed; "; return; }
  r = cuMemMap_(va, bytes_, 0, mc_handle_, 0);
  if (r != CUDA_SUCCESS) { log_ += "mc map failed (" + std::to_string((int)r) + "); "; return; }
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = cuMemSetAccess_(va, bytes_, &acc, 1);
  if (r != CUDA_SUCCESS) { log_ += "mc set-access failed (" + std::to_string((int)r) + "); "; return; }
  mc_ptr_ = reinterpret_cast<void*>(va);
}

// Collective teardown (the caller synchronises the device and barriers the group first: no kernel of any rank may
// still touch a mapping): multicast view -> unbind -> multicast object, then every peer mapping and the local
// allocation.  Errors are logged, not thrown: a half-released region must not take the training job down.
void SymmRegion::close_region() {
  if (closed_) return;
  closed_ = true;
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  PFN_cuMemUnmap cuMemUnmap_ = nullptr;
  PFN_cuMemAddressFree cuMemAddressFree_ = nullptr;
  PFN_cuMemRelease cuMemRelease_ = nullptr;
  try {
    cuMemUnmap_ = drv<PFN_cuMemUnmap>("cuMemUnmap");
    cuMemAddressFree_ = drv<PFN_cuMemAddressFree>("cuMemAddressFree");
    cuMemRelease_ = drv<PFN_cuMemRelease>("cuMemRelease");
  } catch (const std::exception& e) { log_ += std::string("close: ") + e.what() + "; "; return; }
  auto note = [&](CUresult r, const char* what) { if (r != CUDA_SUCCESS) log_ += std::string(what) + " failed (" + std::to_string((int)r) + "); "; };
  if (mc_ptr_) {
    const CUdeviceptr va = reinterpret_cast<CUdeviceptr>(mc_ptr_);
    note(cuMemUnmap_(va, bytes_), "cuMemUnmap(mc)");
    note(cuMemAddressFree_(va, bytes_), "cuMemAddressFree(mc)");
    try {
      auto cuMulticastUnbind_ = drv<PFN_cuMulticastUnbind>("cuMulticastUnbind");
      auto cuDeviceGet_ = drv<PFN_cuDeviceGet>("cuDeviceGet");
      CUdevice dev;
      if (cuDeviceGet_(&dev, device_) == CUDA_SUCCESS) note(cuMulticastUnbind_(mc_handle_, dev, 0, bytes_), "cuMulticastUnbind");
    } catch (const std::exception& e) { log_ += std::string("close: ") + e.what() + "; "; }
    mc_ptr_ = nullptr;
  }
  if (mc_handle_) { note(cuMemRelease_(mc_handle_), "cuMemRelease(mc)"); mc_handle_ = 0; }
  for (int r = 0; r < world_; ++r) {
    if (ptrs_[r]) {
      const CUdeviceptr va = reinterpret_cast<CUdeviceptr>(ptrs_[r]);
      note(cuMemUnmap_(va, bytes_), "cuMemUnmap");
      note(cuMemAddressFree_(va, bytes_), "cuMemAddressFree");
      ptrs_[r] = nullptr;
    }
    if (handles_[r]) { note(cuMemRelease_(handles_[r]), "cuMemRelease"); handles_[r] = 0; }
  }
}

SymmRegion::~SymmRegion() {
  // without an explicit (collective) close_region() the mappings live for the life of the process
  if (listen_fd_ >= 0) close(listen_fd_);
  if (local_fd_ >= 0) close(local_fd_);
  if (mc_fd_ >= 0) close(mc_fd_);
}

// ---- legacy CUDA IPC ---------------------------------------------------------------
void* IpcRegion::alloc(size_t bytes) {
  void* p = nullptr;
  ckrt(cudaMalloc(&p, bytes), "cudaMalloc");
  ckrt(cudaMemset(p, 0, bytes), "cudaMemset");
  ckrt(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  return p;
}
std::string IpcRegion::export_handle(void* p) {
  cudaIpcMemHandle_t h;
  ckrt(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
  return std::string(reinterpret_cast<const char*>(&h), sizeof(h));
}
void* IpcRegion::open_handle(const std::string& s) {
  if (s.size() != sizeof(cudaIpcMemHandle_t)) throw std::runtime_error("cpd_b200: bad IPC handle size");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, s.data(), sizeof(h));
  void* p = nullptr;
  ckrt(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
  return p;
}
void IpcRegion::close_handle(void* p) { cudaIpcCloseMemHandle(p); }
void IpcRegion::free_local(void* p) { cudaFree(p); }

}  // namespace comm
}  // namespace cpd
