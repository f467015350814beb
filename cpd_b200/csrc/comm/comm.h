// comm.h — native communication runtime: GPU/NVLink topology discovery and the
// symmetric (peer-mapped, optionally NVLS-multicast) memory regions that the fused
// all-reduce kernels operate on.  The reference has nothing comparable: its
// dist_init (CPDtorch/utils/dist_util.py:96-131) only parses SLURM variables and
// calls init_process_group('nccl').
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace cpd {
namespace comm {

// JSON description of the box: per-device name / SMs / memory / L2 / smem, the P2P
// access + performance-rank + native-atomics matrix, NVLink link counts and
// versions (NVML, dlopen'ed), multicast (NVLS) support.
std::string discover_topology_json();

// One symmetric allocation: `bytes` on every rank, every rank's copy mapped into
// every other rank's address space; optionally one multicast mapping that aliases
// all copies (stores are replicated by the NVSwitch).
//
// Collective bring-up, with a caller-provided barrier between the steps:
//   open()            local allocation (cuMemCreate, shareable POSIX fd), listen socket;
//                     rank 0 also creates the multicast object
//   -- barrier --
//   exchange()        send/receive the fds over abstract unix sockets (SCM_RIGHTS),
//                     import + map every peer, add this device to the multicast object
//   -- barrier --
//   bind_multicast()  bind the local allocation into the multicast object and map it
//   -- barrier --
class SymmRegion {
 public:
  SymmRegion(size_t bytes, int rank, int world, int device, const std::string& session, bool want_multicast);
  ~SymmRegion();
  void open();
  void exchange();
  void bind_multicast();
  void close_region();      // collective teardown: unmap / unbind / release everything (caller barriers before and after)
  void* ptr(int r) const { return ptrs_[r]; }
  void* mc_ptr() const { return mc_ptr_; }
  size_t bytes() const { return bytes_; }
  bool multicast() const { return mc_ptr_ != nullptr; }
  const std::string& log() const { return log_; }

 private:
  size_t bytes_;
  int rank_, world_, device_;
  std::string session_;
  bool want_mc_;
  std::vector<void*> ptrs_;
  std::vector<unsigned long long> handles_;
  void* mc_ptr_ = nullptr;
  unsigned long long mc_handle_ = 0;
  int local_fd_ = -1, mc_fd_ = -1, listen_fd_ = -1;
  bool mc_ok_ = false, closed_ = false;
  std::string log_;
};

// Legacy CUDA-IPC variant (no multicast): plain cudaMalloc + cudaIpc handles.
struct IpcRegion {
  static void* alloc(size_t bytes);                       // cudaMalloc + memset 0
  static std::string export_handle(void* p);              // 64 opaque bytes
  static void* open_handle(const std::string& h);         // peer mapping
  static void close_handle(void* p);
  static void free_local(void* p);
};

}  // namespace comm
}  // namespace cpd
