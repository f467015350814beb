// topo.cpp — GPU / NVLink / NVSwitch topology discovery for dist_init().
// CUDA runtime for devices and the P2P matrix, the driver (via
// cudaGetDriverEntryPoint) for multicast support, NVML (dlopen'ed, optional) for
// NVLink link state and version.  Returns JSON; never throws on a GPU-less box.
#include "comm.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <sstream>

namespace cpd {
namespace comm {

namespace {
struct Nvml {
  void* lib = nullptr;
  int (*init)() = nullptr;
  int (*shutdown)() = nullptr;
  int (*handle_by_index)(unsigned, void**) = nullptr;
  int (*link_state)(void*, unsigned, unsigned*) = nullptr;
  int (*link_version)(void*, unsigned, unsigned*) = nullptr;
  int (*driver_version)(char*, unsigned) = nullptr;
  bool ok = false;
  Nvml() {
    lib = dlopen("libnvidia-ml.so.1", RTLD_LAZY | RTLD_LOCAL);
    if (!lib) return;
    init = reinterpret_cast<int (*)()>(dlsym(lib, "nvmlInit_v2"));
    shutdown = reinterpret_cast<int (*)()>(dlsym(lib, "nvmlShutdown"));
    handle_by_index = reinterpret_cast<int (*)(unsigned, void**)>(dlsym(lib, "nvmlDeviceGetHandleByIndex_v2"));
    link_state = reinterpret_cast<int (*)(void*, unsigned, unsigned*)>(dlsym(lib, "nvmlDeviceGetNvLinkState"));
    link_version = reinterpret_cast<int (*)(void*, unsigned, unsigned*)>(dlsym(lib, "nvmlDeviceGetNvLinkVersion"));
    driver_version = reinterpret_cast<int (*)(char*, unsigned)>(dlsym(lib, "nvmlSystemGetDriverVersion"));
    ok = init && handle_by_index && link_state && init() == 0;
  }
  ~Nvml() {
    if (ok && shutdown) shutdown();
    if (lib) dlclose(lib);
  }
};
}  // namespace

std::string discover_topology_json() {
  std::ostringstream os;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
  os << "{\"device_count\": " << n;
  if (n == 0) { os << ", \"devices\": [], \"p2p\": [], \"nvlink\": [], \"multicast\": false}"; return os.str(); }

  int rt = 0, drvv = 0;
  cudaRuntimeGetVersion(&rt);
  cudaDriverGetVersion(&drvv);
  os << ", \"cuda_runtime\": " << rt << ", \"cuda_driver\": " << drvv;

  using PFN_cuDeviceGet = CUresult (*)(CUdevice*, int);
  using PFN_cuDeviceGetAttribute = CUresult (*)(int*, CUdevice_attribute, CUdevice);
  PFN_cuDeviceGet cuDeviceGet_ = nullptr;
  PFN_cuDeviceGetAttribute cuDeviceGetAttribute_ = nullptr;
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuDeviceGet", &fn, cudaEnableDefault, &q) == cudaSuccess) cuDeviceGet_ = reinterpret_cast<PFN_cuDeviceGet>(fn);
    if (cudaGetDriverEntryPoint("cuDeviceGetAttribute", &fn, cudaEnableDefault, &q) == cudaSuccess) cuDeviceGetAttribute_ = reinterpret_cast<PFN_cuDeviceGetAttribute>(fn);
    cudaGetLastError();
  }

  bool all_mc = true;
  os << ", \"devices\": [";
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, d);
    int mc = 0;
    if (cuDeviceGet_ && cuDeviceGetAttribute_) {
      CUdevice dev;
      if (cuDeviceGet_(&dev, d) == CUDA_SUCCESS) cuDeviceGetAttribute_(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
    }
    all_mc = all_mc && mc;
    os << (d ? ", " : "") << "{\"index\": " << d << ", \"name\": \"" << p.name << "\", \"cc\": \"" << p.major << "." << p.minor
       << "\", \"sms\": " << p.multiProcessorCount << ", \"mem_bytes\": " << (unsigned long long)p.totalGlobalMem
       << ", \"l2_bytes\": " << p.l2CacheSize << ", \"smem_per_block_optin\": " << p.sharedMemPerBlockOptin
       << ", \"smem_per_sm\": " << p.sharedMemPerMultiprocessor << ", \"multicast\": " << (mc ? "true" : "false") << "}";
  }
  os << "], \"p2p\": [";
  for (int i = 0; i < n; ++i) {
    os << (i ? ", " : "") << "[";
    for (int j = 0; j < n; ++j) {
      int access = (i == j), rank = 0, atomics = (i == j);
      if (i != j) {
        cudaDeviceCanAccessPeer(&access, i, j);
        cudaDeviceGetP2PAttribute(&rank, cudaDevP2PAttrPerformanceRank, i, j);
        cudaDeviceGetP2PAttribute(&atomics, cudaDevP2PAttrNativeAtomicSupported, i, j);
      }
      os << (j ? ", " : "") << "{\"access\": " << access << ", \"perf_rank\": " << rank << ", \"atomics\": " << atomics << "}";
    }
    os << "]";
  }
  cudaGetLastError();
  os << "], \"nvlink\": [";
  {
    Nvml nv;
    for (int d = 0; d < n; ++d) {
      int active = 0;
      unsigned version = 0;
      if (nv.ok) {
        void* h = nullptr;
        if (nv.handle_by_index((unsigned)d, &h) == 0) {
          for (unsigned l = 0; l < 32; ++l) {
            unsigned st = 0;
            if (nv.link_state(h, l, &st) != 0) break;
            if (st) {
              ++active;
              if (nv.link_version) nv.link_version(h, l, &version);
            }
          }
        }
      }
      os << (d ? ", " : "") << "{\"active_links\": " << active << ", \"version\": " << version << "}";
    }
    os << "], \"nvml\": " << (nv.ok ? "true" : "false");
  }
  os << ", \"multicast\": " << (all_mc ? "true" : "false") << "}";
  return os.str();
}

}  // namespace comm
}  // namespace cpd
