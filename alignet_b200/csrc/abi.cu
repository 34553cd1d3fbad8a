// abi.cu -- utility entry points of the C ABI and the process-wide bookkeeping.
#include "volstn_common.cuh"

namespace volstn {

std::atomic<uint64_t> g_launches{0};

static thread_local cudaError_t t_last_error = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { t_last_error = e; }

}  // namespace volstn

extern "C" {

int volstn_b200_abi_version(void) { return VOLSTN_B200_ABI_VERSION; }

const char *volstn_b200_strerror(int status) {
  switch (status) {
    case VOLSTN_OK: return "ok";
    case VOLSTN_ERR_ARG: return "invalid argument (null pointer, bad dtype/G/flags, negative size)";
    case VOLSTN_ERR_SHAPE: return "tensor shapes violate the module contract (BilinearSamplerVolumetric.lua:26-42)";
    case VOLSTN_ERR_LAYOUT: return "unsupported layout (a tensor this entry point needs contiguous, dense or non-aliasing is not)";
    case VOLSTN_ERR_CUDA: return "CUDA call failed (see volstn_b200_last_cuda_error_string)";
    case VOLSTN_ERR_UNSUPPORTED: return "unsupported problem size";
    case VOLSTN_ERR_NO_DEVICE: return "no sm_100 (B200) device: this library has no fallback path";
    case VOLSTN_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

int volstn_b200_last_cuda_error(void) { return (int)volstn::t_last_error; }

const char *volstn_b200_last_cuda_error_string(void) { return cudaGetErrorString(volstn::t_last_error); }

int volstn_b200_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    volstn::set_last_cuda_error(e);
    (void)cudaGetLastError();
    return VOLSTN_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) return VOLSTN_ERR_NO_DEVICE;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) return volstn::cuda_fail(e);
  return major == 10 ? VOLSTN_OK : VOLSTN_ERR_NO_DEVICE;
}

uint64_t volstn_b200_launch_count(void) { return volstn::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
