// api.cu -- version / error reporting for the C ABI (include/r4r_b200.h).
#include "common.cuh"

static thread_local char g_err[512] = "";

void r4r_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int r4r_abi_version(void) { return R4R_ABI_VERSION; }
extern "C" const char* r4r_last_error(void) { return g_err; }

extern "C" int r4r_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* smem_optin_bytes) {
  int dev = 0;
  R4R_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  R4R_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (smem_optin_bytes) *smem_optin_bytes = (int64_t)p.sharedMemPerBlockOptin;
  return 0;
}
