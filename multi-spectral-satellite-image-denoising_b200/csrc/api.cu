// Library-level entry points: version, thread-local error text, device check.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";

void msd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// programmatic dependent launch of the tensor-core kernel chain (common.cuh); read on every launch so that A/B runs can switch it
bool msd_pdl_enabled() {
  const char* e = getenv("MSD_NO_PDL");
  return !(e != nullptr && e[0] == '1');
}

extern "C" int msd_version(void) { return 2; }

extern "C" const char* msd_last_error(void) { return g_err; }

extern "C" int msd_check_device(void) {
  int dev = 0;
  MSD_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  MSD_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MSD_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    msd_set_error("libmsd_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return MSD_ERR_ARCH;
  }
  return MSD_OK;
}
