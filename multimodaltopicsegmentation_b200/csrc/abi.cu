// C-ABI housekeeping for libmts_b200.so: version, last-error string, device check.
#include <string.h>

#include "common.cuh"

namespace mts {
static thread_local char g_err[256] = "";
void set_error(const char *msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}
}  // namespace mts

extern "C" int mts_version(void) { return 1; }
extern "C" const char *mts_last_error(void) { return mts::g_err; }
extern "C" int mts_device_ok(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    mts::set_error("no CUDA device");
    return MTS_E_NODEVICE;
  }
  if (p.major != 10) {
    mts::set_error("libmts_b200 needs a compute-capability 10.x (Blackwell) device");
    return MTS_E_NODEVICE;
  }
  return 0;
}
