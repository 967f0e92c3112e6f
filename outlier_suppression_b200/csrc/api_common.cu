// api_common.cu -- error reporting, version, device queries.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace osq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace osq

extern "C" {

int osq_version(void) { return OSQ_VERSION; }

const char* osq_last_error(void) { return osq::g_err; }

int osq_sm_count(void) {
  int n = osq::sm_count();
  if (n < 0) {
    osq::set_error("no CUDA device");
    return OSQ_ECUDA;
  }
  return n;
}

int64_t osq_workspace_bytes(void) { return osq::kWorkspaceBytes; }

}  // extern "C"
