// Library-level entry points: version, last-error string, device probe,
// launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace cslam {

static thread_local char t_err[1024] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

}  // namespace cslam

extern "C" {

const char* cslam_version(void) { return "cslam_b200 0.1.0 (sm_100a)"; }

const char* cslam_last_error(void) { return cslam::t_err; }

int cslam_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int64_t cslam_launch_count(void) { return cslam::g_launches.load(); }

}  // extern "C"
