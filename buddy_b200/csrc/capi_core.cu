// Library-wide C ABI plumbing: error string, version, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/buddy_b200.h"
#include "common.cuh"

namespace buddy {
static thread_local char g_err[512] = "";
extern std::atomic<long long> g_launches;

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_last_error("%s: %s", what, cudaGetErrorString(e));
  return BUDDY_ERR_CUDA;
}
}  // namespace buddy

extern "C" const char* buddy_last_error(void) { return buddy::g_err; }
extern "C" int buddy_version(void) { return 100; }
extern "C" int64_t buddy_launch_count(void) { return buddy::g_launches.load(); }
extern "C" void buddy_reset_launch_count(void) { buddy::g_launches.store(0); }
