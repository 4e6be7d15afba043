// Error plumbing shared by every layer of libduckdb_mb_gpu.
// The reference keeps one unsynchronised process-global string (src/duckdb_native.c:22-40,
// read back by duckdb_mb_last_error :240-246); here it is thread-local.

#include <stdarg.h>
#include <stdio.h>

#include "dmb_common.cuh"

namespace dmb {

static thread_local char g_error[512] = {0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int32_t check_cuda(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return -1;
}

}  // namespace dmb

extern "C" const char *duckdb_mb_gpu_last_error(void) { return dmb::g_error; }

extern "C" int32_t duckdb_mb_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" void *duckdb_mb_gpu_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (dmb::check_cuda(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault), "cudaHostAlloc")) return nullptr;
  return p;
}

extern "C" void duckdb_mb_gpu_host_free(void *p) {
  if (p) cudaFreeHost(p);
}
