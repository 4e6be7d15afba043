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

// NUMA placement ("pin staging memory NUMA-local to each GPU").  Binds the CALLING thread -- and with it
// every thread and page-locked allocation it makes from now on (threads inherit the mask, cudaHostAlloc / first touch place
// pages on the running CPU's node) -- to the CPUs of the NUMA node the GPU's PCIe root hangs off.  With one process per GPU
// and 8 GPUs on two sockets, unbound processes put about half of the staging traffic on the socket interconnect.
// Returns the node, or -1 when there is nothing to do (single node, no sysfs entry, an empty CPU set): never an error.
#include <sched.h>
#include <stdio.h>
#include <ctype.h>

extern "C" int32_t duckdb_mb_gpu_bind_numa(int32_t device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return -1; }
  for (char *p = bus; *p; ++p) *p = (char)tolower((unsigned char)*p);
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE *f = fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  if (node < 0) return -1;
  snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  f = fopen(path, "r");
  if (!f) return -1;
  char list[4096] = {0};
  const size_t got = fread(list, 1, sizeof(list) - 1, f);
  fclose(f);
  if (!got) return -1;
  cpu_set_t allowed, want;
  CPU_ZERO(&want);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return -1;
  // "0-31,64-95": ranges and single CPUs
  const char *p = list;
  int n_set = 0;
  while (*p) {
    while (*p && !isdigit((unsigned char)*p)) ++p;
    if (!*p) break;
    char *end = nullptr;
    long lo = strtol(p, &end, 10), hi = lo;
    p = end;
    if (*p == '-') { hi = strtol(p + 1, &end, 10); p = end; }
    for (long c = lo; c <= hi && c < CPU_SETSIZE; ++c)
      if (CPU_ISSET((int)c, &allowed)) { CPU_SET((int)c, &want); ++n_set; }
  }
  if (n_set == 0) return -1;  // the node's CPUs are outside this process's cpuset: leave the mask alone
  if (sched_setaffinity(0, sizeof(want), &want) != 0) return -1;
  return node;
}
