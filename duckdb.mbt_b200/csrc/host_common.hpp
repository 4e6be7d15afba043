// Host-side plumbing shared by the L1/L2 layers of libduckdb_mb_gpu: per-GPU context (streams,
// device + pinned memory pools, pinned staging ring) and the chunk stager.
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "dmb_common.cuh"

namespace dmb {

// NVTX range around a host-side phase (tracing).  Header-only NVTX v3: a no-op unless a profiler is attached.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

// Size-bucketed cache of device or page-locked host allocations.  cudaMalloc / cudaHostAlloc cost
// far more than a conversion (pinning runs at a few GB/s), so buffers are recycled across results
// of the same shape: after the first batch of a stream of equally shaped batches nothing is
// allocated any more.
class Pool {
 public:
  enum Kind { kDevice = 0, kPinned = 1 };
  explicit Pool(Kind kind) : kind_(kind) {}
  ~Pool() { release_all(); }
  void *alloc(size_t bytes);
  void free(void *p);
  void trim();         // drop cached (free) blocks
  void release_all();  // drop everything, including blocks still handed out
  size_t held_bytes() const { return held_; }

 private:
  Kind kind_;
  std::mutex mu_;
  std::multimap<size_t, void *> free_;
  std::unordered_map<void *, size_t> live_;
  size_t held_ = 0;
};

// Persistent host threads for the stager's gather memcpy (a 60 M-row pageable batch is ~650 ring fills: creating
// 15 threads per fill cost more than the copies).  run(n, fn) calls fn(i) for i in [0, n) on the caller + the workers.
class WorkerPool {
 public:
  WorkerPool() {}
  ~WorkerPool();
  void set_threads(int total) { total_ = total < 1 ? 1 : total; }  // participants, the caller included
  int threads() const { return total_; }
  void run(int64_t n, const std::function<void(int64_t)> &fn);

 private:
  void worker();
  int total_ = 1;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_job_, cv_done_;
  const std::function<void(int64_t)> *fn_ = nullptr;
  int64_t n_ = 0;
  std::atomic<int64_t> next_{0};
  uint64_t generation_ = 0;
  int active_ = 0;
  bool stop_ = false;
};

constexpr int kStageBuffers = 4;
constexpr size_t kStageBytes = 32u << 20;  // per ring buffer

// Everything a result / exported Arrow array needs to outlive the ctx handle.
struct CtxCore {
  int device = 0;
  cudaStream_t s_in = nullptr, s_compute = nullptr, s_out = nullptr;
  Pool dev{Pool::kDevice};
  Pool pin{Pool::kPinned};
  uint8_t *ring[kStageBuffers] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ring_free[kStageBuffers] = {nullptr, nullptr, nullptr, nullptr};
  int ring_next = 0;
  int stage_threads = 1;
  WorkerPool pool;
  // where a host-buffer call spends its wall time (ms, summed; printed per materialise when DMB_TRACE_STAGE is set)
  double t_gather = 0, t_ring_wait = 0, t_drain_wait = 0, t_stage = 0, t_launch = 0;
  std::mutex mu;  // one blocking call at a time per context
  ~CtxCore();
  bool bind() const { return check_cuda(cudaSetDevice(device), "cudaSetDevice") == 0; }
  int ring_acquire();  // index of a ring buffer whose previous copy has completed, or -1
};

}  // namespace dmb

struct duckdb_mb_gpu_ctx {
  std::shared_ptr<dmb::CtxCore> core;
};

namespace dmb {

inline double wall_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// run fn(i) for i in [0, n) on the context's host threads (gather memcpy into pinned staging, size passes)
template <typename F>
inline void parallel_for(CtxCore &core, int64_t n, F fn) {
  if (core.stage_threads <= 1 || n < 2) {
    for (int64_t i = 0; i < n; ++i) fn(i);
    return;
  }
  const std::function<void(int64_t)> f = fn;
  core.pool.run(n, f);
}

// Copy per-chunk host pieces into a device slab whose slot k starts at dst + k * slot_bytes.
//   src[k]        host pointer of piece k (NULL: skipped)
//   piece_bytes   bytes of piece k = counts[k] * row_bytes (row_bytes == 0: fixed `slot_bytes`)
//   pinned        pieces are page-locked: runs of full, address-contiguous pieces go by direct DMA
//   fixup         optional hooks on the staged copies (string_t compaction); forces the bounce path
// Everything else is gathered by host threads into the pinned ring and sent in 32 MiB pieces.
// The hook runs inside a gather task, right after the task's pieces [c0, c1) were copied to staged0 + (c - c0) * slot_bytes:
// the copies are still in the core's cache when it reads and rewrites them.
struct StageFixup {
  void (*task)(void *user, int64_t c0, int64_t c1, uint8_t *staged0, size_t slot_bytes);
  void *user;
};
int32_t stage_pieces(CtxCore &core, cudaStream_t stream, const void *const *src, const uint32_t *counts,
                     size_t row_bytes, size_t slot_bytes, int64_t nchunks, uint8_t *dst, bool pinned,
                     const StageFixup *fixup, uint64_t *bytes_moved);

// Allocations and events of ONE blocking call.  The destructor drains the three streams before
// anything goes back to the pools, also on error paths.
struct Scope {
  CtxCore &c;
  std::vector<void *> dev, pin;
  std::vector<cudaEvent_t> events;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kernel_spans;
  explicit Scope(CtxCore &core) : c(core) {}
  ~Scope() {
    cudaStreamSynchronize(c.s_in);
    cudaStreamSynchronize(c.s_compute);
    cudaStreamSynchronize(c.s_out);
    for (void *p : dev) c.dev.free(p);
    for (void *p : pin) c.pin.free(p);
    for (cudaEvent_t e : events) cudaEventDestroy(e);
  }
  void *dalloc(size_t bytes) {
    void *p = c.dev.alloc(bytes + 64);
    if (p) dev.push_back(p);
    return p;
  }
  void *palloc(size_t bytes) {
    void *p = c.pin.alloc(bytes + 64);
    if (p) pin.push_back(p);
    return p;
  }
  void *release_pin(void *p) {  // ownership moves to the caller
    for (size_t i = 0; i < pin.size(); ++i)
      if (pin[i] == p) { pin.erase(pin.begin() + (long)i); break; }
    return p;
  }
  cudaEvent_t event(bool timing) {
    cudaEvent_t e = nullptr;
    if (check_cuda(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming), "cudaEventCreate")) return nullptr;
    events.push_back(e);
    return e;
  }
  double kernel_ms() {
    double ms = 0;
    for (auto &kv : kernel_spans) {
      float f = 0;
      if (cudaEventElapsedTime(&f, kv.first, kv.second) == cudaSuccess) ms += f;
    }
    return ms;
  }
};

// pinned < 0: ask the driver whether `src` is page-locked
bool host_is_pinned(const void *p);
int32_t stage_contiguous(CtxCore &core, cudaStream_t stream, void *dst, const void *src, size_t bytes, int pinned,
                         uint64_t *bytes_moved);

}  // namespace dmb
