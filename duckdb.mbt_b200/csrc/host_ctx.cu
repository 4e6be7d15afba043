// L1 plumbing: per-GPU context (three streams: host->device, kernels, device->host), recycled
// device / pinned allocations, and the chunk stager that turns thousands of <= 32 KiB DuckDB
// vectors into a few large cudaMemcpyAsync calls.
//
// The reference has no equivalent: its C stub reads libduckdb vectors in place, one cell per call
// (src/duckdb_native.c:520-667).  The device boundary is inserted here, inside the C layer, so the
// MoonBit side keeps calling blocking functions (SURVEY.md §8b "Threading").

#include <stdlib.h>
#include <string.h>

#include "host_common.hpp"

namespace dmb {

static size_t round_size(size_t bytes) {
  if (bytes < 256) bytes = 256;
  if (bytes < (1u << 20)) return (bytes + 4095) & ~(size_t)4095;
  return (bytes + ((1u << 20) - 1)) & ~(size_t)((1u << 20) - 1);
}

void *Pool::alloc(size_t bytes) {
  const size_t want = round_size(bytes);
  {
    std::lock_guard<std::mutex> g(mu_);
    auto it = free_.lower_bound(want);
    if (it != free_.end() && it->first <= want + want / 4 + (1u << 20)) {
      void *p = it->second;
      live_[p] = it->first;
      free_.erase(it);
      return p;
    }
  }
  void *p = nullptr;
  cudaError_t e = kind_ == kDevice ? cudaMalloc(&p, want) : cudaHostAlloc(&p, want, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    trim();  // give cached blocks back and retry once
    e = kind_ == kDevice ? cudaMalloc(&p, want) : cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("%s of %zu bytes failed: %s", kind_ == kDevice ? "cudaMalloc" : "cudaHostAlloc", want,
                cudaGetErrorString(e));
      return nullptr;
    }
  }
  std::lock_guard<std::mutex> g(mu_);
  live_[p] = want;
  held_ += want;
  return p;
}

void Pool::free(void *p) {
  if (!p) return;
  std::lock_guard<std::mutex> g(mu_);
  auto it = live_.find(p);
  if (it == live_.end()) return;
  free_.emplace(it->second, p);
  live_.erase(it);
}

void Pool::trim() {
  std::multimap<size_t, void *> drop;
  {
    std::lock_guard<std::mutex> g(mu_);
    drop.swap(free_);
    for (auto &kv : drop) held_ -= kv.first;
  }
  for (auto &kv : drop) {
    if (kind_ == kDevice) cudaFree(kv.second); else cudaFreeHost(kv.second);
  }
}

void Pool::release_all() {
  trim();
  std::unordered_map<void *, size_t> drop;
  {
    std::lock_guard<std::mutex> g(mu_);
    drop.swap(live_);
    held_ = 0;
  }
  for (auto &kv : drop) {
    if (kind_ == kDevice) cudaFree(kv.first); else cudaFreeHost(kv.first);
  }
}

WorkerPool::~WorkerPool() {
  {
    std::lock_guard<std::mutex> g(mu_);
    stop_ = true;
  }
  cv_job_.notify_all();
  for (auto &t : workers_) t.join();
}

void WorkerPool::worker() {
  uint64_t seen = 0;
  for (;;) {
    const std::function<void(int64_t)> *fn;
    int64_t n;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_job_.wait(lk, [&] { return stop_ || generation_ != seen; });
      if (stop_) return;
      seen = generation_;
      fn = fn_;
      n = n_;
    }
    for (;;) {
      const int64_t i = next_.fetch_add(1);
      if (i >= n) break;
      (*fn)(i);
    }
    {
      std::lock_guard<std::mutex> g(mu_);
      if (--active_ == 0) cv_done_.notify_one();
    }
  }
}

void WorkerPool::run(int64_t n, const std::function<void(int64_t)> &fn) {
  const int want = total_ - 1;
  if (want <= 0 || n < 2) {
    for (int64_t i = 0; i < n; ++i) fn(i);
    return;
  }
  while ((int)workers_.size() < want) workers_.emplace_back([this] { worker(); });  // (one blocking call at a time per context)
  {
    std::lock_guard<std::mutex> g(mu_);
    fn_ = &fn;
    n_ = n;
    next_.store(0);
    active_ = (int)workers_.size();
    ++generation_;
  }
  cv_job_.notify_all();
  for (;;) {
    const int64_t i = next_.fetch_add(1);
    if (i >= n) break;
    fn(i);
  }
  std::unique_lock<std::mutex> lk(mu_);
  cv_done_.wait(lk, [&] { return active_ == 0; });
}

CtxCore::~CtxCore() {
  cudaSetDevice(device);
  if (s_in) cudaStreamSynchronize(s_in);
  if (s_compute) cudaStreamSynchronize(s_compute);
  if (s_out) cudaStreamSynchronize(s_out);
  for (int i = 0; i < kStageBuffers; ++i) {
    if (ring_free[i]) cudaEventDestroy(ring_free[i]);
    if (ring[i]) cudaFreeHost(ring[i]);
  }
  dev.release_all();
  pin.release_all();
  if (s_in) cudaStreamDestroy(s_in);
  if (s_compute) cudaStreamDestroy(s_compute);
  if (s_out) cudaStreamDestroy(s_out);
}

int CtxCore::ring_acquire() {
  const int b = ring_next;
  ring_next = (ring_next + 1) % kStageBuffers;
  if (!ring[b]) {
    if (check_cuda(cudaHostAlloc((void **)&ring[b], kStageBytes, cudaHostAllocDefault), "staging ring cudaHostAlloc")) return -1;
    if (check_cuda(cudaEventCreateWithFlags(&ring_free[b], cudaEventDisableTiming), "staging ring event")) return -1;
  } else {
    const double t0 = wall_ms();
    if (check_cuda(cudaEventSynchronize(ring_free[b]), "staging ring wait")) return -1;
    t_ring_wait += wall_ms() - t0;
  }
  return b;
}

int32_t stage_pieces(CtxCore &core, cudaStream_t stream, const void *const *src, const uint32_t *counts,
                     size_t row_bytes, size_t slot_bytes, int64_t nchunks, uint8_t *dst, bool pinned,
                     const StageFixup *fixup, uint64_t *bytes_moved) {
  auto piece = [&](int64_t k) -> size_t { return row_bytes ? (size_t)counts[k] * row_bytes : slot_bytes; };
  const int64_t per_buf = (int64_t)(kStageBytes / slot_bytes);
  if (per_buf < 1) { set_error("stage_pieces: slot of %zu bytes exceeds the staging buffer", slot_bytes); return -1; }
  int64_t k = 0;
  while (k < nchunks) {
    if (!src[k] || piece(k) == 0) { ++k; continue; }
    if (pinned && !fixup) {
      // direct DMA for a run of address-contiguous pieces; every piece but the last must be full
      int64_t m = k;
      while (m + 1 < nchunks && piece(m) == slot_bytes && src[m + 1] &&
             (const uint8_t *)src[m + 1] == (const uint8_t *)src[m] + slot_bytes)
        ++m;
      const size_t bytes = (size_t)(m - k) * slot_bytes + piece(m);
      if (bytes >= (64u << 10) || m == nchunks - 1) {
        if (check_cuda(cudaMemcpyAsync(dst + (size_t)k * slot_bytes, src[k], bytes, cudaMemcpyHostToDevice, stream),
                       "chunk H2D (direct)")) return -1;
        if (bytes_moved) *bytes_moved += bytes;
        k = m + 1;
        continue;
      }
    }
    // bounce: gather pieces k..m-1 into one pinned ring buffer, send them with one copy
    const int b = core.ring_acquire();
    if (b < 0) return -1;
    int64_t m = k + per_buf < nchunks ? k + per_buf : nchunks;
    uint8_t *buf = core.ring[b];
    const int64_t npieces = m - k;
    const int64_t group = 64;  // pieces per parallel task
    const double t_g0 = wall_ms();
    parallel_for(core, (npieces + group - 1) / group, [&](int64_t g) {
      const int64_t i1 = (g + 1) * group < npieces ? (g + 1) * group : npieces;
      for (int64_t i = g * group; i < i1; ++i) {
        const int64_t c = k + i;
        const size_t bytes = piece(c);
        if (!src[c] || !bytes) continue;
        memcpy(buf + (size_t)i * slot_bytes, src[c], bytes);
      }
      if (fixup) fixup->task(fixup->user, k + g * group, k + i1, buf + (size_t)(g * group) * slot_bytes, slot_bytes);
    });
    core.t_gather += wall_ms() - t_g0;
    int64_t last = m - 1;
    while (last > k && (!src[last] || piece(last) == 0)) --last;
    const size_t bytes = (size_t)(last - k) * slot_bytes + piece(last);
    if (check_cuda(cudaMemcpyAsync(dst + (size_t)k * slot_bytes, buf, bytes, cudaMemcpyHostToDevice, stream),
                   "chunk H2D (staged)")) return -1;
    if (check_cuda(cudaEventRecord(core.ring_free[b], stream), "staging ring record")) return -1;
    if (bytes_moved) *bytes_moved += bytes;
    k = m;
  }
  return 0;
}

// One contiguous host buffer -> device.  Page-locked sources go by direct DMA; pageable ones are
// copied into the pinned ring by `stage_threads` host threads, 16 MiB at a time, so that the
// memcpy of piece i+1 overlaps the DMA of piece i.
bool host_is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int32_t stage_contiguous(CtxCore &core, cudaStream_t stream, void *dst, const void *src, size_t bytes, int pinned,
                         uint64_t *bytes_moved) {
  if (!bytes) return 0;
  if (pinned < 0) pinned = host_is_pinned(src) ? 1 : 0;
  if (pinned) {
    if (check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream), "H2D (direct)")) return -1;
    if (bytes_moved) *bytes_moved += bytes;
    return 0;
  }
  for (size_t off = 0; off < bytes; off += kStageBytes) {
    const size_t piece = bytes - off < kStageBytes ? bytes - off : kStageBytes;
    const int b = core.ring_acquire();
    if (b < 0) return -1;
    uint8_t *buf = core.ring[b];
    const size_t sub = 1u << 20;
    const int64_t nsub = (int64_t)((piece + sub - 1) / sub);
    const uint8_t *s = (const uint8_t *)src + off;
    parallel_for(core, nsub, [&](int64_t i) {
      const size_t o = (size_t)i * sub;
      memcpy(buf + o, s + o, piece - o < sub ? piece - o : sub);
    });
    if (check_cuda(cudaMemcpyAsync((uint8_t *)dst + off, buf, piece, cudaMemcpyHostToDevice, stream), "H2D (staged)")) return -1;
    if (check_cuda(cudaEventRecord(core.ring_free[b], stream), "staging ring record")) return -1;
  }
  if (bytes_moved) *bytes_moved += bytes;
  return 0;
}

}  // namespace dmb

using namespace dmb;

extern "C" duckdb_mb_gpu_ctx *duckdb_mb_gpu_ctx_create(int32_t device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("duckdb_mb_gpu_ctx_create: no CUDA device is visible (this library has no CPU fallback)");
    return nullptr;
  }
  if (device < 0 || device >= n) { set_error("duckdb_mb_gpu_ctx_create: device %d out of range (%d visible)", device, n); return nullptr; }
  auto core = std::make_shared<CtxCore>();
  core->device = device;
  if (!core->bind()) return nullptr;
  if (check_cuda(cudaStreamCreateWithFlags(&core->s_in, cudaStreamNonBlocking), "stream create") ||
      check_cuda(cudaStreamCreateWithFlags(&core->s_compute, cudaStreamNonBlocking), "stream create") ||
      check_cuda(cudaStreamCreateWithFlags(&core->s_out, cudaStreamNonBlocking), "stream create"))
    return nullptr;
  const char *env = getenv("DMB_STAGE_THREADS");
  int t = env ? atoi(env) : (int)std::thread::hardware_concurrency();
  core->stage_threads = t < 1 ? 1 : (t > 32 ? 32 : t);
  core->pool.set_threads(core->stage_threads);
  duckdb_mb_gpu_ctx *ctx = new duckdb_mb_gpu_ctx();
  ctx->core = core;
  return ctx;
}

extern "C" void duckdb_mb_gpu_ctx_destroy(duckdb_mb_gpu_ctx *ctx) {
  if (!ctx) return;
  delete ctx;  // results and exported Arrow arrays keep the core (pools, streams) alive
}

extern "C" int32_t duckdb_mb_gpu_ctx_sync(duckdb_mb_gpu_ctx *ctx) {
  if (!ctx) { set_error("duckdb_mb_gpu_ctx_sync: null context"); return 0; }
  CtxCore &c = *ctx->core;
  if (!c.bind()) return 0;
  if (check_cuda(cudaStreamSynchronize(c.s_in), "sync") || check_cuda(cudaStreamSynchronize(c.s_compute), "sync") ||
      check_cuda(cudaStreamSynchronize(c.s_out), "sync"))
    return 0;
  return 1;
}
