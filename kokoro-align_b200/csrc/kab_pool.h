// kab_pool.h -- per-device memory pool of the plan workspaces (internal to kab_api.cu: one
// translation unit, everything in an anonymous namespace; not a public header).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <unordered_map>

namespace {

// ---- device memory pool.  The drop-in call ctc_best_path() builds and destroys a plan per
// lattice, as the reference's per-chapter loop does (run_example.py:247-254): a dozen cudaMalloc /
// cudaFree pairs per call, and every cudaFree synchronises the device.  Freed blocks are kept per
// device (size classes of <= 12.5 % slack) and handed out again; kab_pool_trim() returns them to
// the driver, and the pool trims itself beyond POOL_CAP_BYTES per device.
constexpr int POOL_MAX_DEV = 64;
constexpr size_t POOL_CAP_BYTES = (size_t)16 << 30;
struct DevPool {
  std::mutex mu;
  std::multimap<size_t, void *> free_blocks[POOL_MAX_DEV];
  std::unordered_map<void *, std::pair<size_t, int>> live;  // block -> (class size, device)
  size_t cached[POOL_MAX_DEV] = {};
};
DevPool &pool() {
  static DevPool *p = new DevPool();  // never destroyed: plans may be freed during interpreter exit
  return *p;
}
size_t pool_class(size_t bytes) {
  if (bytes < 256) return 256;
  int lg = 63 - __builtin_clzll((unsigned long long)bytes);
  const size_t step = std::max<size_t>(256, (size_t)1 << (lg > 3 ? lg - 3 : 0));
  return (bytes + step - 1) / step * step;
}
void pool_trim_device(DevPool &P, int dev, size_t keep_bytes) {  // P.mu held
  int cur = 0;
  cudaGetDevice(&cur);
  bool switched = false;
  while (P.cached[dev] > keep_bytes && !P.free_blocks[dev].empty()) {
    auto it = std::prev(P.free_blocks[dev].end());  // largest first
    if (!switched && cur != dev) { cudaSetDevice(dev); switched = true; }
    cudaFree(it->second);
    P.cached[dev] -= it->first;
    P.free_blocks[dev].erase(it);
  }
  if (switched) cudaSetDevice(cur);
}
cudaError_t pool_malloc(void **out, size_t bytes) {  // on the current device
  *out = nullptr;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const size_t cls = pool_class(bytes);
  DevPool &P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  if (dev < POOL_MAX_DEV) {
    auto it = P.free_blocks[dev].lower_bound(cls);
    if (it != P.free_blocks[dev].end() && it->first == cls) {
      *out = it->second;
      P.cached[dev] -= cls;
      P.free_blocks[dev].erase(it);
      P.live[*out] = {cls, dev};
      return cudaSuccess;
    }
  }
  e = cudaMalloc(out, cls);
  if (e != cudaSuccess && dev < POOL_MAX_DEV && P.cached[dev]) {  // out of memory: give the cache back, retry
    cudaGetLastError();
    pool_trim_device(P, dev, 0);
    e = cudaMalloc(out, cls);
  }
  if (e == cudaSuccess) P.live[*out] = {cls, dev};
  return e;
}
void pool_free(void *ptr) {
  if (!ptr) return;
  DevPool &P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  auto it = P.live.find(ptr);
  if (it == P.live.end()) { cudaFree(ptr); return; }
  const size_t cls = it->second.first;
  const int dev = it->second.second;
  P.live.erase(it);
  if (dev >= POOL_MAX_DEV) { cudaFree(ptr); return; }
  P.free_blocks[dev].emplace(cls, ptr);
  P.cached[dev] += cls;
  if (P.cached[dev] > POOL_CAP_BYTES) pool_trim_device(P, dev, POOL_CAP_BYTES / 2);
}

}  // namespace
