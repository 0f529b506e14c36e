// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

namespace xp {

void set_error(const std::string& msg);
int fail(const std::string& msg);  // records msg, returns -1
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count();  // SMs of the current device (cached per device)

// 2-D bf16 tensor map with 128-byte swizzle.  inner/outer are extents in
// elements, row_stride in elements, box = {box_inner, box_outer}.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride,
                      uint32_t box_inner, uint32_t box_outer);

// Bind the CUDA context that owns `device_ptr` to the calling thread if the thread has none.  PyTorch runs
// autograd backward on worker threads that may not have touched CUDA yet; this library links its own
// (static) runtime, so without this a first call from such a thread would see "no current context" (or
// silently use device 0 on a multi-GPU rank).
int ensure_context(const void* device_ptr);
#define XP_ENTER(ptr)                                   \
  do {                                                  \
    if (::xp::ensure_context(ptr) != 0) return -1;      \
  } while (0)

#define XP_CHECK_CUDA(expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return ::xp::fail(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define XP_CHECK_LAUNCH(name)                                                                             \
  do {                                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                                  \
    if (_e != cudaSuccess) return ::xp::fail(std::string(name) + " launch: " + cudaGetErrorString(_e)); \
    ::xp::count_launch();                                                                                 \
  } while (0)

}  // namespace xp
