// Test infrastructure: just enough of the CUDA execution model on the host to run the DEVICE functions of
// dpgo_b200/csrc/kernels.cuh for ONE CTA of 256 threads with real threads -- __syncthreads is a barrier,
// warp shuffles exchange through a per-warp buffer, __shared__ is static storage (one CTA), a TMA bulk copy
// is a memcpy that completes a phase of an emulated mbarrier (with the real parity semantics, so a wrong
// parity bookkeeping deadlocks the test instead of passing).  Never part of the product libraries.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>

#define DPGO_CPU_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __shared__ static
#define __align__(x) alignas(x)

struct EmuDim3 { unsigned x, y, z; };
inline thread_local EmuDim3 threadIdx{0, 0, 0};
inline const EmuDim3 blockIdx{0, 0, 0}, gridDim{1, 1, 1}, blockDim{256, 1, 1};
struct double2 { double x, y; };

namespace emu {
constexpr int kThreads = 256, kWarps = kThreads / 32;
inline std::barrier<> cta_barrier(kThreads);
inline std::barrier<> *warp_barrier[kWarps];
inline double warp_buf[kWarps][32];
inline unsigned char *dsm = nullptr;              // base of the emulated dynamic shared memory
inline std::atomic<uint32_t> mbar_done[64];       // completed phases per mbarrier
inline int mbar_slot(uint32_t bar) { return (int)((bar >> 3) & 63u); }
}  // namespace emu

inline void __syncthreads() { emu::cta_barrier.arrive_and_wait(); }

// all 32 lanes of the calling warp take part (true for every use in kernels.cuh)
inline double emu_shfl(double v, int src) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  emu::warp_buf[w][lane] = v;
  emu::warp_barrier[w]->arrive_and_wait();
  const double out = emu::warp_buf[w][(src >= 0 && src < 32) ? src : lane];
  emu::warp_barrier[w]->arrive_and_wait();
  return out;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier[threadIdx.x >> 5]->arrive_and_wait(); }
inline double __shfl_sync(unsigned, double v, int src) { return emu_shfl(v, src & 31); }
inline double __shfl_down_sync(unsigned, double v, int delta) {
  const int lane = threadIdx.x & 31;
  return emu_shfl(v, lane + delta < 32 ? lane + delta : lane);
}
inline double __shfl_xor_sync(unsigned, double v, int m) { return emu_shfl(v, (int)((threadIdx.x & 31) ^ m)); }

template <typename T>
inline T __ldg(const T *p) { return *p; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }

// ---- the PTX helpers of kernels.cuh -------------------------------------------------------------------
inline uint32_t smem_u32(const void *p) { return (uint32_t)(static_cast<const unsigned char *>(p) - emu::dsm); }
inline void mbar_init(uint32_t bar, uint32_t) { emu::mbar_done[emu::mbar_slot(bar)].store(0); }
inline void mbar_fence_init() {}
inline void mbar_expect_tx(uint32_t, uint32_t) {}
// try_wait.parity(P) succeeds once the phase with parity P has completed, i.e. the current phase has parity != P
inline void mbar_wait(uint32_t bar, uint32_t parity) {
  while ((emu::mbar_done[emu::mbar_slot(bar)].load(std::memory_order_acquire) & 1u) == parity) std::this_thread::yield();
}
// the copy lands at once and completes the barrier's phase (one arrival + all bytes)
inline void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  std::memcpy(emu::dsm + dst, src, bytes);
  emu::mbar_done[emu::mbar_slot(bar)].fetch_add(1, std::memory_order_release);
}
inline void bulk_prefetch_l2(const void *, uint32_t) {}

// ---- what the fused solver kernel (fused_kernel.cuh) needs on top -------------------------------------------
#include <chrono>
#define __launch_bounds__(...)
#define DPGO_DYNAMIC_SMEM(name) unsigned char *name = emu::dsm
namespace cg {
struct grid_group {
  void sync() const { __syncthreads(); }   // one CTA: the grid barrier is the CTA barrier
};
inline grid_group this_grid() { return {}; }
}  // namespace cg
inline unsigned long long gtimer() {
  return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}
