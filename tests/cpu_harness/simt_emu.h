// Minimal SIMT emulation for running small warp-collective CUDA kernels on the CPU (TEST INFRASTRUCTURE: the
// kernel SOURCE is extracted from the .cuh files by tests/test_simt_emulation.py and compiled against this header
// with g++).  One std::thread per CUDA thread, blocks one after another; warp collectives rendezvous on a per-warp
// barrier.  Preconditions as on the hardware: a collective with the full mask must be reached by all 32 lanes of a
// warp (warps may exit early as a whole).  __shared__ becomes a static (blocks are sequential, so it is per block).
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
static thread_local emu_dim3 emu_threadIdx, emu_blockIdx;
static emu_dim3 emu_blockDim, emu_gridDim;
#define threadIdx emu_threadIdx
#define blockIdx emu_blockIdx
#define blockDim emu_blockDim
#define gridDim emu_gridDim

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define S7B_HD inline

struct EmuWarp {
  pthread_barrier_t bar;
  unsigned slot[32];
  EmuWarp() { pthread_barrier_init(&bar, nullptr, 32); }
  ~EmuWarp() { pthread_barrier_destroy(&bar); }
};
static std::vector<EmuWarp>* emu_warps = nullptr;
static inline EmuWarp& emu_warp() { return (*emu_warps)[emu_threadIdx.x >> 5]; }
static inline int emu_lane() { return (int)(emu_threadIdx.x & 31); }

// every lane publishes v, then reads all 32
static inline void emu_exchange(unsigned v, unsigned (&all)[32]) {
  EmuWarp& w = emu_warp();
  w.slot[emu_lane()] = v;
  pthread_barrier_wait(&w.bar);
  for (int i = 0; i < 32; ++i) all[i] = w.slot[i];
  pthread_barrier_wait(&w.bar);
}
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu_warp().bar); }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __match_any_sync(unsigned, int v) {
  unsigned all[32];
  emu_exchange((unsigned)v, all);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) if (all[i] == (unsigned)v) m |= 1u << i;
  return m;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
  unsigned all[32];
  emu_exchange(v, all);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if (mask >> i & 1u) r = std::max(r, all[i]);
  return r;
}
static inline int __shfl_sync(unsigned, int v, int src) {
  unsigned all[32];
  emu_exchange((unsigned)v, all);
  return (int)all[src & 31];
}
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int lanemask) {
  unsigned all[32];
  emu_exchange(v, all);
  return all[(emu_lane() ^ lanemask) & 31];
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
  unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
using std::max;
using std::min;

// run kernel(args...) on a grid x block launch, blocks sequentially
static inline void emu_launch(unsigned grid, unsigned block, const std::function<void()>& kernel) {
  emu_blockDim.x = block;
  emu_gridDim.x = grid;
  for (unsigned b = 0; b < grid; ++b) {
    std::vector<EmuWarp> warps((block + 31) / 32);
    emu_warps = &warps;
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([&, t] {
        emu_threadIdx.x = t;
        emu_blockIdx.x = b;
        kernel();
      });
    for (auto& x : th) x.join();
  }
}
