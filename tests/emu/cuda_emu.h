// Minimal CUDA execution-model emulator for the CPU test-suite (TEST INFRASTRUCTURE ONLY).
//
// The build container has no GPU and every GPU round trip costs minutes, so the kernels under
// zkp-implementation_b200/csrc/ are also compiled by g++ with -DZKP_EMU against this header:
// one OS thread per CUDA thread of a block, blocks executed one after another, `__syncthreads()` as
// a pthread barrier, `__shared__` as a static (one block is live at a time), warp shuffles through
// a per-warp exchange slot.  It checks indexing / control-flow logic of the kernels bit-for-bit;
// it says nothing about performance and is never loaded by the product path (only tests/ build and
// load `tests/emu/_build/libzkp_b200_emu.so`).
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __align__(x) __attribute__((aligned(x)))

struct uint4 { uint32_t x, y, z, w; } __attribute__((aligned(16)));
struct uint2 { uint32_t x, y; } __attribute__((aligned(8)));
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef int cudaStream_t;

namespace zkp_emu {
struct State {
  pthread_barrier_t block_barrier;
  std::vector<pthread_barrier_t> warp_barriers;
  std::vector<uint64_t> warp_slots;  // [warp][lane]
  dim3 grid, block;
  unsigned nthreads = 0;
};
inline State*& state() { static State* s = nullptr; return s; }
inline unsigned char*& dyn_smem() { static unsigned char* p = nullptr; return p; }
}  // namespace zkp_emu

static thread_local dim3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;
static thread_local unsigned zkp_emu_tid;

static inline void __syncthreads() { pthread_barrier_wait(&zkp_emu::state()->block_barrier); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  pthread_barrier_wait(&zkp_emu::state()->warp_barriers[zkp_emu_tid >> 5]);
}
static inline void __threadfence() { __sync_synchronize(); }

template <class T>
static inline T zkp_emu_shfl(T v, unsigned src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  auto* st = zkp_emu::state();
  unsigned w = zkp_emu_tid >> 5, lane = zkp_emu_tid & 31;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  st->warp_slots[w * 32 + lane] = raw;
  pthread_barrier_wait(&st->warp_barriers[w]);
  unsigned base = w * 32;
  unsigned nl = std::min(32u, st->nthreads - base);
  uint64_t got = st->warp_slots[base + (src_lane < nl ? src_lane : lane)];
  pthread_barrier_wait(&st->warp_barriers[w]);
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return zkp_emu_shfl(v, (unsigned)src & 31); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return zkp_emu_shfl(v, (zkp_emu_tid & 31) ^ (unsigned)m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  unsigned lane = zkp_emu_tid & 31;
  return zkp_emu_shfl(v, lane + d < 32 ? lane + d : lane);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  // one exchange round: every lane publishes its predicate, then reads the whole warp's
  auto* st = zkp_emu::state();
  const unsigned w = zkp_emu_tid >> 5, lane = zkp_emu_tid & 31;
  st->warp_slots[w * 32 + lane] = pred ? 1u : 0u;
  pthread_barrier_wait(&st->warp_barriers[w]);
  const unsigned base = w * 32;
  const unsigned nl = std::min(32u, st->nthreads - base);
  unsigned acc = 0;
  for (unsigned l = 0; l < nl; l++) acc |= (unsigned)(st->warp_slots[base + l] & 1u) << l;
  pthread_barrier_wait(&st->warp_barriers[w]);
  return acc;
}

static inline unsigned __match_any_sync(unsigned, unsigned value) {
  // one exchange round: every lane publishes its value, then collects the lanes that published the same one
  auto* st = zkp_emu::state();
  const unsigned w = zkp_emu_tid >> 5, lane = zkp_emu_tid & 31;
  st->warp_slots[w * 32 + lane] = value;
  pthread_barrier_wait(&st->warp_barriers[w]);
  const unsigned base = w * 32;
  const unsigned nl = std::min(32u, st->nthreads - base);
  unsigned acc = 0;
  for (unsigned l = 0; l < nl; l++) acc |= (unsigned)(st->warp_slots[base + l] == (uint64_t)value) << l;
  pthread_barrier_wait(&st->warp_barriers[w]);
  return acc;
}

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicMax(unsigned* p, unsigned v) {
  unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline unsigned __brev(unsigned x) {
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
  return (x >> 16) | (x << 16);
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

namespace zkp_emu {
template <class Body>
void launch(dim3 grid, dim3 block, size_t smem_bytes, Body body) {
  State st;
  st.grid = grid;
  st.block = block;
  st.nthreads = block.x * block.y * block.z;
  unsigned nwarps = (st.nthreads + 31) / 32;
  pthread_barrier_init(&st.block_barrier, nullptr, st.nthreads);
  st.warp_barriers.resize(nwarps);
  for (unsigned w = 0; w < nwarps; w++)
    pthread_barrier_init(&st.warp_barriers[w], nullptr, std::min(32u, st.nthreads - w * 32));
  st.warp_slots.assign(nwarps * 32, 0);
  std::vector<unsigned char> smem(smem_bytes + 64);
  unsigned char* smem_aligned = (unsigned char*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
  state() = &st;
  dyn_smem() = smem_aligned;
  auto worker = [&](unsigned tid) {
    zkp_emu_tid = tid;
    blockDim = block;
    gridDim = grid;
    threadIdx = dim3(tid % block.x, (tid / block.x) % block.y, tid / (block.x * block.y));
    for (unsigned bz = 0; bz < grid.z; bz++)
      for (unsigned by = 0; by < grid.y; by++)
        for (unsigned bx = 0; bx < grid.x; bx++) {
          blockIdx = dim3(bx, by, bz);
          body();
          pthread_barrier_wait(&st.block_barrier);  // block boundary: static __shared__ is reused
        }
  };
  if (st.nthreads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    th.reserve(st.nthreads);
    for (unsigned t = 0; t < st.nthreads; t++) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  pthread_barrier_destroy(&st.block_barrier);
  for (auto& b : st.warp_barriers) pthread_barrier_destroy(&b);
  state() = nullptr;
}
}  // namespace zkp_emu

namespace zkp_emu {
// Kernels that never synchronise (no __syncthreads, no shuffles, no __shared__): the CUDA threads of a block run
// one after another on one OS thread and the blocks are spread over the host cores -- no barrier traffic at all.
// Calling a barrier from such a kernel dereferences the null state and crashes, which is the check.
template <class Body>
void launch_nosync(dim3 grid, dim3 block, Body body) {
  const unsigned nthreads = block.x * block.y * block.z;
  const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  unsigned workers = std::thread::hardware_concurrency();
  if (workers == 0) workers = 4;
  if (workers > 16) workers = 16;
  if ((size_t)workers > nblocks) workers = (unsigned)nblocks;
  auto run = [&](unsigned wi) {
    blockDim = block;
    gridDim = grid;
    for (size_t b = wi; b < nblocks; b += workers) {
      blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((size_t)grid.x * grid.y)));
      for (unsigned tid = 0; tid < nthreads; tid++) {
        zkp_emu_tid = tid;
        threadIdx = dim3(tid % block.x, (tid / block.x) % block.y, tid / (block.x * block.y));
        body();
      }
    }
  };
  if (workers <= 1) {
    run(0);
  } else {
    std::vector<std::thread> th;
    th.reserve(workers);
    for (unsigned w = 0; w < workers; w++) th.emplace_back(run, w);
    for (auto& t : th) t.join();
  }
}
}  // namespace zkp_emu

// Kernel launch + dynamic shared memory vocabulary shared with the real build (see csrc/runtime.h).
#define ZKP_LAUNCH(kernel, grid, block, smem, stream, ...) \
  zkp_emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define ZKP_LAUNCH_NOSYNC(kernel, grid, block, smem, stream, ...) \
  zkp_emu::launch_nosync((grid), (block), [&]() { kernel(__VA_ARGS__); })
#define ZKP_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(zkp_emu::dyn_smem())
