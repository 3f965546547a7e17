// Thin device-runtime vocabulary shared by every kernel file.
//
// Real build (nvcc, sm_100a): CUDA runtime calls, `<<<>>>` launches.
// Emulated build (g++ -DZKP_EMU, tests only): the same kernels run on the CPU thread emulator in
// tests/emu/cuda_emu.h so indexing logic is checked without a GPU.  The emulated library is never
// loaded by the product path.
#pragma once
#include <stddef.h>
#include <stdint.h>

enum ZkpStatus {
  ZKP_OK = 0,
  ZKP_ERR_INVALID_ARG = 1,
  ZKP_ERR_CUDA = 2,
  ZKP_ERR_OOM = 3,
  ZKP_ERR_SRS_TOO_SMALL = 4,   // kzg/src/scheme.rs:86 assert!(g1_points.len() > polynomial.degree())
  ZKP_ERR_DOMAIN_TOO_LARGE = 5,  // GeneralEvaluationDomain::new(..) == None
  ZKP_ERR_NO_DEVICE = 6,
  ZKP_ERR_EMPTY_POLY = 7,      // kzg/src/scheme.rs:112 expect("at least 1")
};

#ifdef ZKP_EMU
#include "cuda_emu.h"
#include <stdlib.h>
#include <string.h>
namespace zkp { namespace rt {
inline int dev_malloc(void** p, size_t bytes) { *p = bytes ? aligned_alloc(256, (bytes + 255) & ~(size_t)255) : nullptr; return (*p || !bytes) ? ZKP_OK : ZKP_ERR_OOM; }
inline void dev_free(void* p) { free(p); }
inline int host_malloc_pinned(void** p, size_t bytes) { return dev_malloc(p, bytes); }
inline void host_free_pinned(void* p) { free(p); }
inline int h2d(void* d, const void* h, size_t bytes, cudaStream_t) { if (bytes) memcpy(d, h, bytes); return ZKP_OK; }
inline int d2h(void* h, const void* d, size_t bytes, cudaStream_t) { if (bytes) memcpy(h, d, bytes); return ZKP_OK; }
inline int d2d(void* dst, const void* src, size_t bytes, cudaStream_t) { if (bytes) memmove(dst, src, bytes); return ZKP_OK; }
inline int copy_any(void* dst, const void* src, size_t bytes, cudaStream_t) { if (bytes) memmove(dst, src, bytes); return ZKP_OK; }
inline int dev_memset(void* d, int v, size_t bytes, cudaStream_t) { if (bytes) memset(d, v, bytes); return ZKP_OK; }
inline int sync(cudaStream_t) { return ZKP_OK; }
inline int check_last() { return ZKP_OK; }
inline int set_device(int) { return ZKP_OK; }
inline int device_count() { return 1; }
inline int sm_count(int) { return 4; }
inline int allow_smem(const void*, size_t) { return ZKP_OK; }
inline int prefer_smem_carveout(const void*) { return ZKP_OK; }
inline int ipc_export(const void* d, uint8_t h[64]) { memset(h, 0, 64); memcpy(h, &d, sizeof(d)); return ZKP_OK; }  // same process only
inline int ipc_open(const uint8_t h[64], void** d) { memcpy(d, h, sizeof(*d)); return ZKP_OK; }
inline int ipc_close(void*) { return ZKP_OK; }
inline const char* last_error_string() { return "emulator"; }
}}  // namespace zkp::rt
#else
#include <cuda_runtime.h>
#include <string.h>
#define ZKP_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// kernels without barriers / shuffles / shared memory: same launch here; the CPU emulator runs them without its
// per-thread barrier machinery (tests/emu/cuda_emu.h)
#define ZKP_LAUNCH_NOSYNC(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define ZKP_DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(16) unsigned char zkp_dyn_smem_raw[]; \
  type* name = reinterpret_cast<type*>(zkp_dyn_smem_raw)
namespace zkp { namespace rt {
inline const char*& last_error_slot() { static thread_local const char* s = ""; return s; }
inline int wrap(cudaError_t e) {
  if (e == cudaSuccess) return ZKP_OK;
  last_error_slot() = cudaGetErrorString(e);
  return e == cudaErrorMemoryAllocation ? ZKP_ERR_OOM : ZKP_ERR_CUDA;
}
inline int dev_malloc(void** p, size_t bytes) { *p = nullptr; if (!bytes) return ZKP_OK; return wrap(cudaMalloc(p, bytes)); }
inline void dev_free(void* p) { if (p) cudaFree(p); }
inline int host_malloc_pinned(void** p, size_t bytes) { *p = nullptr; if (!bytes) return ZKP_OK; return wrap(cudaMallocHost(p, bytes)); }
inline void host_free_pinned(void* p) { if (p) cudaFreeHost(p); }
inline int h2d(void* d, const void* h, size_t bytes, cudaStream_t s) { return bytes ? wrap(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s)) : ZKP_OK; }
inline int d2h(void* h, const void* d, size_t bytes, cudaStream_t s) { return bytes ? wrap(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s)) : ZKP_OK; }
inline int d2d(void* dst, const void* src, size_t bytes, cudaStream_t s) { return bytes ? wrap(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s)) : ZKP_OK; }
// unified-addressing copy: source and destination may live on different devices (peer copy over NVLink when enabled)
inline int copy_any(void* dst, const void* src, size_t bytes, cudaStream_t s) { return bytes ? wrap(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s)) : ZKP_OK; }
inline int dev_memset(void* d, int v, size_t bytes, cudaStream_t s) { return bytes ? wrap(cudaMemsetAsync(d, v, bytes, s)) : ZKP_OK; }
inline int sync(cudaStream_t s) { return wrap(cudaStreamSynchronize(s)); }
inline int check_last() { return wrap(cudaGetLastError()); }
inline int set_device(int d) { return wrap(cudaSetDevice(d)); }
inline int device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
inline int sm_count(int dev) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }
inline int allow_smem(const void* fn, size_t bytes) {
  return wrap(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}
inline int prefer_smem_carveout(const void* fn) {
  return wrap(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
}
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
inline int ipc_export(const void* d, uint8_t h[64]) { return wrap(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)h, (void*)d)); }
inline int ipc_open(const uint8_t h[64], void** d) {
  cudaIpcMemHandle_t hh;
  memcpy(&hh, h, 64);
  return wrap(cudaIpcOpenMemHandle(d, hh, cudaIpcMemLazyEnablePeerAccess));
}
inline int ipc_close(void* d) { return d ? wrap(cudaIpcCloseMemHandle(d)) : ZKP_OK; }
inline const char* last_error_string() { return last_error_slot(); }
}}  // namespace zkp::rt
#endif

#define ZKP_TRY(expr)            \
  do {                           \
    int zkp_status_ = (expr);    \
    if (zkp_status_ != ZKP_OK) return zkp_status_; \
  } while (0)
