// Field-multiplication throughput on sm_100a: the practical ceiling for MSM (Fq) and NTT (Fr).
#include <cstdio>
#include <cuda_runtime.h>
#include "curve.cuh"
using namespace zkp;

template <class F, int ILP, int MINB>
__global__ void __launch_bounds__(128, MINB) kmul(F* io, uint32_t iters) {
  F a[ILP], b[ILP];
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < ILP; i++) { a[i] = io[t]; b[i] = io[t]; a[i].v[0] += i; b[i].v[1] ^= i; }
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = fp_mul(a[i], b[i]); }
#pragma unroll
    for (int i = 0; i < ILP; i++) { b[i] = fp_mul(b[i], a[i]); }
  }
  F r = a[0];
#pragma unroll
  for (int i = 0; i < ILP; i++) { r = fp_add(r, a[i]); r = fp_add(r, b[i]); }
  io[t] = r;
}

// XYZZ mixed addition throughput (the MSM inner loop without memory)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) kmadd(G1Affine* pts, G1Xyzz* out, uint32_t iters) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  G1Affine p = pts[t % 64];
  G1Xyzz acc = G1Xyzz::from_affine(pts[(t + 1) % 64]);
  for (uint32_t it = 0; it < iters; it++) {
    xyzz_madd(acc, p);
    p.x = acc.y;  // keep the operand changing (not a curve point, arithmetic cost identical)
  }
  out[t] = acc;
}

template <class K, class... A>
float timeit(K kern, int blocks, A... args) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); kern<<<blocks, 128>>>(args...); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t iters = 2000;
  void* buf; cudaMalloc(&buf, (size_t)sms * 32 * 128 * 192); cudaMemset(buf, 1, (size_t)sms * 32 * 128 * 192);
#define RUN(F, NAME, ILP, MINB, BPS) { int blocks = sms * BPS; float ms = timeit(kmul<F, ILP, MINB>, blocks, (F*)buf, iters); \
    double muls = (double)blocks * 128 * iters * 2 * ILP; printf("%s mul ILP=%d minBlocks=%d blocks/SM=%d: %.3f ms  %.3e mul/s\n", NAME, ILP, MINB, BPS, ms, muls / (ms * 1e-3)); }
  RUN(Fq, "Fq", 1, 1, 4) RUN(Fq, "Fq", 1, 4, 8) RUN(Fq, "Fq", 2, 1, 4) RUN(Fq, "Fq", 2, 3, 6) RUN(Fq, "Fq", 1, 8, 8) RUN(Fq, "Fq", 1, 8, 16)
  RUN(Fr, "Fr", 1, 1, 4) RUN(Fr, "Fr", 1, 4, 8) RUN(Fr, "Fr", 2, 4, 8) RUN(Fr, "Fr", 4, 2, 4) RUN(Fr, "Fr", 1, 8, 16)
  {
    int blocks = sms * 3; float ms = timeit(kmadd<3>, blocks, (G1Affine*)buf, (G1Xyzz*)((char*)buf + 64 * 96), 500u);
    printf("xyzz_madd minBlocks=3 blocks/SM=3: %.3f ms  %.3e madd/s\n", ms, (double)blocks * 128 * 500 / (ms * 1e-3));
    blocks = sms * 4; ms = timeit(kmadd<4>, blocks, (G1Affine*)buf, (G1Xyzz*)((char*)buf + 64 * 96), 500u);
    printf("xyzz_madd minBlocks=4 blocks/SM=4: %.3f ms  %.3e madd/s\n", ms, (double)blocks * 128 * 500 / (ms * 1e-3));
    blocks = sms * 2; ms = timeit(kmadd<2>, blocks, (G1Affine*)buf, (G1Xyzz*)((char*)buf + 64 * 96), 500u);
    printf("xyzz_madd minBlocks=2 blocks/SM=2: %.3f ms  %.3e madd/s\n", ms, (double)blocks * 128 * 500 / (ms * 1e-3));
  }
  return 0;
}
