// Does a non-inlined Fq multiplier (small code, I-cache resident) beat the fully inlined one inside XYZZ madd?
#include <cstdio>
#include <cuda_runtime.h>
#include "curve.cuh"
using namespace zkp;

__device__ __noinline__ Fq fq_mul_call(const Fq& a, const Fq& b) { return fp_mul(a, b); }

template <bool CALL> __device__ __forceinline__ Fq M(const Fq& a, const Fq& b) { if (CALL) return fq_mul_call(a, b); else return fp_mul(a, b); }

template <bool CALL>
__device__ __forceinline__ void madd(G1Xyzz& acc, const G1Affine& q) {
  Fq u2 = M<CALL>(q.x, acc.zz);
  Fq s2 = M<CALL>(q.y, acc.zzz);
  Fq p = fp_sub(u2, acc.x);
  Fq r = fp_sub(s2, acc.y);
  Fq pp = M<CALL>(p, p);
  Fq ppp = M<CALL>(p, pp);
  Fq qq = M<CALL>(acc.x, pp);
  Fq x3 = fp_sub(fp_sub(M<CALL>(r, r), ppp), fp_dbl(qq));
  Fq y3 = fp_sub(M<CALL>(r, fp_sub(qq, x3)), M<CALL>(acc.y, ppp));
  acc.x = x3; acc.y = y3;
  acc.zz = M<CALL>(acc.zz, pp);
  acc.zzz = M<CALL>(acc.zzz, ppp);
}

template <bool CALL, int MINB>
__global__ void __launch_bounds__(128, MINB) kmadd(const G1Affine* pts, G1Xyzz* out, uint32_t iters) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  G1Affine p = pts[t % 64];
  G1Xyzz acc = G1Xyzz::from_affine(pts[(t + 1) % 64]);
  for (uint32_t it = 0; it < iters; it++) {
    madd<CALL>(acc, p);
    p.x = acc.y;
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
  void* buf; cudaMalloc(&buf, (size_t)sms * 32 * 128 * 192 + 64 * 96); cudaMemset(buf, 1, (size_t)sms * 32 * 128 * 192 + 64 * 96);
  G1Affine* pts = (G1Affine*)buf; G1Xyzz* out = (G1Xyzz*)((char*)buf + 64 * 96);
#define RUN(CALL, MINB, BPS) { int blocks = sms * BPS; float ms = timeit(kmadd<CALL, MINB>, blocks, pts, out, 500u); \
    printf("madd call=%d minBlocks=%d blocks/SM=%d: %.3f ms  %.3e madd/s\n", CALL, MINB, BPS, ms, (double)blocks * 128 * 500 / (ms * 1e-3)); }
  RUN(false, 2, 2) RUN(false, 3, 3) RUN(false, 4, 4) RUN(true, 2, 2) RUN(true, 3, 3) RUN(true, 4, 4) RUN(true, 6, 6) RUN(true, 8, 8)
  return 0;
}
