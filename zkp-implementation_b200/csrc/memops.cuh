// Explicit 128-bit global/shared memory accessors for field elements and curve points.
// Every kernel moves Fr (32 B), Fq (48 B), affine (96 B) and XYZZ (192 B) records through these so
// that accesses are LDG.128 / STG.128 and never depend on how the compiler lowers a struct copy.
#pragma once
#include "curve.cuh"

namespace zkp {

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ Fr ldg_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ Fq ld_fq(const Fq* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1], c = q[2];
  Fq r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  r.v[8] = c.x; r.v[9] = c.y; r.v[10] = c.z; r.v[11] = c.w;
  return r;
}
__device__ __forceinline__ void st_fq(Fq* p, const Fq& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
  q[2] = make_uint4(r.v[8], r.v[9], r.v[10], r.v[11]);
}
__device__ __forceinline__ G1Affine ld_affine(const G1Affine* p) {
  G1Affine r;
  r.x = ld_fq(&p->x);
  r.y = ld_fq(&p->y);
  return r;
}
__device__ __forceinline__ void st_affine(G1Affine* p, const G1Affine& r) {
  st_fq(&p->x, r.x);
  st_fq(&p->y, r.y);
}
__device__ __forceinline__ G1Xyzz ld_xyzz(const G1Xyzz* p) {
  G1Xyzz r;
  r.x = ld_fq(&p->x); r.y = ld_fq(&p->y); r.zz = ld_fq(&p->zz); r.zzz = ld_fq(&p->zzz);
  return r;
}
__device__ __forceinline__ void st_xyzz(G1Xyzz* p, const G1Xyzz& r) {
  st_fq(&p->x, r.x); st_fq(&p->y, r.y); st_fq(&p->zz, r.zz); st_fq(&p->zzz, r.zzz);
}

}  // namespace zkp
