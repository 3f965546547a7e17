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
// Record `idx` of a point array whose records are `stride16` 16-byte words apart: 6 = packed G1Affine (96 B), 8 = the
// fixed-base table's 128-byte records (x || y || 32 B pad), where a gathered point is exactly one 128-byte DRAM line.
__device__ __forceinline__ G1Affine ld_affine_s(const G1Affine* base, size_t idx, uint32_t stride16) {
  const uint4* q = reinterpret_cast<const uint4*>(base) + idx * stride16;
  const uint4 a = q[0], b = q[1], c = q[2], d = q[3], e = q[4], f = q[5];
  G1Affine r;
  r.x.v[0] = a.x; r.x.v[1] = a.y; r.x.v[2] = a.z; r.x.v[3] = a.w;
  r.x.v[4] = b.x; r.x.v[5] = b.y; r.x.v[6] = b.z; r.x.v[7] = b.w;
  r.x.v[8] = c.x; r.x.v[9] = c.y; r.x.v[10] = c.z; r.x.v[11] = c.w;
  r.y.v[0] = d.x; r.y.v[1] = d.y; r.y.v[2] = d.z; r.y.v[3] = d.w;
  r.y.v[4] = e.x; r.y.v[5] = e.y; r.y.v[6] = e.z; r.y.v[7] = e.w;
  r.y.v[8] = f.x; r.y.v[9] = f.y; r.y.v[10] = f.z; r.y.v[11] = f.w;
  return r;
}
__device__ __forceinline__ Fq ld_affine_x_s(const G1Affine* base, size_t idx, uint32_t stride16) {
  return ld_fq(reinterpret_cast<const Fq*>(reinterpret_cast<const uint4*>(base) + idx * stride16));
}
__device__ __forceinline__ void st_affine_s(G1Affine* base, size_t idx, uint32_t stride16, const G1Affine& r) {
  Fq* q = reinterpret_cast<Fq*>(reinterpret_cast<uint4*>(base) + idx * stride16);
  st_fq(q, r.x);
  st_fq(q + 1, r.y);
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
