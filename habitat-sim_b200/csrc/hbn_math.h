// Geometry primitives of the query path, restated for host+device with the exact
// operation order of Detour/Include/DetourCommon.h and Detour/Source/DetourCommon.cpp
// (cited per function).  Float control flow must be bit-identical to the reference, so:
// no FMA contraction (nvcc -fmad=false, gcc -ffp-contract=off), IEEE div/sqrt
// (-prec-div=true -prec-sqrt=true are nvcc defaults), no fast-math.
#pragma once
#include <math.h>
#include "hbn_types.h"

namespace hbn {

constexpr float kFltMax = 3.402823466e+38f;

HBN_HD float fsqrt(float x) { return sqrtf(x); }
HBN_HD bool finitef(float x) { return fabsf(x) <= kFltMax; }  // false for NaN and +-inf
HBN_HD bool vfinite(const float* v) { return finitef(v[0]) && finitef(v[1]) && finitef(v[2]); }
HBN_HD float fclamp(float v, float mn, float mx) { return v < mn ? mn : (v > mx ? mx : v); }  // dtClamp
HBN_HD float sqr(float a) { return a * a; }

// DetourCommon.h:117-122
HBN_HD void vlerp(float* d, const float* a, const float* b, float t) {
  d[0] = a[0] + (b[0] - a[0]) * t;
  d[1] = a[1] + (b[1] - a[1]) * t;
  d[2] = a[2] + (b[2] - a[2]) * t;
}
HBN_HD void vcopy(float* d, const float* a) { d[0] = a[0]; d[1] = a[1]; d[2] = a[2]; }
// DetourCommon.h:217-223
HBN_HD float vdist(const float* a, const float* b) {
  const float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
  return fsqrt(dx * dx + dy * dy + dz * dz);
}
// DetourCommon.h:229-235
HBN_HD float vdistSqr(const float* a, const float* b) {
  const float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
  return dx * dx + dy * dy + dz * dz;
}
// DetourCommon.h:254-259
HBN_HD float vdist2DSqr(const float* a, const float* b) {
  const float dx = b[0] - a[0], dz = b[2] - a[2];
  return dx * dx + dz * dz;
}
// DetourCommon.h:278-283
HBN_HD bool vequal(const float* a, const float* b) {
  const float thr = (1.0f / 16384.0f) * (1.0f / 16384.0f);
  return vdistSqr(a, b) < thr;
}
// DetourCommon.h:338-345
HBN_HD float triArea2D(const float* a, const float* b, const float* c) {
  const float abx = b[0] - a[0], abz = b[2] - a[2];
  const float acx = c[0] - a[0], acz = c[2] - a[2];
  return acx * abz - abx * acz;
}
// DetourCommon.cpp:170-184
HBN_HD float distPtSegSqr2D(const float* pt, const float* p, const float* q, float& t) {
  float pqx = q[0] - p[0];
  float pqz = q[2] - p[2];
  float dx = pt[0] - p[0];
  float dz = pt[2] - p[2];
  float d = pqx * pqx + pqz * pqz;
  t = pqx * dx + pqz * dz;
  if (d > 0) t /= d;
  if (t < 0) t = 0;
  else if (t > 1) t = 1;
  dx = p[0] + t * pqx - pt[0];
  dz = p[2] + t * pqz - pt[2];
  return dx * dx + dz * dz;
}
// DetourCommon.cpp:204-233
HBN_HD bool closestHeightPointTriangle(const float* p, const float* a, const float* b,
                                       const float* c, float& h) {
  const float EPS = 1e-6f;
  const float v0x = c[0] - a[0], v0y = c[1] - a[1], v0z = c[2] - a[2];
  const float v1x = b[0] - a[0], v1y = b[1] - a[1], v1z = b[2] - a[2];
  const float v2x = p[0] - a[0], v2z = p[2] - a[2];
  float denom = v0x * v1z - v0z * v1x;
  if (fabsf(denom) < EPS) return false;
  float u = v1z * v2x - v1x * v2z;
  float v = v0x * v2z - v0z * v2x;
  if (denom < 0) {
    denom = -denom;
    u = -u;
    v = -v;
  }
  if (u >= 0.0f && v >= 0.0f && (u + v) <= denom) {
    h = a[1] + (v0y * u + v1y * v) / denom;
    return true;
  }
  return false;
}
// DetourCommon.cpp:238-252 (verts: nv consecutive xyz triples)
HBN_HD bool pointInPolygon(const float* pt, const float* verts, int nv) {
  bool c = false;
  for (int i = 0, j = nv - 1; i < nv; j = i++) {
    const float* vi = &verts[i * 3];
    const float* vj = &verts[j * 3];
    if (((vi[2] > pt[2]) != (vj[2] > pt[2])) &&
        (pt[0] < (vj[0] - vi[0]) * (pt[2] - vi[2]) / (vj[2] - vi[2]) + vi[0]))
      c = !c;
  }
  return c;
}
// DetourCommon.cpp:373-386
HBN_HD bool intersectSegSeg2D(const float* ap, const float* aq, const float* bp,
                              const float* bq, float& s, float& t) {
  const float ux = aq[0] - ap[0], uz = aq[2] - ap[2];
  const float vx = bq[0] - bp[0], vz = bq[2] - bp[2];
  const float wx = ap[0] - bp[0], wz = ap[2] - bp[2];
  const float d = ux * vz - uz * vx;
  if (fabsf(d) < 1e-6f) return false;
  s = (vx * wz - vz * wx) / d;
  t = (ux * wz - uz * wx) / d;
  return true;
}
// dtOverlapQuantBounds, DetourCommon.h:354-362
HBN_HD bool overlapQuant(const uint16_t* amin, const uint16_t* amax, const uint16_t* bmin,
                         const uint16_t* bmax) {
  bool overlap = true;
  overlap = (amin[0] > bmax[0] || amax[0] < bmin[0]) ? false : overlap;
  overlap = (amin[1] > bmax[1] || amax[1] < bmin[1]) ? false : overlap;
  overlap = (amin[2] > bmax[2] || amax[2] < bmin[2]) ? false : overlap;
  return overlap;
}

// Magnum TypeTraits<Float>::equals (src/deps/magnum/src/Magnum/Math/TypeTraits.h:495-510,
// epsilon 1e-5f :531): the fuzzy compare behind `pathStart == pathEnd`,
// PathFinder.cpp:1434.
HBN_HD bool fuzzyEq(float a, float b) {
  if (a == b) return true;
  const float absA = fabsf(a), absB = fabsf(b), diff = fabsf(a - b);
  const float eps = 1.0e-5f;
  if (a == 0.0f || b == 0.0f || diff < eps) return diff < eps;
  return diff / (absA + absB) < eps;
}
HBN_HD bool vfuzzyEq(const float* a, const float* b) {
  return fuzzyEq(a[0], b[0]) && fuzzyEq(a[1], b[1]) && fuzzyEq(a[2], b[2]);
}
// Magnum Vector3::length() of (a-b): dot accumulates from 0 left to right (Vector.h:106-111)
HBN_HD float mnDist(const float* a, const float* b) {
  const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  float out = 0.0f;
  out += dx * dx;
  out += dy * dy;
  out += dz * dz;
  return fsqrt(out);
}

// Counter-based uniform stream in [0,1], in the form of the reference's frand() (PF.cpp:1232-1234):
// float(r) / float(RAND_MAX) with r a 31-bit value, so that the reference's own code, fed r through
// rand(), computes the identical float (1.0 is reachable: float(2^31 - 1) rounds to 2^31; trap T7).
// This is this library's definition (include/hbn.h: hbn_uniform).
HBN_HD uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
HBN_HD float uniform01(uint64_t seed, uint64_t query, uint32_t draw) {
  uint32_t h = mix32(static_cast<uint32_t>(seed) ^ 0x9e3779b9U);
  h = mix32(h ^ static_cast<uint32_t>(seed >> 32));
  h = mix32(h ^ static_cast<uint32_t>(query));
  h = mix32(h ^ static_cast<uint32_t>(query >> 32) ^ 0x85ebca6bU);
  h = mix32(h ^ draw);
  return static_cast<float>(static_cast<int>(h >> 1)) / 2147483648.0f;
}

}  // namespace hbn
