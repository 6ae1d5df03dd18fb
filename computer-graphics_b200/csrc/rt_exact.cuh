// Reference-order ("exact") device arithmetic of the raytracer hot path.
//
// Every function here reproduces, operation for operation and rounding for
// rounding, what the reference computes on the CPU (no FMA contraction, IEEE
// div/sqrt, the same double "islands"), so that results are bit-identical.
// Citations are file:line in fznsakib/Computer-Graphics.
#pragma once
#include "common.cuh"
#include <float.h>

struct RtHit {
  float t;      // ray parameter of the accepted hit (position = start + t*dir)
  float dist;   // closestIntersection.distance (t*|dir| for triangles, raw t for spheres)
  int idx;      // triangle index, or -1 - sphereIndex; only valid when dist < FLT_MAX
};

// Per-triangle geometry the exact test needs: v0 and the two edges
// e1 = v1 - v0, e2 = v2 - v0 (raytracer/Source/skeleton.cpp:283-284).  The edges
// depend on the triangle only, so they are computed once (same float ops).
struct RtGeom {
  float v0x, v0y, v0z, e1x;
  float e1y, e1z, e2x, e2y;
  float e2z, pad0, pad1, pad2;
};

// One triangle of ClosestIntersection's loop (skeleton.cpp:278-336).
// `len` is glm::length(dir3) of this ray (skeleton.cpp:307).
__device__ __forceinline__ void rt_exact_triangle(float sx0, float sy0, float sz0, float dx, float dy,
                                                  float dz, float len, const float4 g0,
                                                  const float4 g1, const float4 g2, int i,
                                                  RtHit &best) {
  const float v0x = g0.x, v0y = g0.y, v0z = g0.z, e1x = g0.w;
  const float e1y = g1.x, e1z = g1.y, e2x = g1.z, e2y = g1.w;
  const float e2z = g2.x;
  const float sx = xsub(sx0, v0x), sy = xsub(sy0, v0y), sz = xsub(sz0, v0z);  // :296-297
  const float D = xdet3(-dx, -dy, -dz, e1x, e1y, e1z, e2x, e2y, e2z);        // :289
  const float t = xdiv(xdet3(sx, sy, sz, e1x, e1y, e1z, e2x, e2y, e2z), D);   // :305-306
  const float distance = xmul(t, len);                                        // :307
  if (distance < 0.0f) return;                                                // :311
  if (distance >= best.dist || distance > FLT_MAX) return;                    // :313
  const float u = xdiv(xdet3(-dx, -dy, -dz, sx, sy, sz, e2x, e2y, e2z), D);   // :317-318
  const float v = xdiv(xdet3(-dx, -dy, -dz, e1x, e1y, e1z, sx, sy, sz), D);   // :320-321
  if ((u >= 0) && (v >= 0) && (xadd(u, v) <= 1)) {                            // :328
    best.t = t;
    best.dist = distance;
    best.idx = i;
  }
}

// Sphere::solveQuadratic + Sphere::intersect (raytracer/Source/TestModelH.h:24-66).
__device__ __forceinline__ bool rt_exact_sphere(const rt_sphere &s, float sx, float sy, float sz,
                                                float dx, float dy, float dz, float &t_out) {
  const float Lx = xsub(sx, s.centre[0]), Ly = xsub(sy, s.centre[1]), Lz = xsub(sz, s.centre[2]);
  const float a = xdot3(dx, dy, dz, dx, dy, dz);
  const float b = xmul(2.0f, xdot3(dx, dy, dz, Lx, Ly, Lz));
  const float c = xsub(xdot3(Lx, Ly, Lz, Lx, Ly, Lz), s.radius_squared);
  const float disc = xsub(xmul(b, b), xmul(xmul(4.0f, a), c));
  float x0, x1;
  if (disc < 0) return false;
  else if (disc == 0) {
    x0 = x1 = __double2float_rn(__ddiv_rn(__dmul_rn(-0.5, (double)b), (double)a));
  } else {
    float q;
    if (b > 0) q = xmul(-0.5f, xadd(b, xsqrt(disc)));  // x -0.5 is exact in any precision
    else q = xmul(-0.5f, xsub(b, xsqrt(disc)));
    x0 = xdiv(q, a);
    x1 = xdiv(c, q);
  }
  if (x0 > x1) { float tmp = x0; x0 = x1; x1 = tmp; }
  if (x0 < 0) {
    x0 = x1;
    if (x0 < 0) return false;
  }
  t_out = x0;
  return true;
}

// The sphere loop of ClosestIntersection (skeleton.cpp:341-355).
__device__ __forceinline__ void rt_exact_spheres(const rt_sphere *__restrict__ sph, int n_sph, float sx,
                                                 float sy, float sz, float dx, float dy, float dz,
                                                 RtHit &best) {
  for (int i = 0; i < n_sph; ++i) {
    float t;
    if (rt_exact_sphere(sph[i], sx, sy, sz, dx, dy, dz, t)) {
      if (t < best.dist) {
        best.t = t;
        best.dist = t;
        best.idx = -1 - i;
      }
    }
  }
}

// Lighting tail of DirectLight once the shadow ray is known to be clear
// (skeleton.cpp:400-412).  r = light.pos - hit.pos (xyz), n = surface normal.
__device__ __forceinline__ void rt_exact_lambert(float rx, float ry, float rz, float r_mag, float nx,
                                                 float ny, float nz, const float *col,
                                                 const float *lcol, float *power) {
  const float inv = xdiv(1.0f, xsqrt(xdot3(rx, ry, rz, rx, ry, rz)));  // glm::normalize
  float a = xdot3(xmul(rx, inv), xmul(ry, inv), xmul(rz, inv), nx, ny, nz);
  const float b = (float)(4 * 3.14159265358979323846);  // float b = 4 * M_PI  (:404)
  const float area = __double2float_rn(__dmul_rn((double)b, __dmul_rn((double)r_mag, (double)r_mag)));
  if (a <= 0) a = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) power[k] = xdiv(xmul(xmul(col[k], lcol[k]), a), area);
}

// r_magnitude of DirectLight (skeleton.cpp:371): double sum of squares, double
// sqrt, narrowed to float.
__device__ __forceinline__ float rt_exact_rmag(float rx, float ry, float rz) {
  const double s = __dadd_rn(__dadd_rn(__dmul_rn((double)rx, (double)rx), __dmul_rn((double)ry, (double)ry)),
                             __dmul_rn((double)rz, (double)rz));
  return __double2float_rn(__dsqrt_rn(s));
}
