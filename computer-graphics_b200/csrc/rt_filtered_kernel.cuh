// rt_filtered_kernel: see rt_filtered.cuh for the method.  The long, rarely
// executed reference-arithmetic pieces are __noinline__ so that the hot filter
// loops stay small enough for the instruction cache.
#pragma once

// Colour and shading normal of a hit (skeleton.cpp:148-149, 377-387;
// Sphere::getNormal TestModelH.h:68-75).
__device__ __forceinline__ void rt_surface2(const rt_triangle *__restrict__ src, const rt_sphere *__restrict__ sph,
                                            int idx, float px, float py, float pz, float *col, float &nx,
                                            float &ny, float &nz) {
  if (idx >= 0) {
    const rt_triangle *t = src + idx;
    col[0] = __ldg(&t->color[0]); col[1] = __ldg(&t->color[1]); col[2] = __ldg(&t->color[2]);
    nx = __ldg(&t->normal[0]); ny = __ldg(&t->normal[1]); nz = __ldg(&t->normal[2]);
  } else {
    const rt_sphere *s = sph + (-1 - idx);
    col[0] = s->color[0]; col[1] = s->color[1]; col[2] = s->color[2];
    const float ax = xsub(px, s->centre[0]), ay = xsub(py, s->centre[1]), az = xsub(pz, s->centre[2]);
    const float inv = xdiv(1.0f, xsqrt(xdot3(ax, ay, az, ax, ay, az)));
    nx = xmul(ax, inv); ny = xmul(ay, inv); nz = xmul(az, inv);
  }
}

// Reference arithmetic for one (primary ray, triangle) pair; `inside` = the
// filter proved `check` (:328) true, so u and v need not be formed.
__device__ __noinline__ RtHit rt_ex_primary(const float4 *__restrict__ geom, const float *__restrict__ dt_cam,
                                            float cx, float cy, float cz, int tri, int inside, float dx, float dy,
                                            float dz, float len, RtHit best) {
  const float4 g0 = __ldg(geom + 3 * tri), g1 = __ldg(geom + 3 * tri + 1), g2 = __ldg(geom + 3 * tri + 2);
  const float e1x = g0.w, e1y = g1.x, e1z = g1.y, e2x = g1.z, e2y = g1.w, e2z = g2.x;
  const float D = xdet3(-dx, -dy, -dz, e1x, e1y, e1z, e2x, e2y, e2z);
  const float t = xdiv(__ldg(dt_cam + tri), D);   // numerator: the same for every primary ray (rt_prep_planes_kernel)
  const float distance = xmul(t, len);
  if (distance < 0.0f) return best;
  if (distance >= best.dist || distance > FLT_MAX) return best;
  if (!inside) {
    const float sx = xsub(cx, g0.x), sy = xsub(cy, g0.y), sz = xsub(cz, g0.z);
    const float u = xdiv(xdet3(-dx, -dy, -dz, sx, sy, sz, e2x, e2y, e2z), D);
    const float v = xdiv(xdet3(-dx, -dy, -dz, e1x, e1y, e1z, sx, sy, sz), D);
    if (!((u >= 0) && (v >= 0) && (xadd(u, v) <= 1))) return best;
  }
  best.t = t; best.dist = distance; best.idx = tri;
  return best;
}

// Reference arithmetic for one (shadow ray, triangle) pair: does this triangle
// make DirectLight return black (skeleton.cpp:394-397)?  Any accepted hit with
// distance < r_magnitude does, because ClosestIntersection keeps the minimum.
__device__ __noinline__ int rt_ex_shadow(const float4 *__restrict__ geom, const rt_triangle *__restrict__ src,
                                         const rt_sphere *__restrict__ sph, int tri, int hit_idx, float px,
                                         float py, float pz, float Lx, float Ly, float Lz) {
  float col[3], nx, ny, nz;
  rt_surface2(src, sph, hit_idx, px, py, pz, col, nx, ny, nz);
  const float rx = xsub(Lx, px), ry = xsub(Ly, py), rz = xsub(Lz, pz);                          // :370
  const float ox = xadd(px, xmul(nx, 0.00001f)), oy = xadd(py, xmul(ny, 0.00001f)),
              oz = xadd(pz, xmul(nz, 0.00001f));                                                 // :394
  const float4 g0 = __ldg(geom + 3 * tri), g1 = __ldg(geom + 3 * tri + 1), g2 = __ldg(geom + 3 * tri + 2);
  const float e1x = g0.w, e1y = g1.x, e1z = g1.y, e2x = g1.z, e2y = g1.w, e2z = g2.x;
  const float sx = xsub(ox, g0.x), sy = xsub(oy, g0.y), sz = xsub(oz, g0.z);
  const float D = xdet3(-rx, -ry, -rz, e1x, e1y, e1z, e2x, e2y, e2z);
  const float Dt = xdet3(sx, sy, sz, e1x, e1y, e1z, e2x, e2y, e2z);
  const float rr = xdot3(rx, ry, rz, rx, ry, rz);
  // The usual outcome for the triangle the ray starts on: t = Dt / D is negative, so
  // `distance < 0` (:311) rejects it.  The sign of an IEEE quotient is the sign of
  // its operands and neither t nor t * len can underflow to -0 under the guards
  // below, so the division, the length and the double-precision r_magnitude are
  // not needed to know that.
  if (((Dt < 0.0f) != (D < 0.0f)) && fabsf(Dt) > 1e-20f * fabsf(D) && fabsf(D) < 1e30f && rr > 1e-30f && rr < 1e30f)
    return 0;
  const float len = xsqrt(rr);
  const float r_mag = rt_exact_rmag(rx, ry, rz);
  const float t = xdiv(Dt, D);
  const float distance = xmul(t, len);
  if (distance < 0.0f) return 0;
  if (!(distance < r_mag) || distance > FLT_MAX) return 0;
  const float u = xdiv(xdet3(-rx, -ry, -rz, sx, sy, sz, e2x, e2y, e2z), D);
  const float v = xdiv(xdet3(-rx, -ry, -rz, e1x, e1y, e1z, sx, sy, sz), D);
  return ((u >= 0) && (v >= 0) && (xadd(u, v) <= 1)) ? 1 : 0;
}

// Conservative sphere pre-test: the reference reports no root when
// discriminant = b*b - 4*a*c < 0 (TestModelH.h:27-28).  disc = 4 q with
// q = (d.L)^2 - (d.d)(L.L - r^2); reject when q is negative beyond the rounding
// error of either evaluation (|error| < 2e-6 (d.d)(L.L + r^2), tolerance 1e-4).
__device__ __forceinline__ bool rt_sphere_may_hit(const rt_sphere &s, float sx, float sy, float sz, float dx,
                                                  float dy, float dz) {
  const float Lx = sx - s.centre[0], Ly = sy - s.centre[1], Lz = sz - s.centre[2];
  const float dL = fmaf(dx, Lx, fmaf(dy, Ly, dz * Lz));
  const float dd = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
  const float LL = fmaf(Lx, Lx, fmaf(Ly, Ly, Lz * Lz));
  const float q = fmaf(dL, dL, -dd * (LL - s.radius_squared));
  const float tol = 1e-4f * fmaf(dL, dL, dd * (LL + s.radius_squared));
  return !(q < -tol);
}

struct RtSphT { int hit; float t; };
__device__ __noinline__ RtSphT rt_ex_sphere(const rt_sphere *__restrict__ s, float sx, float sy, float sz, float dx,
                                            float dy, float dz) {
  RtSphT r;
  r.t = 0.f;
  r.hit = rt_exact_sphere(*s, sx, sy, sz, dx, dy, dz, r.t) ? 1 : 0;
  return r;
}

// DirectLight for one hit once the triangles' verdict (`occluded`) is known:
// sphere occluders, then the Lambert term (skeleton.cpp:366-415).  Also returns
// the surface colour for the indirect term (:156).
struct RtShade { float pw[3]; float col[3]; };
__device__ __noinline__ RtShade rt_ex_shade(const rt_triangle *__restrict__ src, const rt_sphere *__restrict__ sph,
                                            int n_sph, int hit_idx, float px, float py, float pz, float Lx,
                                            float Ly, float Lz, float lr, float lg, float lb, int occluded,
                                            int want_light) {
  RtShade o;
  float nx, ny, nz;
  rt_surface2(src, sph, hit_idx, px, py, pz, o.col, nx, ny, nz);
  o.pw[0] = o.pw[1] = o.pw[2] = 0.f;
  if (!want_light || occluded) return o;
  const float rx = xsub(Lx, px), ry = xsub(Ly, py), rz = xsub(Lz, pz);
  const float r_mag = rt_exact_rmag(rx, ry, rz);
  if (n_sph > 0) {
    const float ox = xadd(px, xmul(nx, 0.00001f)), oy = xadd(py, xmul(ny, 0.00001f)),
                oz = xadd(pz, xmul(nz, 0.00001f));
    for (int s = 0; s < n_sph; ++s) {
      if (!rt_sphere_may_hit(sph[s], ox, oy, oz, rx, ry, rz)) continue;
      float t;
      if (rt_exact_sphere(sph[s], ox, oy, oz, rx, ry, rz, t) && t < r_mag) return o;   // :348 + :395
    }
  }
  const float lcol[3] = {lr, lg, lb};
  rt_exact_lambert(rx, ry, rz, r_mag, nx, ny, nz, o.col, lcol, o.pw);
  return o;
}

// The TMA tile ring shared by the primary and the shadow passes.
struct RtRing {
  float4 *smem;
  uint64_t *bars;
  uint32_t phase_bits;
  int issued;
};

__device__ __forceinline__ void rt_ring_issue(RtRing &ring, const float4 *src, int cnt) {
  const int buf = ring.issued & 1;
  mbar_expect_tx(&ring.bars[buf], cnt * 48u);
  tma_bulk_g2s(ring.smem + (size_t)buf * RT_TILE * RT_REC_F4, src, cnt * 48u, &ring.bars[buf]);
}

template <bool MULTI>
#ifndef RT_MIN_BLOCKS
#define RT_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(RT_THREADS, RT_MIN_BLOCKS) rt_filtered_kernel(const __grid_constant__ RtKParams p) {
  extern __shared__ __align__(128) float4 tile_smem[];  // 2 x RT_TILE x 3 float4
  __shared__ __align__(8) uint64_t bars[2];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp = 8x4 pixel patch; 8 warps tile a 16x16 block as 2 columns x 4 rows
  const int u = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
  const int v = p.row0 + blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
  const bool live = (u < p.W) && (v < p.row1);
  const size_t pid = (size_t)v * p.W + u;
  const int n_tiles = (p.n_tris + RT_TILE - 1) / RT_TILE;
  const size_t origin_stride = (size_t)n_tiles * RT_TILE * RT_REC_F4;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  RtRing ring;
  ring.smem = tile_smem; ring.bars = bars; ring.phase_bits = 0; ring.issued = 0;

  // dir = R * vec4(u - W/2, v - H/2, f, 1)   (skeleton.cpp:126-128)
  const float x = (float)(u - p.W / 2), y = (float)(v - p.H / 2);
  const float dir0 = xadd(xadd(xmul(p.R[0], x), xmul(p.R[4], y)), xadd(xmul(p.R[8], p.focal), xmul(p.R[12], 1.0f)));
  const float dir1 = xadd(xadd(xmul(p.R[1], x), xmul(p.R[5], y)), xadd(xmul(p.R[9], p.focal), xmul(p.R[13], 1.0f)));
  const float dz = p.focal;
  const float cx = p.cam[0], cy = p.cam[1], cz = p.cam[2];
  // the nine sample directions (:134-137): three x offsets, three y offsets
  const float dxs[3] = {xadd(dir0, xmul(0.5f, -1.0f)), xadd(dir0, xmul(0.5f, 0.0f)), xadd(dir0, xmul(0.5f, 1.0f))};
  const float dys[3] = {xadd(dir1, xmul(0.5f, -1.0f)), xadd(dir1, xmul(0.5f, 0.0f)), xadd(dir1, xmul(0.5f, 1.0f))};

  RtHit best[9];
  float len[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    best[k].t = 0.f; best[k].dist = FLT_MAX; best[k].idx = 0;
    len[k] = xsqrt(xdot3(dxs[k / 3], dys[k % 3], dz, dxs[k / 3], dys[k % 3], dz));
  }
  unsigned n_exact = 0;
#ifdef RT_PROFILE_COUNTERS
  unsigned prof_cnt[7] = {0, 0, 0, 0, 0, 0, 0};
#define RT_PC(i) (++prof_cnt[i])
#else
#define RT_PC(i)
#endif

  // ============================ primary rays ============================
  {
    // bounding box of the warp's direction bundle in the (dx, dy) plane
    const float lo0 = warp_min(live ? dir0 : INFINITY), hi0 = warp_max(live ? dir0 : -INFINITY);
    const float lo1 = warp_min(live ? dir1 : INFINITY), hi1 = warp_max(live ? dir1 : -INFINITY);
    const bool warp_live = hi0 >= lo0;
    const float wc0 = 0.5f * (lo0 + hi0), wc1 = 0.5f * (lo1 + hi1);
    const float wh0 = (0.5f * (hi0 - lo0) + 0.5f) * 1.0001f + 1e-3f;
    const float wh1 = (0.5f * (hi1 - lo1) + 0.5f) * 1.0001f + 1e-3f;

    const float4 *src = p.planes;  // origin 0
    if (threadIdx.x == 0 && n_tiles > 0) rt_ring_issue(ring, src, min(RT_TILE, p.n_tris));
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int buf = ring.issued & 1;
      ++ring.issued;
      __syncthreads();   // everyone is done with the other buffer (it held tile-1): refill it
      if (threadIdx.x == 0 && tile + 1 < n_tiles)
        rt_ring_issue(ring, src + (size_t)(tile + 1) * RT_TILE * RT_REC_F4, min(RT_TILE, p.n_tris - (tile + 1) * RT_TILE));
      mbar_wait(&bars[buf], (ring.phase_bits >> buf) & 1u);
      ring.phase_bits ^= 1u << buf;
      const float4 *T = tile_smem + (size_t)buf * RT_TILE * RT_REC_F4;
      const int base = tile * RT_TILE;
      const int cnt = min(RT_TILE, p.n_tris - base);
      if (!warp_live) continue;
      for (int r0 = 0; r0 < cnt; r0 += 32) {
        // ---- L0: one triangle per lane vs the warp's bundle box ----
        bool pass = false;
        if (r0 + lane < cnt) {
          const float4 q0 = T[(r0 + lane) * 3], q1 = T[(r0 + lane) * 3 + 1], q2 = T[(r0 + lane) * 3 + 2];
          const float bu = fmaf(q0.x, wc0, fmaf(q0.y, wc1, q0.z)) + fmaf(fabsf(q0.x), wh0, fabsf(q0.y) * wh1);
          const float bv = fmaf(q0.w, wc0, fmaf(q1.x, wc1, q1.y)) + fmaf(fabsf(q0.w), wh0, fabsf(q1.x) * wh1);
          const float bw = fmaf(q1.z, wc0, fmaf(q1.w, wc1, q2.x)) + fmaf(fabsf(q1.z), wh0, fabsf(q1.w) * wh1);
          // E already covers the rounding of the centre evaluation (|wc| <= dmax);
          // the half-width terms are sums of non-negative products, inflated above
          pass = !(fminf(fminf(bu, bv), bw) < -q2.y);
        }
        unsigned mask = __ballot_sync(0xffffffffu, pass);
        while (mask) {
          const int j = __ffs(mask) - 1;
          mask &= mask - 1;
          const int tri = base + r0 + j;
          const float4 q0 = T[(r0 + j) * 3], q1 = T[(r0 + j) * 3 + 1], q2 = T[(r0 + j) * 3 + 2];
          if (!live) continue;
          // ---- L1: pixel centre, margin widened by the +-0.5 jitter ----
          const float cu = fmaf(q0.x, dir0, fmaf(q0.y, dir1, q0.z));
          const float cv = fmaf(q0.w, dir0, fmaf(q1.x, dir1, q1.y));
          const float cw = fmaf(q1.z, dir0, fmaf(q1.w, dir1, q2.x));
          RT_PC(0);
          if (fminf(fminf(cu, cv), cw) < -q2.w) continue;
          RT_PC(1);
          const float E = q2.y, dt_lo = q2.z;
          // ---- L2 / EX: per ray ----
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const float dx = dxs[k / 3], dy = dys[k % 3];
            const float mU = fmaf(q0.x, dx, fmaf(q0.y, dy, q0.z));
            const float mV = fmaf(q0.w, dx, fmaf(q1.x, dy, q1.y));
            const float mW = fmaf(q1.z, dx, fmaf(q1.w, dy, q2.x));
            const float m3 = fminf(fminf(mU, mV), mW);
            if (m3 >= -E) {
              RT_PC(2);
              const float mN = mU + mV + mW;
              // reference distance >= dt_lo*len/(mN+E): cannot beat the current closest
              const bool farther = (mN > E) && (dt_lo * len[k] >= best[k].dist * (mN + E));
              if (!farther) {
                ++n_exact;
                best[k] = rt_ex_primary(p.geom, p.dt_cam, cx, cy, cz, tri, m3 >= E ? 1 : 0, dx, dy, dz, len[k], best[k]);
              }
            }
          }
        }
      }
    }
  }

  // spheres (skeleton.cpp:341-355)
  if (live) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float dx = dxs[k / 3], dy = dys[k % 3];
      for (int s = 0; s < p.n_sph; ++s) {
        if (!rt_sphere_may_hit(p.sph[s], cx, cy, cz, dx, dy, dz)) continue;
        const RtSphT h = rt_ex_sphere(p.sph + s, cx, cy, cz, dx, dy, dz);
        if (h.hit && h.t < best[k].dist) { best[k].t = h.t; best[k].dist = h.t; best[k].idx = -1 - s; }
      }
    }
  }
  unsigned active = 0;
  float ht[9];   // ray parameter of each hit; position = start + t * dir (:326 / :345)
  int idx[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (live && best[k].dist < FLT_MAX) active |= 1u << k;
    ht[k] = best[k].t;
    idx[k] = best[k].idx;
  }
  if (live) {
    const bool hit4 = (active >> 4) & 1u;
    if (p.depth) p.depth[pid] = hit4 ? best[4].dist : INFINITY;
    if (p.index) p.index[pid] = hit4 ? best[4].idx : INT32_MIN;
  }

  // ============================ shadow rays ============================
  constexpr int NL = MULTI ? B200_MAX_LIGHTS : 1;
  float dl[MULTI ? NL * 27 : 1];
  float pix[3] = {0.f, 0.f, 0.f};
  const unsigned warp_active = __ballot_sync(0xffffffffu, active != 0);
  for (int l = 0; l < p.n_lights; ++l) {
    const float Lx = p.lights[l][0], Ly = p.lights[l][1], Lz = p.lights[l][2];
    unsigned occluded = 0;
    // g = hit - light per ray; boxes of the pixel's and the warp's bundles
    float glo[3] = {INFINITY, INFINITY, INFINITY}, ghi[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if ((active >> k) & 1u) {
        const float gx = -xsub(Lx, xadd(cx, xmul(ht[k], dxs[k / 3])));
        const float gy = -xsub(Ly, xadd(cy, xmul(ht[k], dys[k % 3])));
        const float gz = -xsub(Lz, xadd(cz, xmul(ht[k], dz)));
        glo[0] = fminf(glo[0], gx); ghi[0] = fmaxf(ghi[0], gx);
        glo[1] = fminf(glo[1], gy); ghi[1] = fmaxf(ghi[1], gy);
        glo[2] = fminf(glo[2], gz); ghi[2] = fmaxf(ghi[2], gz);
      }
    }
    float pc[3], ph[3], wc[3], wh[3];
    float pnorm = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      pc[c] = 0.5f * (glo[c] + ghi[c]);
      ph[c] = 0.5f * (ghi[c] - glo[c]) * 1.0001f + 1e-7f * fmaxf(fabsf(glo[c]), fabsf(ghi[c]));
      pnorm = fmaxf(pnorm, fmaxf(fabsf(glo[c]), fabsf(ghi[c])));
      const float wl = warp_min(glo[c]), wu = warp_max(ghi[c]);
      wc[c] = 0.5f * (wl + wu);
      wh[c] = 0.5f * (wu - wl) * 1.0001f + 1e-7f * fmaxf(fabsf(wl), fabsf(wu));
    }
    if (!active) { pnorm = 0.f; pc[0] = pc[1] = pc[2] = 0.f; ph[0] = ph[1] = ph[2] = 0.f; }
    const float wnorm = warp_max(pnorm) * 1.0001f;
    pnorm *= 1.0001f;

    const float4 *src = p.planes + (size_t)(1 + l) * origin_stride;
    // the buffer about to be refilled last held the tile before the previous
    // phase's final one; every thread left it at that phase's last barrier
    __syncthreads();
    if (threadIdx.x == 0 && n_tiles > 0) rt_ring_issue(ring, src, min(RT_TILE, p.n_tris));
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int buf = ring.issued & 1;
      ++ring.issued;
      __syncthreads();
      if (threadIdx.x == 0 && tile + 1 < n_tiles)
        rt_ring_issue(ring, src + (size_t)(tile + 1) * RT_TILE * RT_REC_F4, min(RT_TILE, p.n_tris - (tile + 1) * RT_TILE));
      mbar_wait(&bars[buf], (ring.phase_bits >> buf) & 1u);
      ring.phase_bits ^= 1u << buf;
      const float4 *T = tile_smem + (size_t)buf * RT_TILE * RT_REC_F4;
      const int base = tile * RT_TILE;
      const int cnt = min(RT_TILE, p.n_tris - base);
      if (!warp_active) continue;
      for (int r0 = 0; r0 < cnt; r0 += 32) {
        // ---- L0: one triangle per lane vs the warp's shadow-bundle box ----
        bool pass = false;
        if (r0 + lane < cnt) {
          const float4 q0 = T[(r0 + lane) * 3], q1 = T[(r0 + lane) * 3 + 1], q2 = T[(r0 + lane) * 3 + 2];
          const float Eg = q2.y * wnorm;
          const float cu = fmaf(q0.x, wc[0], fmaf(q0.y, wc[1], q0.z * wc[2]));
          const float cv = fmaf(q0.w, wc[0], fmaf(q1.x, wc[1], q1.y * wc[2]));
          const float cw = fmaf(q1.z, wc[0], fmaf(q1.w, wc[1], q2.x * wc[2]));
          const float hu = fmaf(fabsf(q0.x), wh[0], fmaf(fabsf(q0.y), wh[1], fabsf(q0.z) * wh[2]));
          const float hv = fmaf(fabsf(q0.w), wh[0], fmaf(fabsf(q1.x), wh[1], fabsf(q1.y) * wh[2]));
          const float hw = fmaf(fabsf(q1.z), wh[0], fmaf(fabsf(q1.w), wh[1], fabsf(q2.x) * wh[2]));
          const float m3 = fminf(fminf(cu + hu, cv + hv), cw + hw);
          const float mN = (cu + cv + cw) + (hu + hv + hw);
          pass = !(m3 < -Eg) && !(mN + Eg < q2.z);
        }
        unsigned mask = __ballot_sync(0xffffffffu, pass);
        while (mask) {
          const int j = __ffs(mask) - 1;
          mask &= mask - 1;
          const int tri = base + r0 + j;
          const float4 q0 = T[(r0 + j) * 3], q1 = T[(r0 + j) * 3 + 1], q2 = T[(r0 + j) * 3 + 2];
          const unsigned todo = active & ~occluded;
          if (!todo) continue;
          const float Eg = q2.y * pnorm;
          {
            // ---- L1: the pixel's own bundle box ----
            const float cu = fmaf(q0.x, pc[0], fmaf(q0.y, pc[1], q0.z * pc[2]));
            const float cv = fmaf(q0.w, pc[0], fmaf(q1.x, pc[1], q1.y * pc[2]));
            const float cw = fmaf(q1.z, pc[0], fmaf(q1.w, pc[1], q2.x * pc[2]));
            const float hu = fmaf(fabsf(q0.x), ph[0], fmaf(fabsf(q0.y), ph[1], fabsf(q0.z) * ph[2]));
            const float hv = fmaf(fabsf(q0.w), ph[0], fmaf(fabsf(q1.x), ph[1], fabsf(q1.y) * ph[2]));
            const float hw = fmaf(fabsf(q1.z), ph[0], fmaf(fabsf(q1.w), ph[1], fabsf(q2.x) * ph[2]));
            const float m3 = fminf(fminf(cu + hu, cv + hv), cw + hw);
            const float mN = (cu + cv + cw) + (hu + hv + hw);
            RT_PC(3);
            if ((m3 < -Eg) || (mN + Eg < q2.z)) continue;
            RT_PC(4);
          }
          // ---- L2 / EX: per shadow ray ----
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            if (!((todo >> k) & 1u) || ((occluded >> k) & 1u)) continue;
            const float px = xadd(cx, xmul(ht[k], dxs[k / 3])), py = xadd(cy, xmul(ht[k], dys[k % 3])),
                        pz = xadd(cz, xmul(ht[k], dz));
            const float gx = -xsub(Lx, px), gy = -xsub(Ly, py), gz = -xsub(Lz, pz);
            const float mU = fmaf(q0.x, gx, fmaf(q0.y, gy, q0.z * gz));
            const float mV = fmaf(q0.w, gx, fmaf(q1.x, gy, q1.y * gz));
            const float mW = fmaf(q1.z, gx, fmaf(q1.w, gy, q2.x * gz));
            const float m3 = fminf(fminf(mU, mV), mW);
            const float mN = mU + mV + mW;
            RT_PC(5);
            if (m3 < -Eg || mN + Eg < q2.z) continue;          // definite miss / behind the start
            RT_PC(6);
            if (m3 >= Eg && mN - Eg >= q2.w && q2.z >= 1e-4f * mN) {
              occluded |= 1u << k;                              // definite occluder
              continue;
            }
            ++n_exact;
            if (rt_ex_shadow(p.geom, p.src, p.sph, tri, idx[k], px, py, pz, Lx, Ly, Lz)) occluded |= 1u << k;
          }
        }
      }
    }

    // sphere occluders + the lighting tail of DirectLight, then (single light)
    // the pixelColour accumulation in the reference's order (:151-156)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (!((active >> k) & 1u)) continue;
      const float px = xadd(cx, xmul(ht[k], dxs[k / 3])), py = xadd(cy, xmul(ht[k], dys[k % 3])),
                  pz = xadd(cz, xmul(ht[k], dz));
      const RtShade s = rt_ex_shade(p.src, p.sph, p.n_sph, idx[k], px, py, pz, Lx, Ly, Lz, p.lights[l][4],
                                    p.lights[l][5], p.lights[l][6], (occluded >> k) & 1u, 1);
      if (MULTI) {
        dl[(l * 9 + k) * 3 + 0] = s.pw[0]; dl[(l * 9 + k) * 3 + 1] = s.pw[1]; dl[(l * 9 + k) * 3 + 2] = s.pw[2];
      } else {
        pix[0] = xadd(pix[0], s.pw[0]); pix[1] = xadd(pix[1], s.pw[1]); pix[2] = xadd(pix[2], s.pw[2]);
        pix[0] = xadd(pix[0], xmul(s.col[0], 0.5f));
        pix[1] = xadd(pix[1], xmul(s.col[1], 0.5f));
        pix[2] = xadd(pix[2], xmul(s.col[2], 0.5f));
      }
    }
  }

  if (live) {
    if (MULTI || p.n_lights == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        if (!((active >> k) & 1u)) continue;
        if (MULTI)
          for (int l = 0; l < p.n_lights; ++l) {
            pix[0] = xadd(pix[0], dl[(l * 9 + k) * 3 + 0]);
            pix[1] = xadd(pix[1], dl[(l * 9 + k) * 3 + 1]);
            pix[2] = xadd(pix[2], dl[(l * 9 + k) * 3 + 2]);
          }
        const float px = xadd(cx, xmul(ht[k], dxs[k / 3])), py = xadd(cy, xmul(ht[k], dys[k % 3])),
                    pz = xadd(cz, xmul(ht[k], dz));
        const RtShade s = rt_ex_shade(p.src, p.sph, 0, idx[k], px, py, pz, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0, 0);
        pix[0] = xadd(pix[0], xmul(s.col[0], 0.5f));
        pix[1] = xadd(pix[1], xmul(s.col[1], 0.5f));
        pix[2] = xadd(pix[2], xmul(s.col[2], 0.5f));
      }
    }
    rt_store_pixel(p, pid, active != 0, pix);
  }
  rt_count(p.counters + 0, (unsigned long long)__popc(active) * p.n_lights);
  rt_count(p.counters + 1, (unsigned long long)n_exact);
#ifdef RT_PROFILE_COUNTERS
  if (live) for (int i = 0; i < 6; ++i) atomicAdd(p.counters + 2 + i, (unsigned long long)prof_cnt[i + (i >= 2 ? 1 : 0)]);
#endif
}

int rt_launch_filtered(b200_ctx *ctx, const RtFrame &f, RtKParams &p);
