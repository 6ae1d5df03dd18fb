// Production raytracer kernel: brute force over every (ray, triangle) pair, but
// almost every pair is decided by a conservative edge-function filter; the
// reference arithmetic (rt_exact.cuh) runs only where its outcome can matter.
// The frame is bit-identical to rt_bruteforce_kernel and to the reference.
//
// Why this is legitimate.  ClosestIntersection (raytracer/Source/skeleton.cpp:
// 263-363) keeps the hit with the smallest `distance`, lowest index on ties; a
// triangle that fails `check` (:328) or `distance < 0` (:311) never changes its
// state.  So any pair that PROVABLY fails one of those tests in the reference's
// own float arithmetic can be skipped, and any pair that PROVABLY passes `check`
// needs only the t/distance part.  All primary rays share the camera origin and
// all shadow rays of one light end at that light, so per (origin, triangle) the
// three Cramer numerators are linear in the ray direction d:
//     sigma*Du = d.PU   sigma*Dv = d.PV   sigma*(D-Du-Dv) = d.PW      (planes)
// with sigma the sign that t >= 0 forces on det(A).  A pair is a definite miss
// when one of the three is below -E, a definite inside when all are above +E,
// where E bounds (reference rounding error + filter rounding error); see
// rt_prep_planes_kernel for the bound.  Pairs in the +-E band run the reference
// arithmetic unchanged.
//
// Hierarchy (per warp = 8x4 pixel patch, 9 samples per pixel):
//   L0  each LANE tests a different triangle's planes against the bounding box
//       of the whole warp's ray bundle; __ballot_sync gives the survivors
//       (warp vote early-out: 32 triangles x 288 rays per ~20 instructions);
//   L1  for each survivor, each lane tests its pixel's 9-ray bundle box;
//   L2  per ray: the three plane values decide miss / inside / ambiguous;
//   EX  reference arithmetic for what is left (about two per ray).
// Plane records are streamed through shared memory in tiles by 1-D TMA bulk
// copies (cp.async.bulk + mbarrier), double buffered.
#pragma once
#include "rt_common.cuh"
#include "rt_exact.cuh"

constexpr int RT_TILE = 256;      // triangles per shared-memory tile
constexpr int RT_THREADS = 256;   // 8 warps: a 16x16 pixel block
constexpr int RT_REC_F4 = 3;      // float4s per plane record (48 bytes)

// ---- mbarrier / TMA bulk-copy primitives (sm_90+/sm_100a PTX) ------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on the mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- per-frame plane records ------------------------------------------------------
// origin 0 = camera (directions (dx, dy, f): the z term is folded into a constant)
// origin 1+l = light l (directions g = hit - light, general 3-D)
//   q0 = (Ux, Uy, Uz|cU, Vx)  q1 = (Vy, Vz|cV, Wx, Wy)  q2 = (Wz|cW, E, T0, T1)
// camera: E absolute (max |d| folded in), T0 = lower bound of |Dt|, T1 = E + 0.5*max(|Px|+|Py|)
// light : E per unit |g|_inf, T0/T1 = reject / definite thresholds of d.PN (t >= 0 test)
struct RtPrepParams {
  const float4 *geom;
  int n_tris;
  float cam[3];
  float focal;
  float dmax;        // upper bound of |d|_inf over the frame's primary rays
  float world_S;     // upper bound of |start - v0|_inf for any ray start in the scene
  int n_lights;
  float lights[B200_MAX_LIGHTS][3];
  float4 *planes;    // [(1 + n_lights)][n_tiles * RT_TILE][3]
  size_t origin_stride_f4;
};

__global__ void rt_prep_planes_kernel(const __grid_constant__ RtPrepParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int o = blockIdx.y;  // origin
  if (i >= p.n_tris) return;
  const float4 g0 = p.geom[3 * i], g1 = p.geom[3 * i + 1], g2 = p.geom[3 * i + 2];
  const double e1[3] = {g0.w, g1.x, g1.y}, e2[3] = {g1.z, g1.w, g2.x};
  const float v0[3] = {g0.x, g0.y, g0.z};
  double s[3];
  if (o == 0) {
    // exactly the reference's `start - v0` (skeleton.cpp:296)
    for (int k = 0; k < 3; ++k) s[k] = (double)xsub(p.cam[k], v0[k]);
  } else {
    for (int k = 0; k < 3; ++k) s[k] = (double)p.lights[o - 1][k] - (double)v0[k];
  }
  const double N[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2],
                       e1[0] * e2[1] - e1[1] * e2[0]};
  const double Dt = s[0] * N[0] + s[1] * N[1] + s[2] * N[2];
  const double sg = Dt < 0 ? -1.0 : 1.0;
  // PU = -sg (s x e2), PV = -sg (e1 x s), PN = -sg (e1 x e2), PW = PN - PU - PV
  double PU[3] = {-sg * (s[1] * e2[2] - s[2] * e2[1]), -sg * (s[2] * e2[0] - s[0] * e2[2]),
                  -sg * (s[0] * e2[1] - s[1] * e2[0])};
  double PV[3] = {-sg * (e1[1] * s[2] - e1[2] * s[1]), -sg * (e1[2] * s[0] - e1[0] * s[2]),
                  -sg * (e1[0] * s[1] - e1[1] * s[0])};
  double PW[3];
  for (int k = 0; k < 3; ++k) PW[k] = -sg * N[k] - PU[k] - PV[k];

  const double eps = 5.9604644775390625e-08;  // 2^-24
  const double a1 = fmax(fabs(e1[0]), fmax(fabs(e1[1]), fabs(e1[2])));
  const double a2 = fmax(fabs(e2[0]), fmax(fabs(e2[1]), fabs(e2[2])));
  const double N1 = fabs(N[0]) + fabs(N[1]) + fabs(N[2]);
  const double N2 = sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
  float4 q0, q1, q2;
  bool degenerate = !(isfinite(Dt) && isfinite(N1) && a1 < 1e18 && a2 < 1e18);
  if (o == 0) {
    const double sm = fmax(fabs(s[0]), fmax(fabs(s[1]), fabs(s[2])));
    // Rounding-error bound of the reference's three determinants plus this
    // filter's own (planes rounded to float, FMA chains): each is below
    // 54 eps |d|inf X Y; 256 leaves a wide safety factor.
    const double E = 256.0 * eps * (a1 * a2 + sm * (a1 + a2)) * (double)p.dmax * 1.01;
    const double dt_lo = fabs(Dt) * (1.0 - 1e-6) - 256.0 * eps * sm * a1 * a2;
    if (!(dt_lo > 0.0)) degenerate = true;   // camera (numerically) in the triangle's plane
    const double f = (double)p.focal;
    const double h = 0.5 * fmax(fabs(PU[0]) + fabs(PU[1]), fmax(fabs(PV[0]) + fabs(PV[1]), fabs(PW[0]) + fabs(PW[1])));
    q0 = make_float4((float)PU[0], (float)PU[1], (float)(PU[2] * f), (float)PV[0]);
    q1 = make_float4((float)PV[1], (float)(PV[2] * f), (float)PW[0], (float)PW[1]);
    q2 = make_float4((float)(PW[2] * f), __double2float_ru(E), __double2float_rd(dt_lo),
                     __double2float_ru((E + h) * (1.0 + 1e-6)));
  } else {
    const double S = (double)p.world_S;
    const double E = 256.0 * eps * (a1 * a2 + S * (a1 + a2)) + 6.2e-5 * (a1 + a2);
    const double slackT = 256.0 * eps * S * a1 * a2 + 1.01e-5 * N1 + 1e-6 * fabs(Dt);
    // light closer than 1e-3 S to the plane: the "beyond the light" argument
    // loses its margin, leave the triangle to the reference arithmetic
    if (!(fabs(Dt) > 1e-3 * S * N2) || !(fabs(Dt) - slackT > 0.0)) degenerate = true;
    q0 = make_float4((float)PU[0], (float)PU[1], (float)PU[2], (float)PV[0]);
    q1 = make_float4((float)PV[1], (float)PV[2], (float)PW[0], (float)PW[1]);
    q2 = make_float4((float)PW[2], __double2float_ru(E), __double2float_rd(fabs(Dt) - slackT),
                     __double2float_ru(fabs(Dt) + slackT));
  }
  if (degenerate || !isfinite(q0.x + q0.y + q0.z + q0.w + q1.x + q1.y + q1.z + q1.w + q2.x)) {
    // always "ambiguous": every ray runs the reference arithmetic on this triangle
    q0 = make_float4(0.f, 0.f, 0.f, 0.f);
    q1 = make_float4(0.f, 0.f, 0.f, 0.f);
    q2 = make_float4(0.f, INFINITY, -INFINITY, INFINITY);
  }
  float4 *dst = p.planes + (size_t)o * p.origin_stride_f4 + (size_t)i * RT_REC_F4;
  dst[0] = q0; dst[1] = q1; dst[2] = q2;
}

// ---- warp helpers -------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Reference arithmetic for one (primary ray, triangle) pair; `inside` = the
// filter proved `check` (:328) true, so u and v need not be formed.
__device__ __forceinline__ void rt_ex_primary(const RtKParams &p, int tri, bool inside, float dx, float dy,
                                              float dz, float len, RtHit &best) {
  const float4 g0 = __ldg(p.geom + 3 * tri), g1 = __ldg(p.geom + 3 * tri + 1), g2 = __ldg(p.geom + 3 * tri + 2);
  const float e1x = g0.w, e1y = g1.x, e1z = g1.y, e2x = g1.z, e2y = g1.w, e2z = g2.x;
  const float sx = xsub(p.cam[0], g0.x), sy = xsub(p.cam[1], g0.y), sz = xsub(p.cam[2], g0.z);
  const float D = xdet3(-dx, -dy, -dz, e1x, e1y, e1z, e2x, e2y, e2z);
  const float t = xdiv(xdet3(sx, sy, sz, e1x, e1y, e1z, e2x, e2y, e2z), D);
  const float distance = xmul(t, len);
  if (distance < 0.0f) return;
  if (distance >= best.dist || distance > FLT_MAX) return;
  if (!inside) {
    const float u = xdiv(xdet3(-dx, -dy, -dz, sx, sy, sz, e2x, e2y, e2z), D);
    const float v = xdiv(xdet3(-dx, -dy, -dz, e1x, e1y, e1z, sx, sy, sz), D);
    if (!((u >= 0) && (v >= 0) && (xadd(u, v) <= 1))) return;
  }
  best.t = t; best.dist = distance; best.idx = tri;
}

// Reference arithmetic for one (shadow ray, triangle) pair: does this triangle
// make DirectLight return black (skeleton.cpp:394-397)?  Any accepted hit with
// distance < r_magnitude does, because ClosestIntersection keeps the minimum.
__device__ __forceinline__ bool rt_ex_shadow(const RtKParams &p, int tri, float ox, float oy, float oz,
                                             float rx, float ry, float rz, float len, float r_mag) {
  const float4 g0 = __ldg(p.geom + 3 * tri), g1 = __ldg(p.geom + 3 * tri + 1), g2 = __ldg(p.geom + 3 * tri + 2);
  const float e1x = g0.w, e1y = g1.x, e1z = g1.y, e2x = g1.z, e2y = g1.w, e2z = g2.x;
  const float sx = xsub(ox, g0.x), sy = xsub(oy, g0.y), sz = xsub(oz, g0.z);
  const float D = xdet3(-rx, -ry, -rz, e1x, e1y, e1z, e2x, e2y, e2z);
  const float t = xdiv(xdet3(sx, sy, sz, e1x, e1y, e1z, e2x, e2y, e2z), D);
  const float distance = xmul(t, len);
  if (distance < 0.0f) return false;
  if (!(distance < r_mag) || distance > FLT_MAX) return false;
  const float u = xdiv(xdet3(-rx, -ry, -rz, sx, sy, sz, e2x, e2y, e2z), D);
  const float v = xdiv(xdet3(-rx, -ry, -rz, e1x, e1y, e1z, sx, sy, sz), D);
  return (u >= 0) && (v >= 0) && (xadd(u, v) <= 1);
}

// Conservative sphere pre-test: the reference reports no root when
// discriminant = b*b - 4*a*c < 0 (TestModelH.h:27-28).  disc = 4 q with
// q = (d.L)^2 - (d.d)(L.L - r^2); reject when q is negative beyond the rounding
// error of either evaluation.
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

template <bool MULTI>
__global__ void __launch_bounds__(RT_THREADS) rt_filtered_kernel(const __grid_constant__ RtKParams p) {
  extern __shared__ __align__(128) float4 tile_smem[];  // 2 x RT_TILE x 3 float4
  __shared__ __align__(8) uint64_t bars[2];
  constexpr int NL = MULTI ? B200_MAX_LIGHTS : 1;

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
  uint32_t phase_bits = 0;  // bit b = parity to wait for on bars[b]
  int issued = 0;           // tiles issued so far across all phases (buffer = issued & 1)

  // dir = R * vec4(u - W/2, v - H/2, f, 1)   (skeleton.cpp:126-128)
  const float x = (float)(u - p.W / 2), y = (float)(v - p.H / 2);
  const float dir0 = xadd(xadd(xmul(p.R[0], x), xmul(p.R[4], y)), xadd(xmul(p.R[8], p.focal), xmul(p.R[12], 1.0f)));
  const float dir1 = xadd(xadd(xmul(p.R[1], x), xmul(p.R[5], y)), xadd(xmul(p.R[9], p.focal), xmul(p.R[13], 1.0f)));
  const float dz = p.focal;

  RtHit best[9];
  float len[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    best[k].t = 0.f; best[k].dist = FLT_MAX; best[k].idx = 0;
    const float dx = xadd(dir0, xmul(0.5f, (float)(k / 3 - 1)));
    const float dy = xadd(dir1, xmul(0.5f, (float)(k % 3 - 1)));
    len[k] = xsqrt(xdot3(dx, dy, dz, dx, dy, dz));
  }
  unsigned long long n_exact = 0;

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
    if (threadIdx.x == 0 && n_tiles > 0) {
      const int cnt = min(RT_TILE, p.n_tris);
      mbar_expect_tx(&bars[issued & 1], cnt * 48u);
      tma_bulk_g2s(tile_smem + (size_t)(issued & 1) * RT_TILE * RT_REC_F4, src, cnt * 48u, &bars[issued & 1]);
    }
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int buf = issued & 1;
      ++issued;
      // everyone is done with the other buffer (it held tile-1): refill it
      __syncthreads();
      if (threadIdx.x == 0 && tile + 1 < n_tiles) {
        const int cnt = min(RT_TILE, p.n_tris - (tile + 1) * RT_TILE);
        mbar_expect_tx(&bars[buf ^ 1], cnt * 48u);
        tma_bulk_g2s(tile_smem + (size_t)(buf ^ 1) * RT_TILE * RT_REC_F4,
                     src + (size_t)(tile + 1) * RT_TILE * RT_REC_F4, cnt * 48u, &bars[buf ^ 1]);
      }
      mbar_wait(&bars[buf], (phase_bits >> buf) & 1u);
      phase_bits ^= 1u << buf;
      const float4 *T = tile_smem + (size_t)buf * RT_TILE * RT_REC_F4;
      const int base = tile * RT_TILE;
      const int cnt = min(RT_TILE, p.n_tris - base);
      if (!warp_live) continue;
      for (int r0 = 0; r0 < cnt; r0 += 32) {
        // ---- L0: one triangle per lane vs the warp's bundle box ----
        bool pass = false;
        if (r0 + lane < cnt) {
          const float4 q0 = T[(r0 + lane) * 3], q1 = T[(r0 + lane) * 3 + 1], q2 = T[(r0 + lane) * 3 + 2];
          const float mE = -q2.y;
          const float bu = fmaf(q0.x, wc0, fmaf(q0.y, wc1, q0.z)) + fmaf(fabsf(q0.x), wh0, fabsf(q0.y) * wh1);
          const float bv = fmaf(q0.w, wc0, fmaf(q1.x, wc1, q1.y)) + fmaf(fabsf(q0.w), wh0, fabsf(q1.x) * wh1);
          const float bw = fmaf(q1.z, wc0, fmaf(q1.w, wc1, q2.x)) + fmaf(fabsf(q1.z), wh0, fabsf(q1.w) * wh1);
          // E already covers the rounding of the centre evaluation (|wc| <= dmax);
          // the half-width terms are sums of non-negative products, inflated above
          pass = !(fminf(fminf(bu, bv), bw) < mE);
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
          if (fminf(fminf(cu, cv), cw) < -q2.w) continue;
          const float E = q2.y, dt_lo = q2.z;
          // ---- L2 / EX: per ray ----
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const float dx = xadd(dir0, xmul(0.5f, (float)(k / 3 - 1)));
            const float dy = xadd(dir1, xmul(0.5f, (float)(k % 3 - 1)));
            const float mU = fmaf(q0.x, dx, fmaf(q0.y, dy, q0.z));
            const float mV = fmaf(q0.w, dx, fmaf(q1.x, dy, q1.y));
            const float mW = fmaf(q1.z, dx, fmaf(q1.w, dy, q2.x));
            const float m3 = fminf(fminf(mU, mV), mW);
            if (m3 >= -E) {
              const float mN = mU + mV + mW;
              // reference distance >= dt_lo*len/(mN+E): cannot beat the current closest
              const bool farther = (mN > E) && (dt_lo * len[k] >= best[k].dist * (mN + E));
              if (!farther) {
                ++n_exact;
                rt_ex_primary(p, tri, m3 >= E, dx, dy, dz, len[k], best[k]);
              }
            }
          }
        }
      }
    }
  }

  // spheres (skeleton.cpp:341-355) + hit positions (skeleton.cpp:326/345)
  float px[9], py[9], pz[9];
  unsigned active = 0;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float dx = xadd(dir0, xmul(0.5f, (float)(k / 3 - 1)));
    const float dy = xadd(dir1, xmul(0.5f, (float)(k % 3 - 1)));
    if (live) {
      for (int s = 0; s < p.n_sph; ++s) {
        if (!rt_sphere_may_hit(p.sph[s], p.cam[0], p.cam[1], p.cam[2], dx, dy, dz)) continue;
        float t;
        if (rt_exact_sphere(p.sph[s], p.cam[0], p.cam[1], p.cam[2], dx, dy, dz, t)) {
          if (t < best[k].dist) { best[k].t = t; best[k].dist = t; best[k].idx = -1 - s; }
        }
      }
    }
    const bool hit = live && best[k].dist < FLT_MAX;
    if (hit) active |= 1u << k;
    px[k] = xadd(p.cam[0], xmul(best[k].t, dx));
    py[k] = xadd(p.cam[1], xmul(best[k].t, dy));
    pz[k] = xadd(p.cam[2], xmul(best[k].t, dz));
    if (!hit) { px[k] = p.cam[0]; py[k] = p.cam[1]; pz[k] = p.cam[2]; }
  }
  if (live) {
    const bool hit4 = (active >> 4) & 1u;
    if (p.depth) p.depth[pid] = hit4 ? best[4].dist : INFINITY;
    if (p.index) p.index[pid] = hit4 ? best[4].idx : INT32_MIN;
  }
  int idx[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) idx[k] = best[k].idx;

  // ============================ shadow rays ============================
  float dl[NL][9][3];
  const unsigned warp_active = __ballot_sync(0xffffffffu, active != 0);
  for (int l = 0; l < p.n_lights; ++l) {
    const int ls = MULTI ? l : 0;
    const float Lx = p.lights[l][0], Ly = p.lights[l][1], Lz = p.lights[l][2];
    unsigned occluded = 0;
    // g = hit - light per ray; boxes of the pixel's and the warp's bundles
    float glo[3] = {INFINITY, INFINITY, INFINITY}, ghi[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if ((active >> k) & 1u) {
        const float gx = -xsub(Lx, px[k]), gy = -xsub(Ly, py[k]), gz = -xsub(Lz, pz[k]);
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
    if (threadIdx.x == 0 && n_tiles > 0) {
      const int cnt = min(RT_TILE, p.n_tris);
      mbar_expect_tx(&bars[issued & 1], cnt * 48u);
      tma_bulk_g2s(tile_smem + (size_t)(issued & 1) * RT_TILE * RT_REC_F4, src, cnt * 48u, &bars[issued & 1]);
    }
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int buf = issued & 1;
      ++issued;
      __syncthreads();
      if (threadIdx.x == 0 && tile + 1 < n_tiles) {
        const int cnt = min(RT_TILE, p.n_tris - (tile + 1) * RT_TILE);
        mbar_expect_tx(&bars[buf ^ 1], cnt * 48u);
        tma_bulk_g2s(tile_smem + (size_t)(buf ^ 1) * RT_TILE * RT_REC_F4,
                     src + (size_t)(tile + 1) * RT_TILE * RT_REC_F4, cnt * 48u, &bars[buf ^ 1]);
      }
      mbar_wait(&bars[buf], (phase_bits >> buf) & 1u);
      phase_bits ^= 1u << buf;
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
            if ((m3 < -Eg) || (mN + Eg < q2.z)) continue;
          }
          // ---- L2 / EX: per shadow ray ----
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            if (!((todo >> k) & 1u) || ((occluded >> k) & 1u)) continue;
            const float rx = xsub(Lx, px[k]), ry = xsub(Ly, py[k]), rz = xsub(Lz, pz[k]);  // :370
            const float gx = -rx, gy = -ry, gz = -rz;
            const float mU = fmaf(q0.x, gx, fmaf(q0.y, gy, q0.z * gz));
            const float mV = fmaf(q0.w, gx, fmaf(q1.x, gy, q1.y * gz));
            const float mW = fmaf(q1.z, gx, fmaf(q1.w, gy, q2.x * gz));
            const float m3 = fminf(fminf(mU, mV), mW);
            const float mN = mU + mV + mW;
            if (m3 < -Eg || mN + Eg < q2.z) continue;          // definite miss / behind the start
            if (m3 >= Eg && mN - Eg >= q2.w && q2.z >= 1e-4f * mN) {
              occluded |= 1u << k;                              // definite occluder
              continue;
            }
            ++n_exact;
            float col[3], nx, ny, nz;
            rt_surface(p, idx[k], px[k], py[k], pz[k], col, nx, ny, nz);
            const float ox = xadd(px[k], xmul(nx, 0.00001f)), oy = xadd(py[k], xmul(ny, 0.00001f)),
                        oz = xadd(pz[k], xmul(nz, 0.00001f));   // :394
            const float slen = xsqrt(xdot3(rx, ry, rz, rx, ry, rz));
            const float r_mag = rt_exact_rmag(rx, ry, rz);
            if (rt_ex_shadow(p, tri, ox, oy, oz, rx, ry, rz, slen, r_mag)) occluded |= 1u << k;
          }
        }
      }
    }

    // spheres as occluders + the lighting tail of DirectLight
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      dl[ls][k][0] = dl[ls][k][1] = dl[ls][k][2] = 0.f;
      if (!((active >> k) & 1u)) continue;
      const float rx = xsub(Lx, px[k]), ry = xsub(Ly, py[k]), rz = xsub(Lz, pz[k]);
      const float r_mag = rt_exact_rmag(rx, ry, rz);
      float col[3], nx, ny, nz;
      rt_surface(p, idx[k], px[k], py[k], pz[k], col, nx, ny, nz);
      bool occ = (occluded >> k) & 1u;
      if (!occ && p.n_sph > 0) {
        const float ox = xadd(px[k], xmul(nx, 0.00001f)), oy = xadd(py[k], xmul(ny, 0.00001f)),
                    oz = xadd(pz[k], xmul(nz, 0.00001f));
        for (int s = 0; s < p.n_sph && !occ; ++s) {
          if (!rt_sphere_may_hit(p.sph[s], ox, oy, oz, rx, ry, rz)) continue;
          float t;
          if (rt_exact_sphere(p.sph[s], ox, oy, oz, rx, ry, rz, t)) occ = t < r_mag;  // :348 + :395
        }
      }
      if (occ) continue;
      rt_exact_lambert(rx, ry, rz, r_mag, nx, ny, nz, col, &p.lights[l][4], dl[ls][k]);
    }
  }

  // pixelColour accumulation in the reference's order (skeleton.cpp:134-158)
  if (live) {
    float pix[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (!((active >> k) & 1u)) continue;
      for (int l = 0; l < p.n_lights; ++l) {
        const int ls = MULTI ? l : 0;
        pix[0] = xadd(pix[0], dl[ls][k][0]); pix[1] = xadd(pix[1], dl[ls][k][1]); pix[2] = xadd(pix[2], dl[ls][k][2]);
      }
      float col[3], nx, ny, nz;
      rt_surface(p, idx[k], px[k], py[k], pz[k], col, nx, ny, nz);
      pix[0] = xadd(pix[0], xmul(col[0], 0.5f));
      pix[1] = xadd(pix[1], xmul(col[1], 0.5f));
      pix[2] = xadd(pix[2], xmul(col[2], 0.5f));
    }
    rt_store_pixel(p, pid, active != 0, pix);
  }
  rt_count(p.counters + 0, (unsigned long long)__popc(active) * p.n_lights);
  rt_count(p.counters + 1, n_exact);
}

int rt_launch_filtered(b200_ctx *ctx, const RtFrame &f, RtKParams &p);
