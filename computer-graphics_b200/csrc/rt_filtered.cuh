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
// direction grids (rt_grid.cuh)
constexpr int RT_GRID_G = 64;              // light cube map: cells per face edge
constexpr int RT_GRID_FACE = RT_GRID_G * RT_GRID_G;
constexpr int RT_WTILE = 64;               // records per tile of a warp's private ring (grid kernels)
constexpr int RT_GRID_MAX_CELLS = 96;      // shadow phase: distinct cells one warp (8x4 pixels, 288 rays) may walk, else it streams the scene
constexpr int RT_GRID_TABLE_LOG2 = 8;      // hash set used to collect them
constexpr int RT_SEEN_LOG2 = 8, RT_SEEN = 1 << RT_SEEN_LOG2;   // per-warp set of triangles already met in this light phase

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
  float n_scale;     // max(1, max |normal component|): the shadow-ray start is hit + 1e-5 * normal (:394)
  int n_lights;
  float lights[B200_MAX_LIGHTS][3];
  float4 *planes;    // [(1 + n_lights)][n_tiles * RT_TILE][3]
  size_t origin_stride_f4;
  float *dt_cam;     // [n_tris] det(start - v0, e1, e2) for start = camera, in the reference's own arithmetic
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
    const double ns = (double)p.n_scale;
    const double E = 256.0 * eps * (a1 * a2 + S * (a1 + a2)) + 6.2e-5 * ns * (a1 + a2);
    const double slackT = 256.0 * eps * S * a1 * a2 + 1.01e-5 * ns * N1 + 1e-6 * fabs(Dt);
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
  if (o == 0) {
    // numerator of t (skeleton.cpp:305-306): the same for every primary ray, exact float ops
    const float sx = xsub(p.cam[0], v0[0]), sy = xsub(p.cam[1], v0[1]), sz = xsub(p.cam[2], v0[2]);
    p.dt_cam[i] = xdet3(sx, sy, sz, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x);
  }
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

// ---- bundle-box tests shared by the render kernel (L0) and the grid builder -------
// A triangle can matter to a ray only if all three edge values are >= -E there.  Over a
// box of directions (centre c, half-widths h >= 0) each edge value is linear, so its
// maximum is known exactly; the same holds for the sums of two and of all three, which
// must then be >= -2E / -3E.  The sums catch what the single planes cannot: triangles
// seen nearly edge-on, whose two long edge planes almost coincide and straddle the box.
// (Coefficient sums are rounded once more; the factors leave room for that.)
// PAIRS = false (small scenes streamed whole, where a few extra L0 survivors cost less than
// the test) checks the single planes only.
template <bool PAIRS = true>
__device__ __forceinline__ bool rt_box_may_hit_light(const float4 q0, const float4 q1, const float4 q2, const float *c,
                                                     const float *h, float Eg) {
  const float U[3] = {q0.x, q0.y, q0.z}, V[3] = {q0.w, q1.x, q1.y}, W[3] = {q1.z, q1.w, q2.x};
  auto hi = [&](float x, float y, float z) {
    return fmaf(x, c[0], fmaf(y, c[1], z * c[2])) + fmaf(fabsf(x), h[0], fmaf(fabsf(y), h[1], fabsf(z) * h[2]));
  };
  if (fminf(fminf(hi(U[0], U[1], U[2]), hi(V[0], V[1], V[2])), hi(W[0], W[1], W[2])) < -Eg) return false;
  if (!PAIRS) return true;
  const float uv = hi(U[0] + V[0], U[1] + V[1], U[2] + V[2]);
  const float vw = hi(V[0] + W[0], V[1] + W[1], V[2] + W[2]);
  const float uw = hi(U[0] + W[0], U[1] + W[1], U[2] + W[2]);
  if (fminf(fminf(uv, vw), uw) < -2.1f * Eg) return false;
  return true;
}
// camera: directions (dx, dy, f) with the z term folded into the constant; E absolute
template <bool PAIRS = true>
__device__ __forceinline__ bool rt_box_may_hit_cam(const float4 q0, const float4 q1, const float4 q2, float c0, float c1,
                                                   float h0, float h1) {
  auto hi = [&](float x, float y, float k) {
    return fmaf(x, c0, fmaf(y, c1, k)) + fmaf(fabsf(x), h0, fabsf(y) * h1);
  };
  const float E = q2.y;
  if (fminf(fminf(hi(q0.x, q0.y, q0.z), hi(q0.w, q1.x, q1.y)), hi(q1.z, q1.w, q2.x)) < -E) return false;
  if (!PAIRS) return true;
  const float uv = hi(q0.x + q0.w, q0.y + q1.x, q0.z + q1.y);
  const float vw = hi(q0.w + q1.z, q1.x + q1.w, q1.y + q2.x);
  const float uw = hi(q0.x + q1.z, q0.y + q1.w, q0.z + q2.x);
  if (fminf(fminf(uv, vw), uw) < -2.1f * E) return false;
  return true;
}

#include "rt_filtered_kernel.cuh"
