// Direction grids for large raytracer scenes: per-origin lists of the triangles
// that can matter to the rays of one direction cell, built every frame on the GPU
// with atomic appends.  rt_filtered_kernel<.., GRID = true> then streams only the
// lists of the cells its pixel block looks through instead of the whole scene.
//
// A triangle is left out of a cell's list only when the SAME conservative
// edge-function test the render kernel uses (rt_filtered.cuh, level L0) proves a
// definite miss for the cell's whole bundle of directions, so the proof that the
// frame equals the reference's (skeleton.cpp:263-363) carries over unchanged: a
// listed triangle is still decided per ray by L1 / L2 / the reference arithmetic.
//
//   origin 0 (camera)  cells = the render kernel's 16x16 pixel blocks; the
//                      bundle is the block's (dx, dy) box with the +-0.5 jitter
//   origin 1+l (light) cells = a cube map around the light, RT_GRID_G x RT_GRID_G
//                      per face; the edge tests are homogeneous in g = hit - light
//                      (E is per unit |g|inf), so they are run on the face plane
//                      |g_a| = 1; the t >= 0 test is used by sign only
// Lists are unordered (atomic cursor); the render kernel breaks distance ties by
// triangle index explicitly, which is what the reference's ascending loop with a
// strict `<` does (skeleton.cpp:313).
#pragma once
#include "rt_filtered.cuh"

struct RtGridParams {
  const float4 *planes;
  size_t origin_stride_f4;
  int n_tris;
  float R[16];
  float focal;
  int W, H, row0, row1;
  int gx, gy;                 // camera grid = the render kernel's block grid
  int n_lights;
  unsigned *cell_cnt;         // [cells] entries per cell
  unsigned *cell_cursor;      // [cells] fill pass
  const unsigned *cell_off;   // [cells + 1] first entry (multiples of 4)
  float4 *cell_rec;           // [entries][3] plane records
  int *cell_idx;              // [entries] triangle indices
  unsigned long long cap;     // entries allocated
  // counting pass, optional: every (cell, record) pair it finds is also written down, so that the lists can be
  // filled by rt_grid_place_kernel from the pairs instead of by a second descent of every triangle
  uint2 *pairs;               // (cell, origin * n_tris + triangle)
  unsigned long long pair_cap;
  unsigned long long *pair_cursor;   // pairs found (may exceed pair_cap: then the lists are filled the old way)
};

__device__ __forceinline__ float rt_grid_dir0(const RtGridParams &p, int u, int v) {
  const float x = (float)(u - p.W / 2), y = (float)(v - p.H / 2);
  return xadd(xadd(xmul(p.R[0], x), xmul(p.R[4], y)), xadd(xmul(p.R[8], p.focal), xmul(p.R[12], 1.0f)));
}
__device__ __forceinline__ float rt_grid_dir1(const RtGridParams &p, int u, int v) {
  const float x = (float)(u - p.W / 2), y = (float)(v - p.H / 2);
  return xadd(xadd(xmul(p.R[1], x), xmul(p.R[5], y)), xadd(xmul(p.R[9], p.focal), xmul(p.R[13], 1.0f)));
}

// Camera: can any primary ray of blocks [bx0, bx1) x [by0, by1) need this triangle?
// Every float operation of dir0 / dir1 is monotone in u and in v, so the four corner
// pixels bound every pixel of the range exactly; the rest is level L0 of the kernel.
__device__ __forceinline__ bool rt_grid_test_cam(const RtGridParams &p, const float4 q0, const float4 q1,
                                                 const float4 q2, int bx0, int by0, int bx1, int by1) {
  const int u0 = bx0 * 16, v0 = p.row0 + by0 * 16;
  const int u1 = min(bx1 * 16, p.W) - 1, v1 = min(p.row0 + by1 * 16, p.row1) - 1;
  if (u1 < u0 || v1 < v0) return false;
  const float a0 = rt_grid_dir0(p, u0, v0), a1 = rt_grid_dir0(p, u1, v0), a2 = rt_grid_dir0(p, u0, v1),
              a3 = rt_grid_dir0(p, u1, v1);
  const float b0 = rt_grid_dir1(p, u0, v0), b1 = rt_grid_dir1(p, u1, v0), b2 = rt_grid_dir1(p, u0, v1),
              b3 = rt_grid_dir1(p, u1, v1);
  const float lo0 = fminf(fminf(a0, a1), fminf(a2, a3)), hi0 = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
  const float lo1 = fminf(fminf(b0, b1), fminf(b2, b3)), hi1 = fmaxf(fmaxf(b0, b1), fmaxf(b2, b3));
  const float wc0 = 0.5f * (lo0 + hi0), wc1 = 0.5f * (lo1 + hi1);
  const float wh0 = (0.5f * (hi0 - lo0) + 0.5f) * 1.0001f + 1e-3f;
  const float wh1 = (0.5f * (hi1 - lo1) + 0.5f) * 1.0001f + 1e-3f;
  return rt_box_may_hit_cam(q0, q1, q2, wc0, wc1, wh0, wh1);
}

// Light: cube-map cells [i0, i1) x [j0, j1) (in finest-level units) of `face`.
// face = 2*axis + (negative ? 1 : 0); the other two axes follow cyclically.
__device__ __forceinline__ bool rt_grid_test_light(const float4 q0, const float4 q1, const float4 q2, int face,
                                                   int i0, int j0, int i1, int j1) {
  const float cw = 2.0f / (float)RT_GRID_G;
  const float blo = -1.0f + cw * (float)i0, bhi = -1.0f + cw * (float)i1;
  const float clo = -1.0f + cw * (float)j0, chi = -1.0f + cw * (float)j1;
  const int a = face >> 1, b = (a + 1) % 3;
  const float sa = (face & 1) ? -1.0f : 1.0f;
  const float cb = 0.5f * (blo + bhi), hb = 0.5f * (bhi - blo) * 1.0001f + 0.02f * cw;
  const float cc = 0.5f * (clo + chi), hc = 0.5f * (chi - clo) * 1.0001f + 0.02f * cw;
  float wc[3], wh[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {       // k is a compile-time index: no local memory
    wc[k] = k == a ? sa : (k == b ? cb : cc);
    wh[k] = k == a ? 0.0f : (k == b ? hb : hc);
  }
  const float Eg = q2.y * 1.05f;   // |g|inf <= 1 + the widening on this face
  if (!rt_box_may_hit_light(q0, q1, q2, wc, wh, Eg)) return false;
  // the kernel rejects on mN + Eg*s < |Dt| - slack with the ray's own scale s; |Dt| - slack > 0
  // for every record this can reject (degenerate records carry E = inf), so a bundle whose
  // mN + Eg is negative at scale 1 fails it at every scale
  const float nx = q0.x + q0.w + q1.z, ny = q0.y + q1.x + q1.w, nz = q0.z + q1.y + q2.x;
  const float mN = fmaf(nx, wc[0], fmaf(ny, wc[1], nz * wc[2])) +
                   fmaf(fabsf(nx), wh[0], fmaf(fabsf(ny), wh[1], fabsf(nz) * wh[2]));
  return !(mN + 1.1f * Eg < 0.0f);
}

// One warp per (triangle, origin): descend the origin's grid 8x8 cells at a time.
template <bool FILL>
__global__ void __launch_bounds__(256) rt_grid_bin_kernel(const __grid_constant__ RtGridParams p) {
  const int lane = threadIdx.x & 31;
  const int tri = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int o = blockIdx.y;
  if (tri >= p.n_tris) return;
  const float4 *rec = p.planes + (size_t)o * p.origin_stride_f4 + (size_t)tri * RT_REC_F4;
  const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2);
  // counting pass: the warp's pairs are collected in shared memory and leave with one atomic per warp
  constexpr int PAIR_BUF = 96;
  __shared__ unsigned s_pair[FILL ? 1 : 8][FILL ? 1 : PAIR_BUF];
  unsigned *my_pairs = s_pair[FILL ? 0 : (threadIdx.x >> 5)];
  int n_pair = 0;                                   // warp-uniform
  const bool keep_pairs = !FILL && p.pairs != nullptr;
  auto flush = [&]() {                              // called by the whole warp
    if (n_pair == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(p.pair_cursor, (unsigned long long)n_pair);
    base = __shfl_sync(0xffffffffu, base, 0);
    __syncwarp();
    for (int i = lane; i < n_pair; i += 32)
      if (base + i < p.pair_cap) p.pairs[base + i] = make_uint2(my_pairs[i], (unsigned)o * (unsigned)p.n_tris + (unsigned)tri);
    __syncwarp();
    n_pair = 0;
  };
  // whole warp: `hit` lanes name a cell each
  auto note = [&](bool hit, unsigned cell) {
    if (!keep_pairs) return;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (!m) return;
    if (n_pair + __popc(m) > PAIR_BUF) flush();
    if (hit) my_pairs[n_pair + __popc(m & ((1u << lane) - 1))] = cell;
    n_pair += __popc(m);
  };

  auto emit = [&](unsigned cell) {
    if (!FILL) {
      atomicAdd(p.cell_cnt + cell, 1u);
    } else {
      const unsigned long long e = (unsigned long long)p.cell_off[cell] + atomicAdd(p.cell_cursor + cell, 1u);
      if (e < p.cap) {
        float4 *dst = p.cell_rec + e * RT_REC_F4;
        dst[0] = q0; dst[1] = q1; dst[2] = q2;
        p.cell_idx[e] = tri;
      }
    }
  };

  if (o == 0) {
    const int n2x = (p.gx + 63) >> 6, n2y = (p.gy + 63) >> 6, n2 = n2x * n2y;
    for (int c2base = 0; c2base < n2; c2base += 32) {
      const int c2 = c2base + lane;
      const int x2 = c2 % n2x, y2 = c2 / n2x;
      unsigned m2 = __ballot_sync(0xffffffffu, c2 < n2 && rt_grid_test_cam(p, q0, q1, q2, x2 * 64, y2 * 64,
                                                                          min(x2 * 64 + 64, p.gx), min(y2 * 64 + 64, p.gy)));
      while (m2) {
        const int s2 = c2base + __ffs(m2) - 1;
        m2 &= m2 - 1;
        const int X2 = (s2 % n2x) * 8, Y2 = (s2 / n2x) * 8;    // in units of 8 blocks
#pragma unroll 1
        for (int r1 = 0; r1 < 2; ++r1) {
          const int x1 = X2 + (lane & 7), y1 = Y2 + r1 * 4 + (lane >> 3);
          const bool in1 = x1 * 8 < p.gx && y1 * 8 < p.gy;
          unsigned m1 = __ballot_sync(0xffffffffu, in1 && rt_grid_test_cam(p, q0, q1, q2, x1 * 8, y1 * 8,
                                                                           min(x1 * 8 + 8, p.gx), min(y1 * 8 + 8, p.gy)));
          while (m1) {
            const int j1 = __ffs(m1) - 1;
            m1 &= m1 - 1;
            const int X1 = (X2 + (j1 & 7)) * 8, Y1 = (Y2 + r1 * 4 + (j1 >> 3)) * 8;   // in blocks
#pragma unroll 1
            for (int r0 = 0; r0 < 2; ++r0) {
              const int bx = X1 + (lane & 7), by = Y1 + r0 * 4 + (lane >> 3);
              const bool hit = bx < p.gx && by < p.gy && rt_grid_test_cam(p, q0, q1, q2, bx, by, bx + 1, by + 1);
              if (hit) emit((unsigned)(by * p.gx + bx));
              note(hit, (unsigned)(by * p.gx + bx));
            }
          }
        }
      }
    }
  } else {
    const unsigned base = (unsigned)(p.gx * p.gy) + (unsigned)(o - 1) * 6u * RT_GRID_FACE;
    constexpr int G1 = RT_GRID_G / 8;      // coarse cells per face edge
    static_assert(G1 == 8, "one 8x8 coarse level");
    // whole faces first (lanes 0..5), then 8x8 coarse cells, then the cells
    unsigned faces = __ballot_sync(0xffffffffu, lane < 6 && rt_grid_test_light(q0, q1, q2, lane, 0, 0, RT_GRID_G, RT_GRID_G));
    while (faces) {
      const int face = __ffs(faces) - 1;
      faces &= faces - 1;
#pragma unroll 1
      for (int r1 = 0; r1 < 2; ++r1) {
        const int x1 = lane & 7, y1 = r1 * 4 + (lane >> 3);
        unsigned m1 = __ballot_sync(0xffffffffu, rt_grid_test_light(q0, q1, q2, face, x1 * 8, y1 * 8, x1 * 8 + 8, y1 * 8 + 8));
        while (m1) {
          const int j1 = __ffs(m1) - 1;
          m1 &= m1 - 1;
          const int X1 = (j1 & 7) * 8, Y1 = (r1 * 4 + (j1 >> 3)) * 8;
#pragma unroll 1
          for (int r0 = 0; r0 < 2; ++r0) {
            const int i = X1 + (lane & 7), j = Y1 + r0 * 4 + (lane >> 3);
            const bool hit = rt_grid_test_light(q0, q1, q2, face, i, j, i + 1, j + 1);
            if (hit) emit(base + (unsigned)(face * RT_GRID_FACE + j * RT_GRID_G + i));
            note(hit, base + (unsigned)(face * RT_GRID_FACE + j * RT_GRID_G + i));
          }
        }
      }
    }
  }
  if (keep_pairs) flush();
}

// Fills the cells' lists from the pairs the counting pass wrote down: one thread per pair claims a place in its
// cell (atomic cursor: the lists are unordered anyway) and copies the 48-byte record and the triangle index.
__global__ void __launch_bounds__(256) rt_grid_place_kernel(const __grid_constant__ RtGridParams p, unsigned long long n_pairs) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const uint2 pr = __ldg(p.pairs + i);
  const unsigned o = pr.y / (unsigned)p.n_tris, tri = pr.y - o * (unsigned)p.n_tris;
  const float4 *rec = p.planes + (size_t)o * p.origin_stride_f4 + (size_t)tri * RT_REC_F4;
  const unsigned long long e = (unsigned long long)p.cell_off[pr.x] + atomicAdd(p.cell_cursor + pr.x, 1u);
  if (e < p.cap) {
    float4 *dst = p.cell_rec + e * RT_REC_F4;
    dst[0] = __ldg(rec); dst[1] = __ldg(rec + 1); dst[2] = __ldg(rec + 2);
    p.cell_idx[e] = (int)tri;
  }
}

// counts -> counts rounded up to whole groups of four entries (16-byte aligned index lists)
__global__ void rt_grid_pad_kernel(const unsigned *__restrict__ cnt, unsigned *__restrict__ padded, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) padded[i] = (cnt[i] + 3u) & ~3u;
}

// ---- the order of a gridded frame's blocks, from the cost of the frame before --------------------------------
// On the 100 800-triangle scene 0.1 % of the warps run 20 - 70 times longer than the mean (8 x 4 pixel patches on a
// silhouette whose shadow rays graze a finely tessellated face: thousands of candidates, each lane on a different
// one).  Their blocks sit two thirds down the frame, so in launch order they start late and the kernel ends on them;
// shared among N GPUs, one such warp alone is longer than a rank's whole share.  Every GRID kernel therefore leaves
// the cycles of each block's slowest warp behind, and the next frame of the same shape is launched from a plan:
// blocks whose slowest warp took more than RT_PLAN_HEAVY times the mean come FIRST and are SPLIT -- eight entries
// per block, entry c rendering only pixel column c of every 8 x 4 patch (4 live lanes per warp: the lanes' divergent
// candidate walks, which a warp serialises, are spread over eight times as many warps; each of them streams the
// cells again, so only true outliers are worth it) -- then all other blocks in their usual order.  A split block
// reports eight times its slowest warp, which keeps it split while it stays expensive.  Pixels do not depend on the
// order or the split; a frame without a predecessor runs unplanned.
constexpr int RT_PLAN_MAX_SPLIT = 2048;      // blocks that may be split
constexpr unsigned RT_PLAN_HEAVY = 10;       // x the mean cost (measured on the 100 800-triangle frame: 4 .. 32 tried, N = 1 and 8)
// cost[] is indexed by a block's place in the WHOLE frame (16-row block of the frame x block column), so that a
// launch that renders another row range or another interleave class than the last one (the adaptive bands of a
// multi-GPU context move from frame to frame) still finds what is known about its blocks; `first_row`, `il_n`,
// `il_r` locate this launch's block rows in the frame.  A block without history counts as cheap.
__global__ void __launch_bounds__(1024) rt_plan_kernel(const unsigned *__restrict__ cost_frame, int gx, int n_blocks, int first_row,
                                                       int il_n, int il_r, unsigned *__restrict__ plan, unsigned *__restrict__ plan_n,
                                                       unsigned heavy) {
  auto cost_of = [&](int b) -> unsigned { return cost_frame[(size_t)(first_row + (b / gx) * il_n + il_r) * gx + b % gx]; };
  __shared__ unsigned long long s_sum;
  __shared__ unsigned s_heavy, s_base, s_warp[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_sum = 0; s_heavy = 0; }
  __syncthreads();
  unsigned long long sum = 0;
  for (int b = threadIdx.x; b < n_blocks; b += 1024) sum += cost_of(b);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) atomicAdd(&s_sum, sum);
  __syncthreads();
  unsigned long long limit = s_sum / (unsigned long long)n_blocks * heavy + 1ull;
  // how many are heavy?  More than the budget (not a frame with a few outliers): nothing is split
  unsigned mine = 0;
  for (int b = threadIdx.x; b < n_blocks; b += 1024) mine += cost_of(b) > limit ? 1u : 0u;
  if (mine) atomicAdd(&s_heavy, mine);
  __syncthreads();
  if (s_heavy > (unsigned)RT_PLAN_MAX_SPLIT) limit = ~0ull;
  const unsigned n_split = s_heavy > (unsigned)RT_PLAN_MAX_SPLIT ? 0u : s_heavy;
  __syncthreads();
  if (threadIdx.x == 0) { s_heavy = 0; s_base = 8 * n_split; }
  __syncthreads();
  // the heavy blocks, split, at the front
  for (int b = threadIdx.x; b < n_blocks; b += 1024) {
    if (cost_of(b) <= limit) continue;
    const unsigned at = atomicAdd(&s_heavy, 1u);
    const unsigned e = (unsigned)(b % gx) | ((unsigned)(b / gx) << 12) | (1u << 27);
    for (unsigned c = 0; c < 8; ++c) plan[8 * at + c] = e | (c << 24);
  }
  // the others in their usual order
  for (int b0 = 0; b0 < n_blocks; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const bool other = b < n_blocks && cost_of(b) <= limit;
    const unsigned m = __ballot_sync(0xffffffffu, other);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    unsigned before = 0, total = 0;
    for (int w = 0; w < 32; ++w) { before += w < warp ? s_warp[w] : 0u; total += s_warp[w]; }
    if (other) plan[s_base + before + __popc(m & ((1u << lane) - 1))] = (unsigned)(b % gx) | ((unsigned)(b / gx) << 12);
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *plan_n = s_base;
}
