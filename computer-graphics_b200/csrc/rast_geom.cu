// Geometry stage of the rasteriser's Draw on sm_100a
// (rasteriser/Source/skeleton.cpp:205-241): toCameraSpace :701-716,
// createShadowVolume :1676-1722, rotation :223-228, toClipSpace :691-699 and the
// six-plane clip :720-1673, order preserving.
//
// The reference clips the whole list plane after plane, each pass mapping a
// triangle to 0, 1 or 2 triangles and keeping list order.  A triangle's
// descendants never interact with another triangle's, so one thread carries one
// pre-clip triangle through all six planes with a small local list (at most 32
// descendants) and the final list is the concatenation in input order: a count
// pass, an exclusive scan, and a write pass.
//
// The reference spells the clip out as 6 planes x 7 hand-copied cases; they all
// instantiate one pattern, written once here, with the reference's two slips in
// the far plane kept (:1607 tests v2.x instead of v2.w; :1615 divides by
// (w1 - w0) instead of (w1 - w2)).
#include "common.cuh"

struct GeomParams {
  int W, H;
  float focal;
  float cam[4];
  float R[16];
  float light_cam[4];   // sceneCoordinatesLightPos - cameraPos, w = 1 (before rotation)
  const rast_triangle *room;
  int n_room;
  const rast_triangle *boxes;
  int n_boxes;
  unsigned *counts;     // [n_pre + 1]
  unsigned *offs;       // [n_pre + 1] exclusive scan, total at [n_pre]
  rast_triangle *out;
  unsigned out_cap;
  unsigned long long *flags;   // bit0: a triangle with texture != 0, bit1: a shadow-coloured input triangle
  // single-pass mode (pipelined frames): chained scan over the blocks
  unsigned *ticket;            // hands out block ids in scheduling order (forward progress of the look-back)
  unsigned long long *desc;    // per block: state << 62 | value; state 1 = block sum, 2 = inclusive prefix
  unsigned long long *total;   // list length
};

struct GV { float x, y, z, w; };
struct GTri { GV v[3]; };

__device__ __forceinline__ GV gv_lerp(const GV &a, const GV &b, float t) {   // a + t * (b - a)
  GV o;
  o.x = xadd(a.x, xmul(t, xsub(b.x, a.x)));
  o.y = xadd(a.y, xmul(t, xsub(b.y, a.y)));
  o.z = xadd(a.z, xmul(t, xsub(b.z, a.z)));
  o.w = xadd(a.w, xmul(t, xsub(b.w, a.w)));
  return o;
}

// glm mat4 * vec4 (glm/glm/detail/type_mat4x4.inl:640-652)
__device__ __forceinline__ GV gv_rotate(const float *R, const GV &v) {
  GV o;
  o.x = xadd(xadd(xmul(R[0], v.x), xmul(R[4], v.y)), xadd(xmul(R[8], v.z), xmul(R[12], v.w)));
  o.y = xadd(xadd(xmul(R[1], v.x), xmul(R[5], v.y)), xadd(xmul(R[9], v.z), xmul(R[13], v.w)));
  o.z = xadd(xadd(xmul(R[2], v.x), xmul(R[6], v.y)), xadd(xmul(R[10], v.z), xmul(R[14], v.w)));
  o.w = xadd(xadd(xmul(R[3], v.x), xmul(R[7], v.y)), xadd(xmul(R[11], v.z), xmul(R[15], v.w)));
  return o;
}

struct ClipCtx { int plane, W, H; float focal; };

__device__ __forceinline__ float clip_coord(const ClipCtx &c, const GV &v) {
  return c.plane <= 2 ? v.x : (c.plane <= 4 ? v.y : v.w);
}
// the plane's limit for this vertex: `dot[k]` (:732, :922, :1115, :1307) or wlimit (:1509)
__device__ __forceinline__ float clip_limit(const ClipCtx &c, const GV &v) {
  switch (c.plane) {
    case 1: return xdiv(xmul(v.w, (float)(-c.W)), 2.0f);
    case 2: return xdiv(xmul(v.w, (float)(c.W)), 2.0f);
    case 3: return xdiv(xmul(v.w, (float)(c.H)), 2.0f);
    case 4: return xdiv(xmul(v.w, (float)(-c.H)), 2.0f);
    default: return xdiv(5.0f, c.focal);
  }
}
__device__ __forceinline__ bool clip_in(const ClipCtx &c, const GV &v) {
  const float a = clip_coord(c, v), l = clip_limit(c, v);
  return (c.plane == 1 || c.plane == 4) ? a > l : (c.plane == 6 ? a <= l : a < l);
}
__device__ __forceinline__ bool clip_out(const ClipCtx &c, const GV &v) {
  const float a = clip_coord(c, v), l = clip_limit(c, v);
  return (c.plane == 1 || c.plane == 4) ? a <= l : (c.plane == 6 ? a > l : a >= l);
}
// intersection parameter along a (inside) -> b (outside), e.g. t_01 at :753 / :943 / :1524
__device__ __forceinline__ float clip_t(const ClipCtx &c, const GV &a, const GV &b) {
  if (c.plane == 6) return xdiv(xsub(xdiv(5.0f, c.focal), a.w), xsub(b.w, a.w));
  const int S = c.plane <= 2 ? c.W : c.H;
  const float h = (float)(S / 2), nh = (float)((-S) / 2);
  const float ca = c.plane <= 2 ? a.x : a.y, cb = c.plane <= 2 ? b.x : b.y;
  if (c.plane == 1 || c.plane == 4)
    return xdiv(xadd(ca, xmul(h, a.w)), xadd(xsub(xadd(xmul(nh, b.w), xmul(h, a.w)), cb), ca));
  return xdiv(xsub(ca, xmul(h, a.w)), xadd(xsub(xsub(xmul(h, b.w), xmul(h, a.w)), cb), ca));
}

// One triangle through one plane: 0, 1 or 2 results.
__device__ __forceinline__ int clip_one(const ClipCtx &c, const GTri &in, GTri &o0, GTri &o1) {
  const GV &a = in.v[0], &b = in.v[1], &d = in.v[2];
  if (c.plane == 5) {                                           // :1497-1505: dropped, never split
    if (a.z > 0.01f && b.z > 0.01f && d.z > 0.01f) { o0 = in; return 1; }
    return 0;
  }
  const bool i0 = clip_in(c, a), i1 = clip_in(c, b), i2 = clip_in(c, d);
  const bool x0 = clip_out(c, a), x1 = clip_out(c, b), x2 = clip_out(c, d);
  o0 = in;
  if (i0 && i1 && i2) return 1;
  if (i0 && x1 && x2) { o0.v[1] = gv_lerp(a, b, clip_t(c, a, b)); o0.v[2] = gv_lerp(a, d, clip_t(c, a, d)); return 1; }
  if (x0 && i1 && x2) { o0.v[0] = gv_lerp(b, a, clip_t(c, b, a)); o0.v[2] = gv_lerp(b, d, clip_t(c, b, d)); return 1; }
  if (x0 && x1 && i2) { o0.v[1] = gv_lerp(d, b, clip_t(c, d, b)); o0.v[0] = gv_lerp(d, a, clip_t(c, d, a)); return 1; }
  if (i0 && i1 && x2) {                                         // :814-846
    const GV np12 = gv_lerp(b, d, clip_t(c, b, d)), np02 = gv_lerp(a, d, clip_t(c, a, d));
    o0.v[2] = np02;
    o1.v[0] = np02; o1.v[1] = np12; o1.v[2] = b;
    return 2;
  }
  const bool c02 = c.plane == 6 ? (i0 && x1 && d.x <= xdiv(5.0f, c.focal)) : (i0 && x1 && i2);   // :1607
  if (c02) {                                                    // :849-881
    const float t01 = clip_t(c, a, b);
    const float t21 = c.plane == 6 ? xdiv(xsub(xdiv(5.0f, c.focal), d.w), xsub(b.w, a.w))        // :1615
                                   : clip_t(c, d, b);
    const GV np01 = gv_lerp(a, b, t01), np21 = gv_lerp(d, b, t21);
    o0.v[1] = np01;
    o1.v[0] = np01; o1.v[1] = np21; o1.v[2] = d;
    return 2;
  }
  if (x0 && i1 && i2) {                                         // :883-916
    const GV np10 = gv_lerp(b, a, clip_t(c, b, a)), np20 = gv_lerp(d, a, clip_t(c, d, a));
    o0.v[0] = np10;
    o1.v[0] = np10; o1.v[1] = np20; o1.v[2] = d;
    return 2;
  }
  return 0;   // NaNs, or the far plane's unmatched combination: dropped
}

// The general case: the triangle's descendants through the six planes, in the
// reference's list order.  cur[0] holds the input; returns the count.
__device__ __noinline__ int clip_six_planes(int W, int H, float focal, GTri *cur) {
  GTri nxt[32];
  int n_cur = 1;
  for (int plane = 1; plane <= 6; ++plane) {
    ClipCtx c;
    c.plane = plane; c.W = W; c.H = H; c.focal = focal;
    int n_nxt = 0;
    for (int i = 0; i < n_cur; ++i) {
      GTri o0, o1;
      const int m = clip_one(c, cur[i], o0, o1);
      if (m >= 1) nxt[n_nxt++] = o0;
      if (m == 2) nxt[n_nxt++] = o1;
    }
    for (int i = 0; i < n_nxt; ++i) cur[i] = nxt[i];
    n_cur = n_nxt;
    if (n_cur == 0) break;
  }
  return n_cur;
}

// MODE 0: count the triangles each pre-clip triangle turns into; MODE 1: write them at the
// offsets a scan of the counts gave (exact sizes: two passes and a host read-back in between);
// MODE 2: one pass -- count, block scan, decoupled look-back over the preceding blocks'
// published sums, write -- into a list whose capacity was guessed (pipelined frames).
template <int MODE>
__global__ void __launch_bounds__(128) rast_geom_kernel(const __grid_constant__ GeomParams p) {
  constexpr bool WRITE = MODE != 0;
  constexpr int GT = 128, TW = sizeof(rast_triangle) / 4;
  __shared__ uint32_t stage[GT * TW];   // coalesced word streams in and out of the 84-byte records
  __shared__ unsigned s_bid, s_warp[4], s_prefix;
  if (MODE == 2) {
    if (threadIdx.x == 0) s_bid = atomicAdd(p.ticket, 1u);
    __syncthreads();
  }
  const int bid = MODE == 2 ? (int)s_bid : (int)blockIdx.x;
  const int j0 = bid * GT;
  const int j = j0 + threadIdx.x;
  const int n_pre = p.n_room + 7 * p.n_boxes;
  // a block whose triangles all come from `room` reads them through shared memory
  const bool staged_in = j0 + GT <= p.n_room;
  if (staged_in) {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(p.room + j0);
    for (int i = threadIdx.x; i < GT * TW; i += GT) stage[i] = __ldg(src + i);
    __syncthreads();
  }
  // ---- the pre-clip triangle j (Draw :205-220) ----
  const rast_triangle *srcp;
  int s = 0;
  const int jj = j < n_pre ? j : n_pre - 1;   // tail threads repeat the last triangle, their output is dropped
  if (staged_in) srcp = reinterpret_cast<const rast_triangle *>(stage + threadIdx.x * TW);
  else if (jj < p.n_room) srcp = p.room + jj;
  else { srcp = p.boxes + (jj - p.n_room) / 7; s = (jj - p.n_room) % 7; }
  GTri t;
  float attr[9];   // normal[4], color[3], texture, index (bit-cast)
  {
    GV o[3];
    const float *vin[3] = {srcp->v0, srcp->v1, srcp->v2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {                                // toCameraSpace :701-711
      o[k].x = xsub(vin[k][0], p.cam[0]); o[k].y = xsub(vin[k][1], p.cam[1]); o[k].z = xsub(vin[k][2], p.cam[2]);
      o[k].w = 1.0f;
    }
    if (s == 0) {
      t.v[0] = o[0]; t.v[1] = o[1]; t.v[2] = o[2];
#pragma unroll
      for (int k = 0; k < 4; ++k) attr[k] = srcp->normal[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) attr[4 + k] = srcp->color[k];
      attr[7] = __int_as_float(srcp->texture);
      attr[8] = __int_as_float(srcp->index);
    } else {
      GV n[3];                                                   // createShadowVolume :1695-1697
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        n[k].x = xmul(xsub(o[k].x, p.light_cam[0]), 100.0f);
        n[k].y = xmul(xsub(o[k].y, p.light_cam[1]), 100.0f);
        n[k].z = xmul(xsub(o[k].z, p.light_cam[2]), 100.0f);
        n[k].w = xmul(xsub(o[k].w, p.light_cam[3]), 100.0f);
      }
      const int e = (s - 1) >> 1, e1 = (e + 1) % 3;              // :1705-1710
      if ((s - 1) & 1) { t.v[0] = n[e]; t.v[1] = o[e1]; t.v[2] = n[e1]; }
      else { t.v[0] = o[e]; t.v[1] = n[e]; t.v[2] = o[e1]; }
      // Triangle::ComputeNormal (rasteriser/Source/TestModelH.h:32-41)
      const float ax = xsub(t.v[1].x, t.v[0].x), ay = xsub(t.v[1].y, t.v[0].y), az = xsub(t.v[1].z, t.v[0].z);
      const float bx = xsub(t.v[2].x, t.v[0].x), by = xsub(t.v[2].y, t.v[0].y), bz = xsub(t.v[2].z, t.v[0].z);
      const float cx = xsub(xmul(by, az), xmul(ay, bz)), cy = xsub(xmul(bz, ax), xmul(az, bx)),
                  cz = xsub(xmul(bx, ay), xmul(ax, by));
      const float inv = xdiv(1.0f, xsqrt(xdot3(cx, cy, cz, cx, cy, cz)));
      attr[0] = xmul(cx, inv); attr[1] = xmul(cy, inv); attr[2] = xmul(cz, inv); attr[3] = 1.0f;
      attr[4] = attr[5] = attr[6] = -1.0f;
      attr[7] = __int_as_float(0);
      attr[8] = __int_as_float(0);   // uninitialised in the reference; never read for texture 0
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    t.v[k] = gv_rotate(p.R, t.v[k]);                             // :224-228
    t.v[k].w = xdiv(t.v[k].z, p.focal);                          // :695-697
  }
  // ---- six planes, list order preserved (:236-241) ----
  // Fast path: a triangle inside all six planes comes out unchanged.
  GTri cur[32];
  int n_cur = 1;
  bool all_in = true;
#pragma unroll
  for (int plane = 1; plane <= 6; ++plane) {
    ClipCtx c;
    c.plane = plane; c.W = p.W; c.H = p.H; c.focal = p.focal;
    if (plane == 5) all_in = all_in && t.v[0].z > 0.01f && t.v[1].z > 0.01f && t.v[2].z > 0.01f;
    else all_in = all_in && clip_in(c, t.v[0]) && clip_in(c, t.v[1]) && clip_in(c, t.v[2]);
  }
  cur[0] = t;
  if (!all_in) n_cur = clip_six_planes(p.W, p.H, p.focal, cur);
  if (MODE != 1 && j < n_pre) {
    if (__float_as_int(attr[7]) != 0) atomicOr(p.flags, 1ull);
    if (s == 0 && !(attr[4] >= 0.0f)) atomicOr(p.flags, 2ull);
  }
  if (MODE == 0) {
    if (j < n_pre) p.counts[j] = (unsigned)n_cur;
    return;
  }
  unsigned my_off = 0;   // where this thread's first output goes
  if (MODE == 2) {
    const unsigned cnt = j < n_pre ? (unsigned)n_cur : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned before = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) { before += w < warp ? s_warp[w] : 0u; block_sum += s_warp[w]; }
    if (warp == 0) {
      // Decoupled look-back by the whole warp: lane l inspects the block l places back; the
      // window moves 32 blocks at a time until it contains a block whose inclusive prefix is known.
      // Every block before this one already holds its ticket, so it will publish.
      const unsigned long long VAL = (1ull << 62) - 1;
      if (lane == 0 && bid > 0) atomicExch(p.desc + bid, (1ull << 62) | block_sum);
      unsigned long long prefix = 0;
      for (int base = bid - 1;;) {
        const int k = base - lane;
        const unsigned long long v = k >= 0 ? atomicAdd(p.desc + k, 0ull) : (2ull << 62);   // "block -1": prefix 0
        const unsigned state = (unsigned)(v >> 62);
        const unsigned ready = __ballot_sync(0xffffffffu, state != 0), done = __ballot_sync(0xffffffffu, state == 2);
        const int first = done ? __ffs(done) - 1 : 31;                 // nearest block with an inclusive prefix
        const unsigned need = first == 31 ? 0xffffffffu : ((2u << first) - 1u);
        if ((ready & need) != need) continue;                            // someone in the window has not published yet
        unsigned long long part = lane <= first ? (v & VAL) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        prefix += part;
        if (done) break;
        base -= 32;
      }
      if (lane == 0) {
        atomicExch(p.desc + bid, (2ull << 62) | (prefix + block_sum));
        if (bid == (int)gridDim.x - 1) *p.total = prefix + block_sum;
        s_prefix = (unsigned)prefix;
      }
    }
    __syncthreads();
    my_off = s_prefix + before + incl - cnt;
  }
  // Common case: the whole block is unclipped (one output each, contiguous in the
  // list): write the records through shared memory as one coalesced stream.
  if (__syncthreads_and(staged_in && n_cur == 1)) {
    float *o = reinterpret_cast<float *>(stage) + threadIdx.x * TW;
    const GTri &g = cur[0];
    o[0] = g.v[0].x; o[1] = g.v[0].y; o[2] = g.v[0].z; o[3] = g.v[0].w;
    o[4] = g.v[1].x; o[5] = g.v[1].y; o[6] = g.v[1].z; o[7] = g.v[1].w;
    o[8] = g.v[2].x; o[9] = g.v[2].y; o[10] = g.v[2].z; o[11] = g.v[2].w;
#pragma unroll
    for (int k = 0; k < 9; ++k) o[12 + k] = attr[k];
    __syncthreads();
    const unsigned off0 = MODE == 2 ? s_prefix : p.offs[j0];
    uint32_t *dst = reinterpret_cast<uint32_t *>(p.out + off0);
    if (off0 + GT <= p.out_cap)
      for (int i = threadIdx.x; i < GT * TW; i += GT) dst[i] = stage[i];
    return;
  }
  if (j >= n_pre) return;
  const unsigned off = MODE == 2 ? my_off : p.offs[j];
  for (int i = 0; i < n_cur; ++i) {
    if (off + i >= p.out_cap) break;
    rast_triangle *o = p.out + off + i;
    const GTri &g = cur[i];
    o->v0[0] = g.v[0].x; o->v0[1] = g.v[0].y; o->v0[2] = g.v[0].z; o->v0[3] = g.v[0].w;
    o->v1[0] = g.v[1].x; o->v1[1] = g.v[1].y; o->v1[2] = g.v[1].z; o->v1[3] = g.v[1].w;
    o->v2[0] = g.v[2].x; o->v2[1] = g.v[2].y; o->v2[2] = g.v[2].z; o->v2[3] = g.v[2].w;
#pragma unroll
    for (int k = 0; k < 4; ++k) o->normal[k] = attr[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) o->color[k] = attr[4 + k];
    o->texture = __float_as_int(attr[7]);
    o->index = __float_as_int(attr[8]);
  }
}

// Exclusive scan of counts[0..n) into offs[0..n], total at offs[n]: per-block
// sums, a one-block scan of those, then per-block scans with their offsets.
constexpr int SCAN_THREADS = 1024, SCAN_ITEMS = 4, SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

// Exclusive prefix of `v` over the block (all SCAN_THREADS threads call); total in *total.
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned *total) {
  __shared__ unsigned warp_excl[32];
  __shared__ unsigned block_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned m = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += m;
  }
  __syncthreads();   // protects warp_excl / block_total across back-to-back calls
  if (lane == 31) warp_excl[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const unsigned w = lane < nw ? warp_excl[lane] : 0u;
    unsigned wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned m = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += m;
    }
    warp_excl[lane] = wi - w;
    if (lane == 31) block_total = wi;
  }
  __syncthreads();
  *total = block_total;
  return warp_excl[warp] + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const unsigned *__restrict__ counts, int n,
                                                                   unsigned *__restrict__ block_sums) {
  const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  unsigned s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) s += base + k < n ? counts[base + k] : 0u;
  unsigned total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// In-place exclusive scan of block_sums[0..nb), grand total at block_sums[nb]; one block.
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(unsigned *__restrict__ block_sums, int nb) {
  unsigned carry = 0;
  for (int base = 0; base < nb; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const unsigned v = i < nb ? block_sums[i] : 0u;
    unsigned total;
    const unsigned ex = block_excl_scan(v, &total);
    if (i < nb) block_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) block_sums[nb] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const unsigned *__restrict__ counts, int n,
                                                                  const unsigned *__restrict__ block_sums, int nb,
                                                                  unsigned *__restrict__ offs,
                                                                  unsigned long long *__restrict__ total_out) {
  const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  unsigned v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = base + k < n ? counts[base + k] : 0u; s += v[k]; }
  unsigned total;
  unsigned run = block_sums[blockIdx.x] + block_excl_scan(s, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) offs[base + k] = run;
    run += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    offs[n] = block_sums[nb];
    if (total_out) *total_out = block_sums[nb];
  }
}

// offs[0..n] = exclusive scan of counts[0..n); tmp holds nb + 1 block sums.
int scan_exclusive(b200_ctx *ctx, const unsigned *counts, unsigned *offs, int n, unsigned *tmp,
                   unsigned long long *total_out) {
  const int nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  scan_reduce_kernel<<<nb, SCAN_THREADS, 0, ctx->stream>>>(counts, n, tmp);
  scan_sums_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(tmp, nb);
  scan_apply_kernel<<<nb, SCAN_THREADS, 0, ctx->stream>>>(counts, n, tmp, nb, offs, total_out);
  ctx->stats.kernel_launches += 3;
  tl_mark(ctx, "scan_apply_kernel");
  CU_CHECK(ctx, cudaGetLastError());
  return B200_OK;
}

// Host: runs the stage on the uploaded world-space scene; leaves the clipped list
// in ctx->rast_src / ctx->rast_n_tris and the camera-space rotated light in
// light_out (what the triangle loop's calculateIllumination reads, :675).
int rast_geometry(b200_ctx *ctx, const camera_t *cam, const rast_light_t *light, rast_light_t *light_out,
                  bool spec) {
  const int n_room = ctx->rast_n_room, n_boxes = ctx->rast_n_boxes;
  const int n_pre = n_room + 7 * n_boxes;
  GeomParams p;
  memset(&p, 0, sizeof p);
  p.W = cam->width; p.H = cam->height; p.focal = cam->focal;
  memcpy(p.cam, cam->pos, sizeof p.cam);
  memcpy(p.R, cam->R, sizeof p.R);
  // lightPos = sceneCoordinatesLightPos - cameraPos, w = 1 (:211-212, :713-716); then R * lightPos (:223)
  volatile float lx = light->pos[0] - cam->pos[0], ly = light->pos[1] - cam->pos[1], lz = light->pos[2] - cam->pos[2];
  p.light_cam[0] = lx; p.light_cam[1] = ly; p.light_cam[2] = lz; p.light_cam[3] = 1.0f;
  *light_out = *light;
  for (int r = 0; r < 4; ++r) {
    volatile float a = cam->R[0 + r] * p.light_cam[0], b = cam->R[4 + r] * p.light_cam[1];
    volatile float c = cam->R[8 + r] * p.light_cam[2], d = cam->R[12 + r] * p.light_cam[3];
    volatile float ab = a + b, cd = c + d;
    light_out->pos[r] = ab + cd;
  }
  p.room = (const rast_triangle *)ctx->rast_world.p;
  p.n_room = n_room;
  p.boxes = p.room + n_room;
  p.n_boxes = n_boxes;
  const int n_scan_blocks = (n_pre + SCAN_BLOCK - 1) / SCAN_BLOCK;
  const int n_blocks = (n_pre + 127) / 128;
  const size_t tmp_words = 2 * (size_t)(n_pre + 1) + n_scan_blocks + 2;          // counts, offsets, scan scratch
  const size_t chain_off = (tmp_words * sizeof(unsigned) + 15) / 16 * 16;        // then: ticket (16 B), block descriptors
  if (int rc = ensure(ctx, ctx->rast_geom_tmp, chain_off + 16 + sizeof(unsigned long long) * (size_t)(n_blocks + 1))) return rc;
  p.flags = (unsigned long long *)ctx->counters.p + 7;
  p.counts = (unsigned *)ctx->rast_geom_tmp.p;
  p.offs = p.counts + (n_pre + 1);
  p.ticket = (unsigned *)((char *)ctx->rast_geom_tmp.p + chain_off);
  p.desc = (unsigned long long *)((char *)ctx->rast_geom_tmp.p + chain_off + 16);
  p.total = (unsigned long long *)ctx->counters.p + 24;
  if (spec && n_pre > 0) {
    // pipelined: one pass into a list whose capacity comes from the last verified frame; the
    // list length stays on the device ([24]) and the guess is checked when the frame is waited for
    unsigned long long cap = rast_spec_cap(ctx->rast_spec.tris);
    if (cap > 0x3fffffffull) cap = 0x3fffffffull;
    if (int rc = ensure(ctx, ctx->rast_src, sizeof(rast_triangle) * (size_t)cap)) return rc;
    p.out = (rast_triangle *)ctx->rast_src.p;
    p.out_cap = (unsigned)cap;
    CU_CHECK(ctx, cudaMemsetAsync(p.ticket, 0, 16 + sizeof(unsigned long long) * (size_t)n_blocks, ctx->stream));
    rast_geom_kernel<2><<<n_blocks, 128, 0, ctx->stream>>>(p);
    ctx->stats.kernel_launches++;
    tl_mark(ctx, "rast_geom_kernel<2>");
    CU_CHECK(ctx, cudaGetLastError());
    ctx->rast_n_tris = (int)cap;
    return B200_OK;
  }
  unsigned total = 0;
  if (n_pre > 0) {
    rast_geom_kernel<0><<<(n_pre + 127) / 128, 128, 0, ctx->stream>>>(p);
    ctx->stats.kernel_launches++;
    tl_mark(ctx, "rast_geom_kernel");
    CU_CHECK(ctx, cudaGetLastError());
    unsigned long long *dc = (unsigned long long *)ctx->counters.p;
    if (int rc = scan_exclusive(ctx, p.counts, p.offs, n_pre, p.offs + (n_pre + 1), dc + 24)) return rc;
    {
      // one small read-back into pinned memory: [7] validation flags, [24] clipped-list length
      unsigned long long *hc = (unsigned long long *)ctx->pinned;
      CU_CHECK(ctx, cudaMemcpyAsync(hc, dc, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
      CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
      const unsigned long long flags = hc[7];
      total = (unsigned)hc[24];
      if (flags & 1ull) return ctx_fail(ctx, B200_EINVAL, "only texture == 0 is supported");
      ctx->rast_has_shadow = (n_boxes > 0 || (flags & 2ull)) ? 1 : 0;
    }
  }
  if (int rc = ensure(ctx, ctx->rast_src, sizeof(rast_triangle) * (size_t)(total ? total : 1))) return rc;
  p.out = (rast_triangle *)ctx->rast_src.p;
  p.out_cap = total;
  if (total > 0) {
    rast_geom_kernel<1><<<(n_pre + 127) / 128, 128, 0, ctx->stream>>>(p);
    ctx->stats.kernel_launches++;
    tl_mark(ctx, "rast_geom_kernel");
    CU_CHECK(ctx, cudaGetLastError());
  }
  ctx->rast_n_tris = (int)total;
  return B200_OK;
}
