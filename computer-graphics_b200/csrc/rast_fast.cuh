// Fast path of the rasteriser for lists WITHOUT shadow-volume triangles.
//
// With only opaque triangles the reference's fold (`zinv >= depth`, drawn in
// list order, rasteriser/Source/skeleton.cpp:574,665) is order independent: the
// pixel ends up owned by the fragment with the largest zinv, the LATEST triangle
// on ties.  That is a lexicographic maximum of (zinv, index), so fragments can be
// scattered in any order with one 64-bit atomicMax per fragment on
//   key = zinv bits << 32 | (triangle index + 1)          (0 = empty pixel)
// (non-negative floats order like their bit patterns; fragments with zinv < 0 or
// NaN fail `zinv >= depth` against the cleared buffer and are never drawn).
//
//   rast_scatter_kernel  one thread per (triangle,row): the row record (span ends,
//                        zinv and position steps) is derived and stored, then one
//                        atomicMax per pixel of the span
//   rast_resolve_kernel  per 32x8 pixel tile + 1-pixel halo: the winners are
//                        compacted, each winning fragment is re-derived exactly
//                        from its row record (fragment, three
//                        calculateIllumination) into shared memory, then the 5-tap
//                        AA / HDR mean of the post pass (:283-307; no shadow flags
//                        => nothing is darkened) is taken from shared memory and
//                        the frame written once.
// No per-pixel colour buffers and no tile lists touch HBM.
#pragma once

__device__ __forceinline__ unsigned long long rast_key(float zinv, int tri) {
  const unsigned hi = zinv == 0.0f ? 0u : __float_as_uint(zinv);   // -0.0 passes `>= 0` too
  return ((unsigned long long)hi << 32) | (unsigned)(tri + 1);
}

// red.max on a key with an L2 evict-last policy: the 64-bit keys (66 MB at 4K) are what the
// scatter revisits at random, while the setup records it reads and the row records it writes
// stream through once -- the policy keeps the keys resident while those pass.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void red_max_u64_keep(unsigned long long *addr, unsigned long long v, uint64_t pol) {
  asm volatile("red.relaxed.gpu.global.max.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(addr), "l"(v), "l"(pol) : "memory");
}

__global__ void rast_scatter_kernel(const __grid_constant__ RastParams p) {
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned chunk = gid >> RAST_CHUNK_LOG2, sub = gid & (RAST_CHUNK - 1);
  unsigned long long n_frag = 0;
  if (chunk < rast_count_chunks(p)) {
    const int t = (int)min((unsigned)p.chunk_owner[chunk], (unsigned)(p.n_tris - 1));   // stale owner after an overflow
    const RastSetup &s = p.setup[t];
    const int r = (int)((chunk - s.chunk_off) << RAST_CHUNK_LOG2) + (int)sub;
    if (r >= 0 && r < s.nrows) {
      const int y = s.row0 + r;
      float4 A, B;
      rast_row_record<true>(s, y, A, B);
      __stcs(p.rowsA + s.row_off + r, A);   // kept for the resolve pass (the winner's fragment is re-derived
      __stcs(p.rowsB + s.row_off + r, B);   // from its row record); streaming stores: the keys should stay in L2
      const int lx = __float_as_int(A.x), rx = __float_as_int(A.y);
      const int x0 = max(lx, 0), x1 = min(rx, p.W);      // right end excluded (:504); bounds (:573)
      unsigned long long *row = p.keys + (size_t)y * p.W;
      const uint64_t keep = l2_policy_evict_last();
      for (int x = x0; x < x1; ++x) {
        const float zinv = xadd(A.z, xmul(A.w, (float)(x - lx)));   // :543
        if (zinv >= 0.0f) red_max_u64_keep(row + x, rast_key(zinv, t), keep);   // :574 against the cleared buffer
      }
      n_frag = x1 > x0 ? (unsigned long long)(x1 - x0) : 0ull;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_frag += __shfl_xor_sync(0xffffffffu, n_frag, o);
  if ((threadIdx.x & 31) == 0 && n_frag) atomicAdd(p.counters + 16, n_frag);
}

constexpr int RS_W = 32, RS_H = 8, RS_HW = RS_W + 2, RS_HH = RS_H + 2, RS_N = RS_HW * RS_HH;   // 34 x 10 = 340

__global__ void __launch_bounds__(RS_W * RS_H) rast_resolve_kernel(const __grid_constant__ RastParams p) {
  __shared__ float col[RS_N][10];           // screen rgb, low rgb, high rgb, depth
  __shared__ int owner[RS_N];
  __shared__ unsigned short work[RS_N];
  __shared__ int warp_base[RS_W * RS_H / 32 + 1];
  __shared__ int n_work;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * RS_W - 1, y0 = p.row0 + blockIdx.y * RS_H - 1;   // halo origin

  // ---- gather the winners of the tile + halo, compact the covered ones ----
  if (tid == 0) n_work = 0;
  __syncthreads();
  for (int base = 0; base < RS_N; base += RS_W * RS_H) {
    const int pos = base + tid;
    unsigned long long key = 0;
    if (pos < RS_N) {
      const int gx = x0 + pos % RS_HW, gy = y0 + pos / RS_HW;
      if (gx >= 0 && gx < p.W && gy >= p.fb0 && gy < p.fb1) key = p.keys[(size_t)gy * p.W + gx];
      owner[pos] = (int)(unsigned)(key & 0xffffffffull) - 1;
#pragma unroll
      for (int k = 0; k < 10; ++k) col[pos][k] = 0.f;
    }
    const bool covered = key != 0;
    const unsigned m = __ballot_sync(0xffffffffu, covered);
    if (lane == 0) warp_base[warp] = __popc(m);
    __syncthreads();
    if (tid == 0) {
      int acc = n_work;
      for (int w = 0; w < RS_W * RS_H / 32; ++w) { const int c = warp_base[w]; warp_base[w] = acc; acc += c; }
      n_work = acc;
    }
    __syncthreads();
    if (covered) work[warp_base[warp] + __popc(m & ((1u << lane) - 1))] = (unsigned short)pos;
    __syncthreads();
  }

  // ---- deferred PixelShader of every winner (:575-586) ----
  for (int k = tid; k < n_work; k += RS_W * RS_H) {
    const int pos = work[k];
    const int gx = x0 + pos % RS_HW;
    const int t = owner[pos];
    const RastSetup &s = p.setup[t];
    const unsigned rr = s.row_off + (unsigned)(y0 + pos / RS_HW - s.row0);
    const float4 A = p.rowsA[rr], B = p.rowsB[rr];
    const float fi = (float)(gx - __float_as_int(A.x));
    const float zinv = xadd(A.z, xmul(A.w, fi));
    const float pz = xdiv(1.0f, zinv);                              // :546
    const float px = xdiv(xadd(B.x, xmul(B.y, fi)), zinv);          // :547
    const float py = xdiv(xadd(B.z, xmul(B.w, fi)), zinv);          // :548
    const rast_triangle *tr = p.src + t;
    float D[3];
    rast_illum_D(p, px, py, pz, tr->normal[0], tr->normal[1], tr->normal[2], D);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float cc = tr->color[c];
      col[pos][c] = xmul(cc, xadd(D[c], rast_indirect(p, c, t, (size_t)(y0 + pos / RS_HW) * p.W + gx)));   // :580
      col[pos][3 + c] = xmul(cc, xadd(D[c], 0.0f));                 // :581-582
      col[pos][6 + c] = xmul(cc, xadd(D[c], 0.4f));                 // :583-584
    }
    col[pos][9] = zinv;                                             // :665
  }
  __syncthreads();

  // ---- post pass on the core pixels (:283-307, antiAliasing :1736-1753) ----
  const int tx = tid % RS_W, ty = tid / RS_W;
  const int x = x0 + 1 + tx, y = y0 + 1 + ty;
  if (x >= p.W || y >= p.row1) return;
  const int c0 = (ty + 1) * RS_HW + (tx + 1);
  const size_t q = (size_t)y * p.W + x;
  float out[3] = {0.f, 0.f, 0.f};
  const bool interior = y >= 1 && y <= p.H - 2 && x >= 1 && x <= p.W - 2;
  // five empty taps average to +0 exactly: nothing to compute (most pixels of a sparse scene)
  const bool any = (owner[c0] & owner[c0 - RS_HW] & owner[c0 + RS_HW] & owner[c0 - 1] & owner[c0 + 1]) >= 0;
  if (interior && any) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float acc[3];
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int k = 3 * b + c;
        float a = xadd(col[c0][k], col[c0 - RS_HW][k]);
        a = xadd(a, col[c0 + RS_HW][k]);
        a = xadd(a, col[c0 - 1][k]);
        a = xadd(a, col[c0 + 1][k]);
        acc[b] = xdiv_const<5>(a);
      }
      out[c] = xdiv_const<3>(xadd(xadd(acc[0], acc[1]), acc[2]));     // :1750
    }
  }
  if (p.out_rgb) { p.out_rgb[3 * q] = out[0]; p.out_rgb[3 * q + 1] = out[1]; p.out_rgb[3 * q + 2] = out[2]; }
  if (p.out_argb) p.out_argb[q] = interior ? put_pixel_argb(out[0], out[1], out[2]) : 0u;
  if (p.out_depth) p.out_depth[q] = col[c0][9];
  if (p.out_index) p.out_index[q] = owner[c0];
}
