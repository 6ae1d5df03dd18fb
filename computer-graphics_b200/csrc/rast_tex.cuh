// Texture branches of the rasteriser's PixelShader (rasteriser/Source/skeleton.cpp:588-645) and
// findU / findV (:1756-1825), in the reference's operation order (one IEEE rounding per operation).
//
// The images are what the reference's main() holds in its cv::Mat globals (:63-75, :133-155): byte
// images addressed like cv::Mat::at, at<T>(row, col) = *(T *)(data + row * step + col * sizeof(T)), row =
// findU, col = findV.  The caller decodes the files; the library only reads pixels (rast_set_textures).
//
//   texture 1  marble     colour = marble BGR / 255, normal = triangle normal + normalMap_marble[y * rows + x]
//   texture 2  metal grill  a fragment whose thresholded opacity is not 255 is a HOLE: it passes the depth
//   texture 3  woven wood   test, leaves the colours alone and sets depthBuffer to 0 (:619, :643, :665);
//                           otherwise colour = base BGR / 255, normal = normalize(vec4(map BGR / 255, 1)),
//                           woven wood also scales the illumination by occlusion / 255 (:626-638)
#pragma once
#include "common.cuh"

constexpr int RAST_TEX_MARBLE = 0, RAST_TEX_GRILL = 1, RAST_TEX_GRILL_OPACITY = 2, RAST_TEX_GRILL_NORMAL = 3,
              RAST_TEX_WOVEN = 4, RAST_TEX_WOVEN_OCCLUSION = 5, RAST_TEX_WOVEN_OPACITY = 6, RAST_TEX_WOVEN_NORMAL = 7;

// float -> int as the reference build converts (cvttss2si: NaN and out-of-range give INT_MIN)
__device__ __forceinline__ int rast_tex_to_int(float v) {
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT_MIN;
  return (int)v;
}

// the common head of findU / findV (:1759-1769): the fragment's position back in scene coordinates
__device__ __forceinline__ void rast_tex_object_space(const RastTex &t, float px, float py, float pz, float *o) {
  if (t.use_rinv) {
    // glm mat4 * vec4: (m[0] * x + m[1] * y) + (m[2] * z + m[3] * w), w = 1
#pragma unroll
    for (int r = 0; r < 3; ++r)
      o[r] = xadd(xadd(xadd(xmul(t.Rinv[r], px), xmul(t.Rinv[4 + r], py)), xadd(xmul(t.Rinv[8 + r], pz), xmul(t.Rinv[12 + r], 1.0f))), t.cam[r]);
  } else {
    o[0] = xadd(px, t.cam[0]); o[1] = xadd(py, t.cam[1]); o[2] = xadd(pz, t.cam[2]);
  }
}

// findU (:1771-1790) and findV (:1807-1824) for a texture of `size` on the object `index`; a negative
// remainder (outside the unit box the textures are laid over: an out-of-bounds read in the reference)
// wraps into the image
__device__ __forceinline__ void rast_tex_uv(const RastTex &t, float px, float py, float pz, int size, int index, int &u, int &v) {
  float o[3];
  rast_tex_object_space(t, px, py, pz, o);
  const float mh = (float)(-size / 2), ph = (float)(size / 2);
  u = 0; v = 0;
  if (index == 3) { u = rast_tex_to_int(xadd(xmul(mh, o[1]), ph)); v = rast_tex_to_int(xadd(xmul(ph, o[2]), ph)); }
  else if (index == 1 || index == 2) { u = rast_tex_to_int(xadd(xmul(mh, o[0]), ph)); v = rast_tex_to_int(xadd(xmul(mh, o[2]), ph)); }
  else if (index == 4) { u = rast_tex_to_int(xadd(xmul(mh, o[1]), ph)); v = rast_tex_to_int(xadd(xmul(mh, o[2]), ph)); }
  else if (index == 0) { u = rast_tex_to_int(xadd(xmul(mh, o[0]), ph)); v = rast_tex_to_int(xadd(xmul(mh, o[1]), ph)); }
  u %= size; v %= size;
  if (u < 0) u += size;
  if (v < 0) v += size;
}

__device__ __forceinline__ const unsigned char *rast_texel(const RastTex &t, int which, int u, int v, int elem) {
  return t.img[which] + (size_t)u * t.step[which] + (size_t)v * elem;
}

// (:603, :625) is this fragment of a metal-grill / woven-wood triangle a hole?
__device__ __forceinline__ bool rast_tex_hole(const RastTex &t, int texture, int index, float px, float py, float pz) {
  int u, v;
  rast_tex_uv(t, px, py, pz, 1024, index, u, v);
  return __ldg(rast_texel(t, texture == 2 ? RAST_TEX_GRILL_OPACITY : RAST_TEX_WOVEN_OPACITY, u, v, 1)) != 255;
}

// Colour, normal and occlusion the shaded (non-hole) fragment at pixel (gx, gy) uses: in/out colour[3] and
// normal[3] arrive as the triangle's own.
__device__ __forceinline__ void rast_tex_material(const RastTex &t, int texture, int index, float px, float py, float pz,
                                                  int gx, int gy, float *colour, float *normal, float &occlusion) {
  occlusion = 1.0f;
  if (texture == 1) {
    int u, v;
    rast_tex_uv(t, px, py, pz, 2000, index, u, v);
    const unsigned char *c = rast_texel(t, RAST_TEX_MARBLE, u, v, 3);
    colour[0] = xdiv((float)__ldg(c + 2), 255.0f); colour[1] = xdiv((float)__ldg(c + 1), 255.0f); colour[2] = xdiv((float)__ldg(c), 255.0f);
    long long ni = (long long)gy * t.marble_rows + gx;                       // :593
    if (ni >= t.noise_len) ni = t.noise_len - 1;                             // out of bounds in the reference: clamped
    const float4 nz = __ldg(t.noise + ni);
    normal[0] = xadd(normal[0], nz.x); normal[1] = xadd(normal[1], nz.y); normal[2] = xadd(normal[2], nz.z);
  } else if (texture == 2 || texture == 3) {
    int u, v;
    rast_tex_uv(t, px, py, pz, 1024, index, u, v);
    if (texture == 3) occlusion = xdiv((float)__ldg(rast_texel(t, RAST_TEX_WOVEN_OCCLUSION, u, v, 1)), 255.0f);   // :627-628
    const unsigned char *m = rast_texel(t, texture == 2 ? RAST_TEX_GRILL_NORMAL : RAST_TEX_WOVEN_NORMAL, u, v, 3);
    const float vx = xdiv((float)__ldg(m), 255.0f), vy = xdiv((float)__ldg(m + 1), 255.0f), vz = xdiv((float)__ldg(m + 2), 255.0f);
    // glm::normalize(vec4(x, y, z, 1)) = v * (1 / sqrt((xx + yy) + (zz + ww)))
    const float d = xadd(xadd(xmul(vx, vx), xmul(vy, vy)), xadd(xmul(vz, vz), xmul(1.0f, 1.0f)));
    const float inv = xdiv(1.0f, xsqrt(d));
    normal[0] = xmul(vx, inv); normal[1] = xmul(vy, inv); normal[2] = xmul(vz, inv);
    const unsigned char *c = rast_texel(t, texture == 2 ? RAST_TEX_GRILL : RAST_TEX_WOVEN, u, v, 3);
    colour[0] = xdiv((float)__ldg(c + 2), 255.0f); colour[1] = xdiv((float)__ldg(c + 1), 255.0f); colour[2] = xdiv((float)__ldg(c), 255.0f);
  }
}
