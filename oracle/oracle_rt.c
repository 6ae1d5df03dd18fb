/* oracle_rt.c -- CPU restatement of the reference raytracer's hot path.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the checker; never linked into or loaded by
 * the product library.
 *
 * Parity status: PINNED.  tests/test_oracle_rt.py checks this file bit-for-bit
 * against (a) the reference's golden raytracer/screenshot.bmp (committed as
 * tests/golden/rt_screenshot_320x256.npz) and (b) the unmodified reference
 * compiled into oracle/_ref/libref_rt.so, on Cornell and random scenes.
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference repository).  The arithmetic is IEEE binary32, one rounding per
 * operation, in exactly the reference's operation order: the reference is built
 * with g++ -O3 and no -march (raytracer/Makefile:15), so there is no FMA
 * contraction; build this file with -ffp-contract=off.  The few double
 * "islands" that pow()/M_PI introduce are kept as double.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

typedef struct { float v0[4], v1[4], v2[4], normal[4], color[3]; } o_tri;   /* 76 B */
typedef struct { float radius, radius2, centre[3], color[3], normal[3]; } o_sph; /* 44 B */

typedef struct {
  float pos[4];
  float distance;
  int tri;
  int sph;
} o_isect; /* raytracer/Source/skeleton.cpp:40-45 */

/* glm::determinant(mat3) with columns a, b, c
 * (glm/glm/detail/func_matrix.inl:235-238). */
static float det3(const float *a, const float *b, const float *c) {
  float c0 = b[1] * c[2] - c[1] * b[2];
  float c1 = a[1] * c[2] - c[1] * a[2];
  float c2 = a[1] * b[2] - b[1] * a[2];
  return (a[0] * c0 - b[0] * c1) + c[0] * c2;
}

/* glm::dot(vec3) (glm/glm/detail/func_geometric.inl:65-73). */
static float dot3(const float *a, const float *b) {
  return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

/* Sphere::solveQuadratic + Sphere::intersect (raytracer/Source/TestModelH.h:24-66). */
static int sphere_intersect(const o_sph *s, const float *start, const float *dir, float *t) {
  float L[3] = {start[0] - s->centre[0], start[1] - s->centre[1], start[2] - s->centre[2]};
  float a = dot3(dir, dir);
  float b = 2.0f * dot3(dir, L);
  float c = dot3(L, L) - s->radius2;
  float disc = (b * b) - ((4.0f * a) * c);
  float x0, x1;
  if (disc < 0) return 0;
  else if (disc == 0) x0 = x1 = (float)(-0.5 * (double)b / (double)a);
  else {
    float q;
    if (b > 0) q = (float)(-0.5 * (double)(b + sqrtf(disc)));
    else q = (float)(-0.5 * (double)(b - sqrtf(disc)));
    x0 = q / a;
    x1 = c / q;
  }
  if (x0 > x1) { float tmp = x0; x0 = x1; x1 = tmp; }
  if (x0 < 0) {
    x0 = x1;
    if (x0 < 0) return 0;
  }
  *t = x0;
  return 1;
}

/* ClosestIntersection (raytracer/Source/skeleton.cpp:263-363). */
static int closest_intersection(const float *start, const float *dir, const o_tri *tris, int n_tris,
                                const o_sph *sph, int n_sph, o_isect *out) {
  const float bound = FLT_MAX;
  out->distance = bound;
  float ndir[3] = {-dir[0], -dir[1], -dir[2]};
  for (int i = 0; i < n_tris; ++i) {
    const float *v0 = tris[i].v0, *v1 = tris[i].v1, *v2 = tris[i].v2;
    float e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
    float e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
    float s[3] = {start[0] - v0[0], start[1] - v0[1], start[2] - v0[2]};
    float D = det3(ndir, e1, e2);
    float t = det3(s, e1, e2) / D;
    float len = sqrtf(dot3(dir, dir));
    float distance = t * len;
    if (distance < 0.0f) continue;
    else if (distance >= out->distance || distance > bound) continue;
    float u = det3(ndir, s, e2) / D;
    float v = det3(ndir, e1, s) / D;
    if ((u >= 0) && (v >= 0) && ((u + v) <= 1)) {
      out->pos[0] = start[0] + t * dir[0];
      out->pos[1] = start[1] + t * dir[1];
      out->pos[2] = start[2] + t * dir[2];
      out->pos[3] = start[3] + 0.0f;
      out->distance = distance;
      out->tri = i;
      out->sph = -1;
    }
  }
  for (int i = 0; i < n_sph; ++i) {
    float t;
    if (sphere_intersect(&sph[i], start, dir, &t)) {
      if (t < out->distance) {
        out->pos[0] = start[0] + t * dir[0];
        out->pos[1] = start[1] + t * dir[1];
        out->pos[2] = start[2] + t * dir[2];
        out->pos[3] = start[3] + 0.0f;
        out->distance = t;
        out->tri = -1;
        out->sph = i;
      }
    }
  }
  return out->distance < bound;
}

/* DirectLight (raytracer/Source/skeleton.cpp:366-415). light7 = pos[4], colour[3]. */
static void direct_light(const o_isect *is, const o_tri *tris, int n_tris, const o_sph *sph,
                         int n_sph, const float *light7, float *power, uint64_t *n_shadow) {
  float r[4] = {light7[0] - is->pos[0], light7[1] - is->pos[1], light7[2] - is->pos[2],
                light7[3] - is->pos[3]};
  float r_mag = (float)sqrt((double)r[0] * (double)r[0] + (double)r[1] * (double)r[1] +
                            (double)r[2] * (double)r[2]);
  float col[3], normal[4];
  if (is->tri != -1) {
    memcpy(col, tris[is->tri].color, sizeof col);
    memcpy(normal, tris[is->tri].normal, sizeof normal);
  } else {
    /* Sphere::getNormal (TestModelH.h:68-75): normalize(p - centre) */
    const o_sph *s = &sph[is->sph];
    memcpy(col, s->color, sizeof col);
    float n[3] = {is->pos[0] - s->centre[0], is->pos[1] - s->centre[1], is->pos[2] - s->centre[2]};
    float inv = 1.0f / sqrtf(dot3(n, n));
    normal[0] = n[0] * inv; normal[1] = n[1] * inv; normal[2] = n[2] * inv; normal[3] = 0.0f;
  }
  float start[4] = {is->pos[0] + normal[0] * 0.00001f, is->pos[1] + normal[1] * 0.00001f,
                    is->pos[2] + normal[2] * 0.00001f, is->pos[3] + normal[3] * 0.00001f};
  o_isect sh;
  ++*n_shadow;
  if (closest_intersection(start, r, tris, n_tris, sph, n_sph, &sh)) {
    if (sh.distance < r_mag) { power[0] = power[1] = power[2] = 0.0f; return; }
  }
  float inv = 1.0f / sqrtf(dot3(r, r));
  float nd[3] = {r[0] * inv, r[1] * inv, r[2] * inv};
  float a = dot3(nd, normal);
  float b = (float)(4 * M_PI);
  float area = (float)((double)b * ((double)r_mag * (double)r_mag));
  if (a <= 0) a = 0.f;
  for (int k = 0; k < 3; ++k) power[k] = ((col[k] * light7[4 + k]) * a) / area;
}

/* PutPixelSDL quantisation (raytracer/Source/SDLauxiliary.h:149-161). */
static uint32_t put_pixel(const float *c) {
  uint32_t ch[3];
  for (int k = 0; k < 3; ++k) {
    float v = 255 * c[k];
    v = v < 0.f ? 0.f : v;       /* glm::clamp = min(max(x, lo), hi) */
    v = v > 255.f ? 255.f : v;
    ch[k] = (uint32_t)v;
  }
  return (128u << 24) + (ch[0] << 16) + (ch[1] << 8) + ch[2];
}

void oracle_quantise(const float *rgb, size_t n, uint32_t *argb) {
  for (size_t i = 0; i < n; ++i) argb[i] = put_pixel(rgb + 3 * i);
}

/* Draw (raytracer/Source/skeleton.cpp:104-169), rows [row0, row1).
 * Outputs are full-frame arrays (only the band's rows are written); any may be
 * NULL.  dist/index are those of the centre sample (i = j = 0).
 * counts[0] += primary rays, counts[1] += shadow rays. */
int oracle_rt_render(int W, int H, float focal, const float *cam, const float *R,
                     const float *lights7, int n_lights, const void *tris_, int n_tris,
                     const void *sph_, int n_sph, int row0, int row1, float *rgb_out,
                     float *dist_out, int32_t *index_out, uint32_t *argb_out, uint64_t *counts) {
  const o_tri *tris = (const o_tri *)tris_;
  const o_sph *sph = (const o_sph *)sph_;
  uint64_t n_primary = 0, n_shadow = 0;
  for (int v = row0; v < row1; ++v) {
    for (int u = 0; u < W; ++u) {
      /* dir = R * vec4(u - W/2, v - H/2, f, 1): glm mat4*vec4 pairwise sums
       * (glm/glm/detail/type_mat4x4.inl:640-652) */
      float x = (float)(u - W / 2), y = (float)(v - H / 2);
      float dir[4];
      for (int r = 0; r < 4; ++r)
        dir[r] = (R[0 + r] * x + R[4 + r] * y) + (R[8 + r] * focal + R[12 + r] * 1.0f);
      float pix[3] = {0.f, 0.f, 0.f};
      int valid = 0;
      for (int i = -1; i <= 1; ++i) {
        for (int j = -1; j <= 1; ++j) {
          float nd[3] = {dir[0] + (0.5f * (float)i), dir[1] + (0.5f * (float)j), focal};
          o_isect is;
          ++n_primary;
          int hit = closest_intersection(cam, nd, tris, n_tris, sph, n_sph, &is);
          if (i == 0 && j == 0) {
            size_t p = (size_t)v * W + u;
            if (dist_out) dist_out[p] = hit ? is.distance : INFINITY;
            if (index_out) index_out[p] = hit ? (is.tri != -1 ? is.tri : -1 - is.sph) : INT32_MIN;
          }
          if (hit) {
            valid = 1;
            const float *oc = is.tri != -1 ? tris[is.tri].color : sph[is.sph].color;
            for (int l = 0; l < n_lights; ++l) {
              float p3[3];
              direct_light(&is, tris, n_tris, sph, n_sph, lights7 + 7 * l, p3, &n_shadow);
              pix[0] += p3[0]; pix[1] += p3[1]; pix[2] += p3[2];
            }
            pix[0] = pix[0] + oc[0] * 0.5f;
            pix[1] = pix[1] + oc[1] * 0.5f;
            pix[2] = pix[2] + oc[2] * 0.5f;
          }
        }
      }
      float outc[3] = {0.f, 0.f, 0.f};
      if (valid) { outc[0] = pix[0] / 9.0f; outc[1] = pix[1] / 9.0f; outc[2] = pix[2] / 9.0f; }
      size_t p = (size_t)v * W + u;
      if (rgb_out) { rgb_out[3 * p] = outc[0]; rgb_out[3 * p + 1] = outc[1]; rgb_out[3 * p + 2] = outc[2]; }
      if (argb_out) argb_out[p] = put_pixel(outc);
    }
  }
  if (counts) { counts[0] += n_primary; counts[1] += n_shadow; }
  return 0;
}
