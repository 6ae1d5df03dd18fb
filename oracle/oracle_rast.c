/* oracle_rast.c -- CPU restatement of the reference rasteriser's hot path.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the checker; never linked into or loaded by
 * the product library.
 *
 * Parity status: PINNED.  tests/test_oracle_rast.py checks this file against
 * (a) KAT #1, the ComputePolygonRows table of the lab sheet
 * (rasteriser/Source/skeleton.cpp:183-199), and (b) bit-for-bit against the
 * unmodified reference compiled into oracle/_ref/libref_rast_<W>x<H>.so
 * (depth / screen / low / high / shadow buffers, final colour, clipped list),
 * plus the committed outputs of that library under tests/golden/.
 *
 * Citations are file:line in rasteriser/Source/skeleton.cpp unless noted.  IEEE
 * binary32, one rounding per operation (reference build: g++ -O3, no -march =>
 * no FMA; compile this file with -ffp-contract=off); pow()/M_PI double islands
 * kept as double.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float v0[4], v1[4], v2[4], normal[4], color[3];
  int32_t texture, index;
} o_rtri; /* 84 B, rasteriser/Source/TestModelH.h:13-42 */

typedef struct {
  int x, y;
  float zinv;
  float pos[4];
} o_pixel; /* :88-94 */

static int imax(int a, int b) { return a > b ? a : b; }

/* VertexShader :510-522 */
static void vertex_shader(const float *v, float focal, int W, int H, o_pixel *p) {
  float x = (focal * (v[0] / v[2])) + (float)(W / 2);
  float y = (focal * (v[1] / v[2])) + (float)(H / 2);
  p->x = (int)x;
  p->y = (int)y;
  p->zinv = 1 / v[2];
  memcpy(p->pos, v, sizeof p->pos);
}

/* Interpolate :524-551 */
static void interpolate(o_pixel a, o_pixel b, o_pixel *result, int N) {
  a.pos[0] = a.pos[0] * a.zinv;
  a.pos[1] = a.pos[1] * a.zinv;
  b.pos[0] = b.pos[0] * b.zinv;
  b.pos[1] = b.pos[1] * b.zinv;
  float den = (float)imax(N - 1, 1);
  float step_x = (float)(b.x - a.x) / den;
  float step_y = (float)(b.y - a.y) / den;
  float step_z = (b.zinv - a.zinv) / den;
  float step_px = (b.pos[0] - a.pos[0]) / den;
  float step_py = (b.pos[1] - a.pos[1]) / den;
  for (int i = 0; i < N; ++i) {
    result[i].x = (int)floorf((float)a.x + (step_x * (float)i));
    result[i].y = (int)floorf((float)a.y + (step_y * (float)i));
    result[i].zinv = (a.zinv + (step_z * (float)i));
    result[i].pos[2] = 1 / result[i].zinv;
    result[i].pos[0] = (a.pos[0] + (step_px * (float)i)) / result[i].zinv;
    result[i].pos[1] = (a.pos[1] + (step_py * (float)i)) / result[i].zinv;
    result[i].pos[3] = 1.0f;
  }
}

/* ComputePolygonRows :433-498.  Returns the number of rows; *left / *right are
 * malloc'ed arrays the caller frees. */
static int compute_polygon_rows(const o_pixel *vp, o_pixel **left_out, o_pixel **right_out) {
  int max = -INT_MAX, min = +INT_MAX;
  for (int i = 0; i < 3; ++i) {
    if (vp[i].y > max) max = vp[i].y;
    if (vp[i].y < min) min = vp[i].y;
  }
  int rows = (max - min) + 1;
  o_pixel *left = (o_pixel *)calloc((size_t)rows, sizeof(o_pixel));
  o_pixel *right = (o_pixel *)calloc((size_t)rows, sizeof(o_pixel));
  for (int j = 0; j < rows; ++j) {
    left[j].x = +INT_MAX; left[j].y = min + j;
    right[j].x = -INT_MAX; right[j].y = min + j;
  }
  for (int i = 0; i < 3; ++i) {
    o_pixel a = vp[i], b = vp[(i + 1) % 3];
    int dx = abs(a.x - b.x), dy = abs(a.y - b.y);
    int pixels = imax(dx, dy) + 1;
    o_pixel *line = (o_pixel *)malloc((size_t)pixels * sizeof(o_pixel));
    interpolate(a, b, line, pixels);
    for (int j = 0; j < pixels; ++j) {
      int r = line[j].y - min;
      if (r < 0) continue; /* the ":485 SEG FAULT FIXED HERE" guard */
      if (line[j].x <= left[r].x) { left[r].x = line[j].x; left[r].zinv = line[j].zinv; memcpy(left[r].pos, line[j].pos, 16); }
      if (line[j].x >= right[r].x) { right[r].x = line[j].x; right[r].zinv = line[j].zinv; memcpy(right[r].pos, line[j].pos, 16); }
    }
    free(line);
  }
  *left_out = left; *right_out = right;
  return rows;
}

typedef struct {
  int W, H;
  float focal;
  float light[4];
  float power[3];
  float indirect[3];
  float *depth, *screen, *low, *high;
  int32_t *shadow, *index;
  uint64_t fragments;
} o_frame;

/* calculateIllumination :674-688, without the final "+ indirect": returns D. */
static void illumination_D(const o_frame *f, const float *pos, const float *normal, float *D) {
  float r[3] = {f->light[0] - pos[0], f->light[1] - pos[1], f->light[2] - pos[2]};
  float r_mag = (float)((double)r[0] * (double)r[0] + (double)r[1] * (double)r[1] + (double)r[2] * (double)r[2]);
  float vp = (r[0] * normal[0] + r[1] * normal[1]) + r[2] * normal[2];
  float m = vp < 0.0f ? 0.0f : vp; /* glm::max(x, 0) = (x < 0) ? 0 : x */
  float den = (float)((double)4.0f * M_PI * (double)r_mag);
  for (int k = 0; k < 3; ++k) D[k] = (f->power[k] * m) / den;
}

/* ---- texture state (:63-75, built by main() :133-169) and the view globals findU / findV read ----
 * An image is what cv::Mat holds: rows x step bytes, at<T>(r, c) = *(T *)(data + r * step + c * sizeof(T)).
 * which: 0 marble, 1 metalGrill, 2 metalGrillOpacity (thresholded), 3 metalGrillNormalMap, 4 woven,
 * 5 woven_ambientOcclusion, 6 woven_opacity (thresholded), 7 wovenNormal. */
typedef struct { const uint8_t *data; int rows, cols, step; } o_image;
static o_image g_img[8];
static const float *g_noise;      /* normalMap_marble, vec4 per entry (:157-168) */
static long long g_noise_len;
static int g_tex_on;              /* 0: every triangle is drawn as texture == 0 (no images) */
static int g_colour_mode;         /* randColourSelect :81 */
static float g_cam[4], g_Rinv[16];
static int g_yaw_nonzero;

void oracle_rast_set_image(int which, const uint8_t *data, int rows, int cols, int step) {
  if (which < 0 || which > 7) return;
  g_img[which].data = data; g_img[which].rows = rows; g_img[which].cols = cols; g_img[which].step = step;
}
void oracle_rast_set_marble_noise(const float *noise4, long long n) { g_noise = noise4; g_noise_len = n; }
void oracle_rast_enable_textures(int on) { g_tex_on = on; }
/* 0, 1 (random colours) or 2 (night vision), :647-662; modes 1 / 2 call the C library's rand() exactly
 * where PixelShader does (glibc's, like the reference's own build: seed with srand) */
void oracle_rast_set_colour_mode(int mode) { g_colour_mode = mode; }

/* glm::inverse(mat4), glm 0.9.7.2 detail/type_mat4x4.inl:37-92; m and out column-major, m[4 * c + r] */
static void mat4_inverse(const float *m_, float *out) {
#define M(c, r) m_[4 * (c) + (r)]
  float Coef00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
  float Coef02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
  float Coef03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
  float Coef04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
  float Coef06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
  float Coef07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
  float Coef08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
  float Coef10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
  float Coef11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
  float Coef12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
  float Coef14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
  float Coef15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
  float Coef16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
  float Coef18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
  float Coef19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
  float Coef20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
  float Coef22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
  float Coef23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
  float Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07};
  float Fac2[4] = {Coef08, Coef08, Coef10, Coef11}, Fac3[4] = {Coef12, Coef12, Coef14, Coef15};
  float Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
  float Vec0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, Vec1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
  float Vec2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, Vec3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
  static const float SignA[4] = {+1, -1, +1, -1}, SignB[4] = {-1, +1, -1, +1};
  float Inv[16];
  for (int k = 0; k < 4; ++k) {
    Inv[0 + k] = ((Vec1[k] * Fac0[k] - Vec2[k] * Fac1[k]) + Vec3[k] * Fac2[k]) * SignA[k];
    Inv[4 + k] = ((Vec0[k] * Fac0[k] - Vec2[k] * Fac3[k]) + Vec3[k] * Fac4[k]) * SignB[k];
    Inv[8 + k] = ((Vec0[k] * Fac1[k] - Vec1[k] * Fac3[k]) + Vec3[k] * Fac5[k]) * SignA[k];
    Inv[12 + k] = ((Vec0[k] * Fac2[k] - Vec1[k] * Fac4[k]) + Vec2[k] * Fac5[k]) * SignB[k];
  }
  float Dot0[4] = {M(0, 0) * Inv[0], M(0, 1) * Inv[4], M(0, 2) * Inv[8], M(0, 3) * Inv[12]};
  float Dot1 = (Dot0[0] + Dot0[1]) + (Dot0[2] + Dot0[3]);
  float OneOverDeterminant = 1.0f / Dot1;
  for (int k = 0; k < 16; ++k) out[k] = Inv[k] * OneOverDeterminant;
#undef M
}

/* cameraPos, R and "yaw != 0" as findU / findV see them (:1761-1769) */
void oracle_rast_set_view(const float *cam4, const float *R16, int yaw_nonzero) {
  memcpy(g_cam, cam4, sizeof g_cam);
  mat4_inverse(R16, g_Rinv);
  g_yaw_nonzero = yaw_nonzero;
}
void oracle_rast_inverse(const float *R16, float *out16) { mat4_inverse(R16, out16); }

/* the common head of findU / findV :1759-1769 */
static void object_space(const float *pos, float *o) {
  if (g_yaw_nonzero) {
    /* glm mat4 * vec4: (m[0] * x + m[1] * y) + (m[2] * z + m[3] * w) */
    for (int r = 0; r < 4; ++r)
      o[r] = (g_Rinv[r] * pos[0] + g_Rinv[4 + r] * pos[1]) + (g_Rinv[8 + r] * pos[2] + g_Rinv[12 + r] * pos[3]);
    for (int r = 0; r < 4; ++r) o[r] = o[r] + g_cam[r];
  } else {
    for (int r = 0; r < 4; ++r) o[r] = pos[r] + g_cam[r];
  }
  o[3] = 1.0f;
}
/* float -> int as the reference build does it (cvttss2si: out-of-range and NaN give INT_MIN) */
static int to_int(float v) {
  if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT_MIN;
  return (int)v;
}
/* findU :1756-1791: int * float + int, truncated; C's % keeps the sign of u */
static int find_u(const float *pos, int size, int index) {
  float o[4];
  object_space(pos, o);
  int u = 0;
  if (index == 3) u = to_int((float)(-size / 2) * o[1] + (float)(size / 2));
  else if (index == 1) u = to_int((float)(-size / 2) * o[0] + (float)(size / 2));
  else if (index == 4) u = to_int((float)(-size / 2) * o[1] + (float)(size / 2));
  else if (index == 2) u = to_int((float)(-size / 2) * o[0] + (float)(size / 2));
  else if (index == 0) u = to_int((float)(-size / 2) * o[0] + (float)(size / 2));
  return u % size;
}
/* findV :1793-1825 */
static int find_v(const float *pos, int size, int index) {
  float o[4];
  object_space(pos, o);
  int v = 0;
  if (index == 3) v = to_int((float)(size / 2) * o[2] + (float)(size / 2));
  else if (index == 1) v = to_int((float)(-size / 2) * o[2] + (float)(size / 2));
  else if (index == 4) v = to_int((float)(-size / 2) * o[2] + (float)(size / 2));
  else if (index == 2) v = to_int((float)(-size / 2) * o[2] + (float)(size / 2));
  else if (index == 0) v = to_int((float)(-size / 2) * o[1] + (float)(size / 2));
  return v % size;
}
/* Mat::at.  A negative coordinate (a point outside the unit box the textures are laid over) reads out of
 * bounds in the reference -- undefined; here, and in the product, it wraps into the image. */
static const uint8_t *texel(const o_image *im, int u, int v, int size, int elem) {
  if (u < 0) u += size;
  if (v < 0) v += size;
  return im->data + (size_t)u * im->step + (size_t)v * elem;
}
/* glm::normalize(vec4(x, y, z, 1)) (:603, :629) = v * (1 / sqrt(dot(v, v))), dot = (xx + yy) + (zz + ww) */
static void normal_from_map(const uint8_t *t, float *n) {
  float v[4] = {(float)t[0] / 255.0f, (float)t[1] / 255.0f, (float)t[2] / 255.0f, 1.0f};
  float d = (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]);
  float inv = 1.0f / sqrtf(d);
  for (int k = 0; k < 3; ++k) n[k] = v[k] * inv;
}

/* PixelShader :559-672.  Colour mode 0: texture == 0 (:578-586), marble (:588-599), metal grill (:601-621),
 * woven wood (:623-645); colour modes 1 and 2 (:647-662). */
static void pixel_shader(o_frame *f, const o_pixel *p, const o_rtri *t, int tri_index) {
  int x = p->x, y = p->y;
  if (x >= 0 && x < f->W && y >= 0 && y < f->H) {
    size_t q = (size_t)y * f->W + x;
    f->fragments++;
    if (p->zinv >= f->depth[q] && t->color[0] >= 0 && g_colour_mode != 0) {
      /* :647-662: three rand() per accepted fragment; only screenBuffer is written; the texture fields are
       * not looked at and the global indirect light is not reset */
      float LO = 0.2f;
      float HI = 0.5f;
      float r0 = LO + (float)rand() / ((float)(RAND_MAX / HI - LO));
      float r1 = LO + (float)rand() / ((float)(RAND_MAX / HI - LO));
      float r2 = LO + (float)rand() / ((float)(RAND_MAX / HI - LO));
      float c[3] = {r0, r1, r2};
      if (g_colour_mode == 2) { c[0] = r0 - 0.2f; c[1] = 1.0f; c[2] = r2 - 0.2f; }
      float D[3];
      illumination_D(f, p->pos, t->normal, D);
      for (int k = 0; k < 3; ++k) f->screen[3 * q + k] = c[k] * (D[k] + f->indirect[k]);
      if (f->index) f->index[q] = tri_index;
      f->depth[q] = p->zinv;   /* :665 */
    } else if (p->zinv >= f->depth[q] && t->color[0] >= 0) {
      const int tex = g_tex_on ? t->texture : 0;
      float zinv = p->zinv;
      float colour[3] = {t->color[0], t->color[1], t->color[2]};
      float normal[3] = {t->normal[0], t->normal[1], t->normal[2]};
      float occlusion = 1.0f;
      int shade = tex >= 0 && tex <= 3;   /* any other value: no branch of :578-645 runs, only :665 */
      if (tex == 1) {
        const uint8_t *c = texel(&g_img[0], find_u(p->pos, 2000, t->index), find_v(p->pos, 2000, t->index), 2000, 3);
        colour[0] = (float)c[2] / 255.0f; colour[1] = (float)c[1] / 255.0f; colour[2] = (float)c[0] / 255.0f;
        long long ni = (long long)p->y * g_img[0].rows + p->x;   /* :593 */
        if (ni >= g_noise_len) ni = g_noise_len - 1;             /* out of bounds in the reference: clamped */
        for (int k = 0; k < 3; ++k) normal[k] = normal[k] + g_noise[4 * ni + k];
      } else if (tex == 2 || tex == 3) {
        const int u = find_u(p->pos, 1024, t->index), v = find_v(p->pos, 1024, t->index);
        const o_image *op = &g_img[tex == 2 ? 2 : 6], *nm = &g_img[tex == 2 ? 3 : 7], *base = &g_img[tex == 2 ? 1 : 4];
        if (*texel(op, u, v, 1024, 1) == 255) {
          if (tex == 3) { occlusion = (float)*texel(&g_img[5], u, v, 1024, 1); occlusion /= 255.0f; }
          normal_from_map(texel(nm, u, v, 1024, 3), normal);
          const uint8_t *c = texel(base, u, v, 1024, 3);
          colour[0] = (float)c[2] / 255.0f; colour[1] = (float)c[1] / 255.0f; colour[2] = (float)c[0] / 255.0f;
        } else {
          zinv = 0;        /* :619, :643: a hole -- the colours stay, the depth goes back to "empty" */
          shade = 0;
        }
      }
      if (shade) {
        float D[3];
        illumination_D(f, p->pos, normal, D);
        for (int k = 0; k < 3; ++k) {
          if (tex == 3) {   /* :634-638: textureColour * (illumination * occlusion) */
            f->screen[3 * q + k] = colour[k] * ((D[k] + f->indirect[k]) * occlusion);
            f->low[3 * q + k] = colour[k] * ((D[k] + 0.0f) * occlusion);
            f->high[3 * q + k] = colour[k] * ((D[k] + 0.4f) * occlusion);
          } else {
            f->screen[3 * q + k] = colour[k] * (D[k] + f->indirect[k]);   /* :580, the global as it stands */
            f->low[3 * q + k] = colour[k] * (D[k] + 0.0f);
            f->high[3 * q + k] = colour[k] * (D[k] + 0.4f);
          }
        }
        /* :585 leaves the GLOBAL indirectLightPowerPerArea at 0.2f * vec3(1): only the first
         * shaded fragment of a Draw ever sees the value the global had on entry */
        f->indirect[0] = f->indirect[1] = f->indirect[2] = 0.2f * 1.0f;
        if (f->index) f->index[q] = tri_index;
      }
      f->depth[q] = zinv;   /* :665 */
    } else if (p->zinv > f->depth[q] && t->color[0] < 0) {
      f->shadow[q] = 1;
    }
  }
}

/* DrawPolygon :420-431 = VertexShader x3, ComputePolygonRows, DrawPolygonRows :500-508 */
static void draw_polygon(o_frame *f, const o_rtri *t, int tri_index) {
  o_pixel vp[3];
  vertex_shader(t->v0, f->focal, f->W, f->H, &vp[0]);
  vertex_shader(t->v1, f->focal, f->W, f->H, &vp[1]);
  vertex_shader(t->v2, f->focal, f->W, f->H, &vp[2]);
  o_pixel *left, *right;
  int rows = compute_polygon_rows(vp, &left, &right);
  for (int r = 0; r < rows; ++r) {
    int n = right[r].x - left[r].x + 1;
    if (n <= 0) continue;
    o_pixel *line = (o_pixel *)malloc((size_t)n * sizeof(o_pixel));
    interpolate(left[r], right[r], line, n);
    for (int x = 0; x < (right[r].x - left[r].x); ++x) pixel_shader(f, &line[x], t, tri_index);
    free(line);
  }
  free(left); free(right);
}

/* surroundingShadowSum :1725-1733 (note [y+1][x-1] twice, [y+1][x+1] never) */
static float shadow_sum(const int32_t *s, int W, int y, int x) {
#define S(yy, xx) s[(size_t)(yy) * W + (xx)]
  float val = (float)(S(y, x) + S(y - 1, x) + S(y - 1, x - 1) + S(y - 1, x + 1) + S(y + 1, x - 1) +
                      S(y + 1, x) + S(y + 1, x - 1) + S(y, x - 1) + S(y, x + 1));
#undef S
  val /= 9.0f;
  return val;
}

/* antiAliasing :1736-1753 for one buffer: 5-tap cross / 5 */
static void cross5(const float *b, int W, int y, int x, float *out) {
  for (int k = 0; k < 3; ++k) {
#define B(yy, xx) b[3 * ((size_t)(yy) * W + (xx)) + k]
    float v = B(y, x) + B(y - 1, x);
    v = v + B(y + 1, x);
    v = v + B(y, x - 1);
    v = v + B(y, x + 1);
#undef B
    out[k] = v / 5.0f;
  }
}

static uint32_t put_pixel(const float *c) {
  uint32_t ch[3];
  for (int k = 0; k < 3; ++k) {
    float v = 255 * c[k];
    v = v < 0.f ? 0.f : v;
    v = v > 255.f ? 255.f : v;
    ch[k] = (uint32_t)v;
  }
  return (128u << 24) + (ch[0] << 16) + (ch[1] << 8) + ch[2];
}

/* The triangle loop and the post pass of Draw (:243-307) on an already-clipped
 * list.  All buffers are W*H (x3 for colours), caller-allocated; index_out,
 * screen_pre_out (screenBuffer before the in-place darkening), rgb_out and
 * argb_out may be NULL.  `screen` ends up darkened, like the reference's. */
int oracle_rast_draw_clipped(int W, int H, float focal, const float *light_cam4, const float *power3,
                             const float *indirect3, const void *tris_, int n, float *depth,
                             float *screen, float *low, float *high, int32_t *shadow,
                             int32_t *index_out, float *screen_pre_out, float *rgb_out,
                             uint32_t *argb_out, uint64_t *fragments_out) {
  const o_rtri *tris = (const o_rtri *)tris_;
  o_frame f;
  f.W = W; f.H = H; f.focal = focal;
  memcpy(f.light, light_cam4, sizeof f.light);
  memcpy(f.power, power3, sizeof f.power);
  memcpy(f.indirect, indirect3, sizeof f.indirect);
  f.depth = depth; f.screen = screen; f.low = low; f.high = high; f.shadow = shadow; f.index = index_out;
  f.fragments = 0;
  size_t np = (size_t)W * H;
  memset(depth, 0, np * sizeof(float));
  memset(screen, 0, np * 3 * sizeof(float));
  memset(low, 0, np * 3 * sizeof(float));
  memset(high, 0, np * 3 * sizeof(float));
  memset(shadow, 0, np * sizeof(int32_t));
  if (index_out) for (size_t i = 0; i < np; ++i) index_out[i] = -1;
  for (int i = 0; i < n; ++i) draw_polygon(&f, &tris[i], i);   /* :262-281 */
  if (screen_pre_out) memcpy(screen_pre_out, screen, np * 3 * sizeof(float));
  if (rgb_out) memset(rgb_out, 0, np * 3 * sizeof(float));
  if (argb_out) memset(argb_out, 0, np * sizeof(uint32_t));
  for (int y = 1; y < H - 1; ++y) {                            /* :283-307 */
    for (int x = 1; x < W - 1; ++x) {
      size_t q = (size_t)y * W + x;
      if (shadow[q] == 1) {
        float s = shadow_sum(shadow, W, y, x);
        float sub = s < 0.6 ? 0.05f : s < 0.7 ? 0.08f : s < 0.8 ? 0.1f : s < 0.9 ? 0.12f : 0.3f;
        for (int k = 0; k < 3; ++k) screen[3 * q + k] -= sub;
      }
      float a[3], b[3], c[3], out[3];
      cross5(screen, W, y, x, a);
      cross5(low, W, y, x, b);
      cross5(high, W, y, x, c);
      for (int k = 0; k < 3; ++k) out[k] = ((a[k] + b[k]) + c[k]) / 3.0f;
      if (rgb_out) memcpy(rgb_out + 3 * q, out, sizeof out);
      if (argb_out) argb_out[q] = put_pixel(out);
    }
  }
  if (fragments_out) *fragments_out = f.fragments;
  return 0;
}

/* Row table of one triangle (VertexShader + ComputePolygonRows), 8 words per
 * row like ref_rast_rows in oracle/refbuild/ref_rast_harness.cpp. */
int oracle_rast_rows(int W, int H, float focal, const float *verts12, int *y_min, int *n_rows,
                     float *rows_out, int rows_cap) {
  o_pixel vp[3];
  for (int i = 0; i < 3; ++i) vertex_shader(verts12 + 4 * i, focal, W, H, &vp[i]);
  o_pixel *L, *R;
  int rows = compute_polygon_rows(vp, &L, &R);
  *n_rows = rows;
  *y_min = rows ? L[0].y : 0;
  int rc = rows > rows_cap ? -1 : 0;
  for (int r = 0; r < rows && rc == 0; ++r) {
    float *o = rows_out + 8 * r;
    memcpy(o + 0, &L[r].x, 4); memcpy(o + 1, &R[r].x, 4);
    o[2] = L[r].zinv; o[3] = R[r].zinv;
    o[4] = L[r].pos[0]; o[5] = L[r].pos[1];
    o[6] = R[r].pos[0]; o[7] = R[r].pos[1];
  }
  free(L); free(R);
  return rc;
}

/* KAT #1 helper: ComputePolygonRows on bare pixel vertices. */
int oracle_rast_polygon_rows_kat(const int *xy6, int *y_min, int *n_rows, int *left_x, int *right_x, int cap) {
  o_pixel vp[3];
  memset(vp, 0, sizeof vp);
  for (int i = 0; i < 3; ++i) { vp[i].x = xy6[2 * i]; vp[i].y = xy6[2 * i + 1]; }
  o_pixel *L, *R;
  int rows = compute_polygon_rows(vp, &L, &R);
  *n_rows = rows;
  *y_min = rows ? L[0].y : 0;
  int rc = rows > cap ? -1 : 0;
  for (int r = 0; r < rows && rc == 0; ++r) { left_x[r] = L[r].x; right_x[r] = R[r].x; }
  free(L); free(R);
  return rc;
}
