// Host-side scene builders behind the C ABI: the Cornell box of the reference's
// LoadTestModel (both programs) and the synthetic scenes BASELINE.md names.
// Data, not compute: this mirrors where the reference keeps it (host C++).
#include <cmath>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/b200render.h"

namespace {

struct V4 { float x, y, z, w; };

// Triangle::ComputeNormal (raytracer/Source/TestModelH.h:96-105 and
// rasteriser/Source/TestModelH.h:32-41): normalize(cross(e2, e1)), w = 1.
// volatile keeps every product/sum a separately rounded float, like the
// reference's non-FMA build.
void compute_normal(const float *v0, const float *v1, const float *v2, float *n) {
  volatile float e1x = v1[0] - v0[0], e1y = v1[1] - v0[1], e1z = v1[2] - v0[2];
  volatile float e2x = v2[0] - v0[0], e2y = v2[1] - v0[1], e2z = v2[2] - v0[2];
  volatile float a, b;
  a = e2y * e1z; b = e1y * e2z; volatile float cx = a - b;
  a = e2z * e1x; b = e1z * e2x; volatile float cy = a - b;
  a = e2x * e1y; b = e1x * e2y; volatile float cz = a - b;
  volatile float xx = cx * cx, yy = cy * cy, zz = cz * cz;
  volatile float d = xx + yy;
  d = d + zz;
  volatile float inv = 1.0f / std::sqrt((float)d);
  n[0] = cx * inv; n[1] = cy * inv; n[2] = cz * inv; n[3] = 1.0f;
}

struct Quad { V4 a, b, c; V4 d, e, f; float col[3]; int index; };  // two triangles (a,b,c), (d,e,f)

// The 555-unit Cornell box exactly as both LoadTestModel functions list it
// (raytracer/Source/TestModelH.h:143-240).  `with_tall_back` adds the tall
// block's back face, which the raytracer comments out (:231-232) and the
// rasteriser keeps (rasteriser/Source/TestModelH.h:229-237).
void cornell_raw(std::vector<Quad> &q, bool with_tall_back, bool rast_back_wall) {
  const float L = 555;
  const float red[3] = {0.75f, 0.15f, 0.15f}, yellow[3] = {0.75f, 0.75f, 0.15f}, green[3] = {0.15f, 0.75f, 0.15f},
              cyan[3] = {0.15f, 0.75f, 0.75f}, blue[3] = {0.15f, 0.15f, 0.75f}, purple[3] = {0.75f, 0.15f, 0.75f},
              white[3] = {0.75f, 0.75f, 0.75f}, rast_back[3] = {0.03529f, 0.7843f, 0.8078f};
  auto add = [&](V4 a, V4 b, V4 c, V4 d, V4 e, V4 f, const float *col, int index) {
    Quad x; x.a = a; x.b = b; x.c = c; x.d = d; x.e = e; x.f = f;
    memcpy(x.col, col, sizeof x.col); x.index = index; q.push_back(x);
  };
  V4 A{L, 0, 0, 1}, B{0, 0, 0, 1}, C{L, 0, L, 1}, D{0, 0, L, 1}, E{L, L, 0, 1}, F{0, L, 0, 1}, G{L, L, L, 1}, H{0, L, L, 1};
  add(C, B, A, C, D, B, green, 2);    // floor
  add(A, E, C, C, E, G, purple, 3);   // left wall
  add(F, B, D, H, F, D, yellow, 4);   // right wall
  add(E, F, G, F, H, G, cyan, 1);     // ceiling
  add(G, D, C, G, H, D, rast_back_wall ? rast_back : white, 0);  // back wall
  for (int block = 0; block < 2; ++block) {
    const float *col = block == 0 ? red : blue;
    if (block == 0) {
      A = {290, 0, 114, 1}; B = {130, 0, 65, 1}; C = {240, 0, 272, 1}; D = {82, 0, 225, 1};
      E = {290, 165, 114, 1}; F = {130, 165, 65, 1}; G = {240, 165, 272, 1}; H = {82, 165, 225, 1};
    } else {
      A = {423, 0, 247, 1}; B = {265, 0, 296, 1}; C = {472, 0, 406, 1}; D = {314, 0, 456, 1};
      E = {423, 330, 247, 1}; F = {265, 330, 296, 1}; G = {472, 330, 406, 1}; H = {314, 330, 456, 1};
    }
    add(E, B, A, E, F, B, col, 0);
    add(F, D, B, F, H, D, col, 4);
    if (block == 0 || with_tall_back) add(H, C, D, H, G, C, col, 0);
    add(G, E, C, E, A, C, col, 3);
    add(G, F, E, G, H, F, col, 1);
  }
}

// Scale to [-1,1]^3 and flip x, y: the loop at raytracer/Source/TestModelH.h:246-269.
void to_unit_cube(float *v) {
  const float L = 555;
  volatile float s = 2 / L;
  for (int k = 0; k < 3; ++k) {
    volatile float t = v[k] * s;
    t = t - 1.0f;
    v[k] = t;
  }
  v[3] = (v[3] * s) - 1.0f;
  v[0] *= -1; v[1] *= -1; v[3] = 1.0f;
}

void put(float *dst, const V4 &v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }

}  // namespace

extern "C" {

// raytracer LoadTestModel (raytracer/Source/TestModelH.h:121-279): 28 triangles
// + 1 sphere.  Returns the counts; copies at most the given capacities.
int b200_scene_cornell_rt(rt_triangle *tris, int tri_cap, rt_sphere *spheres, int sph_cap, int *n_tris,
                          int *n_spheres) {
  std::vector<Quad> q;
  cornell_raw(q, false, false);
  if (n_tris) *n_tris = (int)q.size() * 2;
  if (n_spheres) *n_spheres = 1;
  int k = 0;
  for (const Quad &x : q) {
    const V4 *vs[2][3] = {{&x.a, &x.b, &x.c}, {&x.d, &x.e, &x.f}};
    for (int h = 0; h < 2; ++h, ++k) {
      if (!tris || k >= tri_cap) continue;
      rt_triangle &t = tris[k];
      put(t.v0, *vs[h][0]); put(t.v1, *vs[h][1]); put(t.v2, *vs[h][2]);
      to_unit_cube(t.v0); to_unit_cube(t.v1); to_unit_cube(t.v2);
      memcpy(t.color, x.col, sizeof t.color);
      compute_normal(t.v0, t.v1, t.v2, t.normal);
    }
  }
  if (spheres && sph_cap >= 1) {
    memset(&spheres[0], 0, sizeof(rt_sphere));
    spheres[0].radius = 0.3f;                       // TestModelH.h:275-277
    spheres[0].radius_squared = 0.3f * 0.3f;
    spheres[0].centre[0] = -0.45f; spheres[0].centre[1] = 0.6f; spheres[0].centre[2] = -0.6f;
    spheres[0].color[0] = spheres[0].color[1] = spheres[0].color[2] = 0.75f;
  }
  return B200_OK;
}

// BASELINE config 5: every quad of the raytracer's Cornell box split into an
// n x n grid of cells, two triangles per cell with the parent quad's triangle
// orientation and colour (n = 60 gives 100 800 triangles), plus the sphere.
// Vertices are bilinear in the unit-cube quad corners.
int b200_scene_cornell_rt_tessellated(int n, rt_triangle *tris, int tri_cap, int *n_tris) {
  if (n < 1) return B200_EINVAL;
  std::vector<Quad> q;
  cornell_raw(q, false, false);
  if (n_tris) *n_tris = (int)q.size() * 2 * n * n;
  if (!tris) return B200_OK;
  int k = 0;
  for (const Quad &x : q) {
    // a quad is two triangles (a,b,c), (d,e,f) sharing an edge; recover its four
    // corners as a parallelogram: P(s,t) = a + s*(b - a) + t*(c - a) covers
    // triangle 1 for s+t<=1 and its mirror image is triangle 2.
    float a[4], b[4], c[4], d[4], e[4], f[4];
    put(a, x.a); put(b, x.b); put(c, x.c); put(d, x.d); put(e, x.e); put(f, x.f);
    to_unit_cube(a); to_unit_cube(b); to_unit_cube(c); to_unit_cube(d); to_unit_cube(e); to_unit_cube(f);
    const float *tri[2][3] = {{a, b, c}, {d, e, f}};
    for (int h = 0; h < 2; ++h) {
      const float *p0 = tri[h][0], *p1 = tri[h][1], *p2 = tri[h][2];
      float parent_normal[4];
      compute_normal(p0, p1, p2, parent_normal);   // children keep the parent's shading normal
      // split triangle (p0,p1,p2) into n*n similar triangles
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < 2 * (n - i) - 1; ++j) {
          if (k >= tri_cap) { ++k; continue; }
          rt_triangle &t = tris[k++];
          const int jj = j / 2;
          auto point = [&](int s, int u, float *out) {   // p0 + s/n (p1-p0) + u/n (p2-p0)
            for (int m = 0; m < 3; ++m)
              out[m] = p0[m] + ((float)s / (float)n) * (p1[m] - p0[m]) + ((float)u / (float)n) * (p2[m] - p0[m]);
            out[3] = 1.0f;
          };
          if ((j & 1) == 0) { point(jj, i, t.v0); point(jj + 1, i, t.v1); point(jj, i + 1, t.v2); }
          else { point(jj + 1, i, t.v0); point(jj + 1, i + 1, t.v1); point(jj, i + 1, t.v2); }
          memcpy(t.color, x.col, sizeof t.color);
          memcpy(t.normal, parent_normal, sizeof t.normal);
        }
    }
  }
  return B200_OK;
}

// rasteriser LoadTestModel with setting = settingBoxes = 0
// (rasteriser/Source/TestModelH.h:48-312): 10 room + 20 box triangles.
int b200_scene_cornell_rast(rast_triangle *room, int room_cap, rast_triangle *boxes, int boxes_cap,
                            int *n_room, int *n_boxes) {
  std::vector<Quad> q;
  cornell_raw(q, true, true);
  if (n_room) *n_room = 10;
  if (n_boxes) *n_boxes = (int)q.size() * 2 - 10;
  int k = 0;
  for (const Quad &x : q) {
    const V4 *vs[2][3] = {{&x.a, &x.b, &x.c}, {&x.d, &x.e, &x.f}};
    for (int h = 0; h < 2; ++h, ++k) {
      rast_triangle *t = nullptr;
      if (k < 10) { if (room && k < room_cap) t = &room[k]; }
      else if (boxes && k - 10 < boxes_cap) t = &boxes[k - 10];
      if (!t) continue;
      put(t->v0, *vs[h][0]); put(t->v1, *vs[h][1]); put(t->v2, *vs[h][2]);
      to_unit_cube(t->v0); to_unit_cube(t->v1); to_unit_cube(t->v2);
      memcpy(t->color, x.col, sizeof t->color);
      compute_normal(t->v0, t->v1, t->v2, t->normal);
      t->texture = 0;
      t->index = x.index;
    }
  }
  return B200_OK;
}

// BASELINE config 4: n-triangle random soup, std::mt19937(seed): centre uniform
// in [-1,1]^3, two edge vectors uniform in [-edge,edge]^3, colour uniform in
// [0.15,0.75]^3, texture = index = 0, normal via ComputeNormal.
int b200_scene_soup_rast(int n, uint32_t seed, float edge, rast_triangle *tris) {
  if (n < 0 || !tris) return B200_EINVAL;
  std::mt19937 gen(seed);
  auto uni = [&](float lo, float hi) {
    // 24 random bits -> [0,1): independent of the standard library's distribution code
    const float u = (float)(gen() >> 8) * (1.0f / 16777216.0f);
    return lo + (hi - lo) * u;
  };
  for (int i = 0; i < n; ++i) {
    rast_triangle &t = tris[i];
    for (int k = 0; k < 3; ++k) t.v0[k] = uni(-1.f, 1.f);
    for (int k = 0; k < 3; ++k) t.v1[k] = t.v0[k] + uni(-edge, edge);
    for (int k = 0; k < 3; ++k) t.v2[k] = t.v0[k] + uni(-edge, edge);
    t.v0[3] = t.v1[3] = t.v2[3] = 1.0f;
    for (int k = 0; k < 3; ++k) t.color[k] = uni(0.15f, 0.75f);
    t.texture = 0; t.index = 0;
    compute_normal(t.v0, t.v1, t.v2, t.normal);
  }
  return B200_OK;
}

}  // extern "C"
