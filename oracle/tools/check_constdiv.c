/* Exhaustive check (all 2^32 float bit patterns) that the three-instruction
 * sequence the CUDA post pass uses for division by the constants 5, 3 and 9,
 *     q = x * rc;  r = fma(-c, q, x);  q' = fma(r, rc, q);      rc = RN(1/c)
 * returns exactly the IEEE-754 correctly rounded quotient x / c that the reference
 * computes (rasteriser/Source/skeleton.cpp:1743-1750 `/= 5.0f`, `/3.0f`;
 * raytracer/Source/skeleton.cpp:161 `/9.0f`), for every finite x whose quotient
 * is a normal number; denormal quotients and non-finite x are listed as the
 * exceptions the kernel must route to the exact division.
 * Build: gcc -O2 -mfma -o check_constdiv check_constdiv.c -lm ; run: ./check_constdiv
 * TEST INFRASTRUCTURE ONLY. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

int main(void) {
  const float cs[3] = {5.0f, 3.0f, 9.0f};
  int fail = 0;
  for (int k = 0; k < 3; ++k) {
    const float c = cs[k];
    volatile float rcv = 1.0f / c;
    const float rc = rcv;
    uint64_t bad = 0, bad_normal = 0;
    float worst = 0.f;
    for (uint64_t b = 0; b < (1ull << 32); ++b) {
      uint32_t u = (uint32_t)b;
      float x;
      memcpy(&x, &u, 4);
      if (!isfinite(x)) continue;
      const float want = x / c;
      const float q = x * rc;
      const float r = fmaf(-c, q, x);
      const float got = fmaf(r, rc, q);
      uint32_t a, w;
      memcpy(&a, &got, 4); memcpy(&w, &want, 4);
      if (a != w) {
        ++bad;
        if (fabsf(want) >= 1.17549435e-38f * 4.0f) { ++bad_normal; if (fabsf(x) > worst) worst = fabsf(x); }
      }
    }
    printf("c = %g: %llu mismatches, %llu of them with |x/c| >= 4*FLT_MIN (largest |x| %g)\n", c,
           (unsigned long long)bad, (unsigned long long)bad_normal, worst);
    if (bad_normal) fail = 1;
  }
  return fail;
}
