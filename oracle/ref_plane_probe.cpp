// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Pins Plane::trace (reference Plane.cpp:36-73).  The reference's Scene has no addPlane, so no render reaches that
// function; this probe calls the UNMODIFIED Plane::trace directly on seeded planes and rays and dumps inputs and outputs.
// Compiled by oracle/Makefile with /root/reference/src/common/*.cpp where they lie (nothing is copied) into
// oracle/_ref/ref_plane_probe.  tests/golden/make_golden.py turns its dump into tests/golden/plane_probe.npz, against which
// tests/test_oracle.py checks the oracle's plane_trace bit for bit.
//
//   ref_plane_probe N SEED OUT     writes N records of 24 floats:
//     pos[3] norm[3] origin[3] ray[3] | hit(0/1) drop[3] outNorm[3] reflected[3] distance | shadowHit(0/1)
//   (shadowHit = the same call with every output pointer NULL, as Scene.cpp:135 makes it)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdint.h>
#include <vector>

#include "Plane.h"

extern "C" int rand(void) noexcept { return 12345; }   // trace_math.h:34 seeds a per-TU LCG from rand(); unused here, pinned anyway

namespace
{
  uint32_t g_s;
  uint32_t next() { g_s = g_s * 1664525u + 1013904223u; return g_s >> 8; }
  float unit() { return float(next()) / 16777216.0f; }            // [0, 1)
  float sym(float a) { return (unit() * 2.0f - 1.0f) * a; }       // (-a, a)
}

int main(int argc, char ** argv)
{
  if (argc != 4) { fprintf(stderr, "usage: ref_plane_probe N SEED OUT\n"); return 2; }
  const int n = atoi(argv[1]);
  g_s = (uint32_t)strtoul(argv[2], NULL, 0);
  FILE * f = fopen(argv[3], "wb");
  if (!f) return 2;
  const Material mat(Material::mtDielectric, Color(0.25f, 0.5f, 0.75f), 0.5f, 0.0f);
  for (int i = 0; i < n; i++)
  {
    Vector3 pos(sym(10.0f), sym(10.0f), sym(10.0f));
    Vector3 norm(sym(1.0f), sym(1.0f), sym(1.0f));                // not normalised on purpose: reflect() divides by n.n
    Vector3 origin(sym(20.0f), sym(20.0f), sym(20.0f));
    Vector3 ray(sym(2.0f), sym(2.0f), sym(2.0f));
    switch (i % 16)                                              // every branch of Plane.cpp:41-72 gets its share
    {
    case 1: norm = Vector3(0.0f, 1.0f, 0.0f); break;             // axis-aligned plane (the floor case)
    case 2: ray = Vector3(ray.x, 0.0f, ray.z); norm = Vector3(0.0f, 1.0f, 0.0f); break;   // parallel ray: a == 0
    case 3: origin = pos; break;                                 // starts on the plane: t == 0
    case 4: ray = ray * 1e-12f; break;                           // tiny ray: fabs(a) near 2^-63, long t
    case 5: ray = ray * 1e-20f; break;                           // fabs(a) <= 2^-63
    case 6: origin = pos + norm * 1e-6f; ray = norm * -1.0f; break;   // DELTA*DELTA boundary (dist ~1e-6 < 1e-4)
    case 7: origin = pos + norm * 1e-4f; ray = norm * -1.0f; break;   // dist ~ DELTA
    case 8: norm = Vector3(0.0f, 0.0f, 0.0f); break;             // zero normal: a == 0
    case 9: norm = norm * 1e-10f; break;                         // n.n ~1e-20 <= 2^-63: reflect() returns its input
    case 10: ray = ray * 1e9f; origin = origin * 1e3f; break;    // shadow-ray magnitudes (Scene.cpp:129)
    case 11: norm = norm * 1e6f; break;
    default: break;
    }
    const Plane plane(pos, norm, mat);
    Vector3 drop(0, 0, 0), onorm(0, 0, 0), refl(0, 0, 0);
    float dist = 0.0f;
    Material om;
    const bool hit = plane.trace(origin, ray, &drop, &onorm, &refl, &dist, &om);
    const bool shadowHit = plane.trace(origin, ray, NULL, NULL, NULL, NULL, NULL);
    const float rec[24] = { pos.x, pos.y, pos.z, norm.x, norm.y, norm.z, origin.x, origin.y, origin.z, ray.x, ray.y, ray.z,
                            hit ? 1.0f : 0.0f, drop.x, drop.y, drop.z, onorm.x, onorm.y, onorm.z, refl.x, refl.y, refl.z, dist,
                            shadowHit ? 1.0f : 0.0f };
    fwrite(rec, sizeof(float), 24, f);
  }
  fclose(f);
  return 0;
}
