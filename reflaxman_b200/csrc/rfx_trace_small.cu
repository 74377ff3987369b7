// rfx_trace_small.cu — K2 for scenes that fit the kernel-parameter constant bank (the reference's demo scene and
// anything up to 16 spheres / 8 triangles / 2 planes / 4 lights / 8 textures).
//
// Design (B200, FP32 CUDA cores; this path has no dense contraction, so no tensor cores):
//   * The whole scene travels as a __grid_constant__ kernel parameter: every sphere/triangle coefficient is a
//     constant-bank operand of the FADD/FMUL that uses it — no loads, no address arithmetic in the test loops.
//   * The reject tests (sphere discriminant, triangle plane-side) are straight-line, fully unrolled and branch-free;
//     they only set a per-lane candidate bit.  The rare candidates are then resolved one object at a time for the
//     whole warp (warp-uniform object index), so the expensive sqrt/divide path is entered once per distinct
//     candidate object per warp instead of once per loop iteration per lane.
//   * The same routine answers closest-hit and shadow (any-hit) queries.
//   * Warps own 4x8 pixel tiles of row-aligned slices (coherent primary/secondary rays; 8x4, 16x2, 32x1 measured slower).
//   * The bounce recursion is the reference's own bounded iterative loop carrying throughput (mulColor).
//
// ARITHMETIC CONTRACT: compiled with --fmad=false, no fast-math: every + - * / sqrtf is the IEEE binary32 RN
// operation, in the reference's evaluation order.  Hoisting a per-ray invariant (|ray|^2,
// 2*ray, 4a, 2a) or skipping work whose result is provably unused does not change any produced bit.
//
// Reference map (path:line under /root/reference/src/common): Sphere.cpp:44-85, Triangle.cpp:53-108, Plane.cpp:36-73,
// Scene.cpp:73-236, Render.cpp:136-215, trace_math.cpp:3-23, Texture.cpp:216-269, Skybox.cpp:39-106, Color.cpp:114-117.
#include "rfx_kernels.h"
#include "rfx_device.cuh"

namespace rfx
{

#ifndef RFX_SMALL_UNROLL
#define RFX_SMALL_UNROLL 0
#endif
#ifndef RFX_TILE_W
#define RFX_TILE_W 4u      // pixel tile of one warp: RFX_TILE_W x RFX_TILE_H = 32 (4x8 measured best, profiles/variants_d_r1.jsonl)
#endif
#define RFX_TILE_H (32u / RFX_TILE_W)
#ifndef RFX_SMALL_THREADS
#define RFX_SMALL_THREADS 256
#endif
#ifndef RFX_SMALL_MINBLOCKS
#define RFX_SMALL_MINBLOCKS 3
#endif

constexpr int SM_TRI_BIT = SMALL_MAX_SPHERES;                       // candidate-mask bit layout: spheres | triangles | planes
constexpr int SM_PLANE_BIT = SMALL_MAX_SPHERES + SMALL_MAX_TRIS;

struct Best
{
  float dist;
  int slot;          // candidate-mask bit of the winning object, -1 = none
  int order;         // insertion index (closest-hit tie-break, reference Scene.cpp:98 walks the list in order)
  float t;
  float u, v;        // triangle barycentrics
  float ax, ay, az;  // sphere centre, or triangle/plane normal — whatever the hit record needs from the object
};


// ---- reject tests -------------------------------------------------------------------------------------------------
// sphere i: discriminant of Sphere.cpp:49-53 with the per-ray invariants hoisted; sets bit i when d >= 0
#define RFX_SPHERE_REJECT(i)                                                         \
  {                                                                                  \
    const float vx = o.x - sc.sph[i].x, vy = o.y - sc.sph[i].y, vz = o.z - sc.sph[i].z; \
    const float b = (r2x * vx + r2y * vy) + r2z * vz;                                \
    const float c = ((vx * vx + vy * vy) + vz * vz) - sc.sph[i].w;                   \
    const float disc = b * b - a4 * c;                                               \
    if (disc >= 0.0f) mask |= 1u << (i);                                             \
  }

// triangle k: third row of axTrans*(origin - v0) and axTrans*ray (Triangle.cpp:56-57, Matrix33.cpp:232-234);
// t = -oz/rz > 2^-63 needs |rz| > 2^-63 and oz, rz of strictly opposite sign (the sign of an IEEE quotient is exact),
// so everything else is skipped without dividing
#define RFX_TRI_REJECT(k)                                                            \
  {                                                                                  \
    const float px = o.x - sc.tri[k].v0[0], py = o.y - sc.tri[k].v0[1], pz = o.z - sc.tri[k].v0[2]; \
    const float oz = (px * sc.tri[k].ax[6] + py * sc.tri[k].ax[7]) + pz * sc.tri[k].ax[8]; \
    const float rz = (d.x * sc.tri[k].ax[6] + d.y * sc.tri[k].ax[7]) + d.z * sc.tri[k].ax[8]; \
    if (fabsf(rz) > RFX_VSN && ((oz < 0.0f && rz > 0.0f) || (oz > 0.0f && rz < 0.0f))) mask |= 1u << (SM_TRI_BIT + (k)); \
  }

__device__ __forceinline__ void considerHit(Best & best, float dist, int slot, int order, float t, float u, float v, float ax, float ay, float az)
{
  if (dist < best.dist || (dist == best.dist && order < best.order))
  {
    best.dist = dist; best.slot = slot; best.order = order; best.t = t; best.u = u; best.v = v;
    best.ax = ax; best.ay = ay; best.az = az;
  }
}

// all objects against one ray; objects whose bit is set in skipMask are ignored (the shadow loop's `*obj != hitObject`)
__device__ __forceinline__ void intersectSmall(const SmallScene & sc, V3 o, V3 d, uint32_t skipMask, Best & best)
{
  const float a = vsqlen(d);                                          // Sphere.cpp:50
  const float r2x = d.x * 2.0f, r2y = d.y * 2.0f, r2z = d.z * 2.0f;   // 2.0f * ray, Sphere.cpp:51
  const float a4 = 4.0f * a, a2 = 2.0f * a;                           // Sphere.cpp:53,57
  uint32_t mask = 0;

#if RFX_SMALL_UNROLL == 2
  // experiment: counts fixed at compile time -> straight-line code, immediate constant-bank offsets, no indirect branch
#pragma unroll
  for (int i = 0; i < RFX_FIX_NS; i++) RFX_SPHERE_REJECT(i)
  if (!(a > RFX_VSN)) mask = 0;
#pragma unroll
  for (int k = 0; k < RFX_FIX_NT; k++) RFX_TRI_REJECT(k)
#elif RFX_SMALL_UNROLL
  switch (sc.nS)   // fall-through: straight-line code for exactly nS spheres
  {
  case 16: RFX_SPHERE_REJECT(15)
  case 15: RFX_SPHERE_REJECT(14)
  case 14: RFX_SPHERE_REJECT(13)
  case 13: RFX_SPHERE_REJECT(12)
  case 12: RFX_SPHERE_REJECT(11)
  case 11: RFX_SPHERE_REJECT(10)
  case 10: RFX_SPHERE_REJECT(9)
  case 9: RFX_SPHERE_REJECT(8)
  case 8: RFX_SPHERE_REJECT(7)
  case 7: RFX_SPHERE_REJECT(6)
  case 6: RFX_SPHERE_REJECT(5)
  case 5: RFX_SPHERE_REJECT(4)
  case 4: RFX_SPHERE_REJECT(3)
  case 3: RFX_SPHERE_REJECT(2)
  case 2: RFX_SPHERE_REJECT(1)
  case 1: RFX_SPHERE_REJECT(0)
  default: break;
  }
  if (!(a > RFX_VSN)) mask = 0;                                       // Sphere.cpp:55 `a > VERY_SMALL_NUMBER`

  switch (sc.nT)
  {
  case 8: RFX_TRI_REJECT(7)
  case 7: RFX_TRI_REJECT(6)
  case 6: RFX_TRI_REJECT(5)
  case 5: RFX_TRI_REJECT(4)
  case 4: RFX_TRI_REJECT(3)
  case 3: RFX_TRI_REJECT(2)
  case 2: RFX_TRI_REJECT(1)
  case 1: RFX_TRI_REJECT(0)
  default: break;
  }
#else
  // rolled: a ~20-instruction body that stays resident in the L0 instruction cache; the loop index is warp-uniform,
  // so sc.sph[i] is a uniform constant-bank load feeding uniform-register operands
  {
    uint32_t bit = 1u;
    const uint32_t endBit = 1u << sc.nS;
#pragma unroll 1
    for (int i = 0; bit != endBit; i++, bit += bit)
    {
      const float vx = o.x - sc.sph[i].x, vy = o.y - sc.sph[i].y, vz = o.z - sc.sph[i].z;
      const float b = (r2x * vx + r2y * vy) + r2z * vz;
      const float c = ((vx * vx + vy * vy) + vz * vz) - sc.sph[i].w;
      const float disc = b * b - a4 * c;
      if (disc >= 0.0f) mask |= bit;
    }
    if (!(a > RFX_VSN)) mask = 0;                                     // Sphere.cpp:55 `a > VERY_SMALL_NUMBER`
    bit = 1u << SM_TRI_BIT;
    const uint32_t endTri = bit << sc.nT;
#pragma unroll 1
    for (int k = 0; bit != endTri; k++, bit += bit)
    {
      const float px = o.x - sc.tri[k].v0[0], py = o.y - sc.tri[k].v0[1], pz = o.z - sc.tri[k].v0[2];
      const float oz = (px * sc.tri[k].ax[6] + py * sc.tri[k].ax[7]) + pz * sc.tri[k].ax[8];
      const float rz = (d.x * sc.tri[k].ax[6] + d.y * sc.tri[k].ax[7]) + d.z * sc.tri[k].ax[8];
      if (fabsf(rz) > RFX_VSN && ((oz < 0.0f && rz > 0.0f) || (oz > 0.0f && rz < 0.0f))) mask |= bit;
    }
  }
#endif
  for (int k = 0; k < sc.nP; k++) mask |= 1u << (SM_PLANE_BIT + k);   // planes are unreachable through the reference's Scene: no reject stage
  mask &= ~skipMask;

  // resolve candidates, one object at a time for every lane that flagged it
  uint32_t todo = __reduce_or_sync(__activemask(), mask);
  while (todo)
  {
    const int i = __ffs(todo) - 1;
    todo &= todo - 1;
    if ((mask >> i) & 1u)
    {
      if (i < SM_TRI_BIT)
      {
        const float4 s = sc.sph[i];
        const float vx = o.x - s.x, vy = o.y - s.y, vz = o.z - s.z;
        const float b = (r2x * vx + r2y * vy) + r2z * vz;
        const float c = ((vx * vx + vy * vy) + vz * vz) - s.w;
        const float disc = b * b - a4 * c;
        const float t = (-b - sqrtf(disc)) / a2;                      // Sphere.cpp:57
        if (t > RFX_VSN)
        {
          const float fx = d.x * t, fy = d.y * t, fz = d.z * t;
          const float dist = sqrtf((fx * fx + fy * fy) + fz * fz);    // fullRay.length(), Sphere.cpp:62
          if (dist > RFX_DELTA) considerHit(best, dist, i, sc.mat[i].order, t, 0.0f, 0.0f, s.x, s.y, s.z);
        }
      }
      else if (i < SM_PLANE_BIT)
      {
        const Triangle & tr = sc.tri[i - SM_TRI_BIT];
        const float px = o.x - tr.v0[0], py = o.y - tr.v0[1], pz = o.z - tr.v0[2];
        const float oz = (px * tr.ax[6] + py * tr.ax[7]) + pz * tr.ax[8];
        const float rz = (d.x * tr.ax[6] + d.y * tr.ax[7]) + d.z * tr.ax[8];
        const float t = -oz / rz;                                     // Triangle.cpp:61
        if (t > RFX_VSN)
        {
          const float ox = (px * tr.ax[0] + py * tr.ax[1]) + pz * tr.ax[2];
          const float rx = (d.x * tr.ax[0] + d.y * tr.ax[1]) + d.z * tr.ax[2];
          const float oy = (px * tr.ax[3] + py * tr.ax[4]) + pz * tr.ax[5];
          const float ry = (d.x * tr.ax[3] + d.y * tr.ax[4]) + d.z * tr.ax[5];
          const float u = ox + t * rx;                                // Triangle.cpp:65-66
          const float v = oy + t * ry;
          if (u >= 0.0f && v >= 0.0f && u + v < 1.0f)
          {
            const float fx = d.x * t, fy = d.y * t, fz = d.z * t;
            const float sq = (fx * fx + fy * fy) + fz * fz;
            if (sq > RFX_DELTA * RFX_DELTA)
              considerHit(best, sqrtf(sq), i, sc.mat[i].order, t, u, v, tr.n[0], tr.n[1], tr.n[2]);
          }
        }
      }
      else
      {
        const Plane & pl = sc.pl[i - SM_PLANE_BIT];                   // Plane.cpp:36-73
        const V3 n = mk(pl.n[0], pl.n[1], pl.n[2]);
        const V3 vop = mk(pl.pos[0] - o.x, pl.pos[1] - o.y, pl.pos[2] - o.z);
        const float den = vdot(n, d);
        if (fabsf(den) > RFX_VSN)
        {
          const float t = vdot(n, vop) / den;
          if (t > RFX_VSN)
          {
            const float fx = d.x * t, fy = d.y * t, fz = d.z * t;
            const float sq = (fx * fx + fy * fy) + fz * fz;
            if (sq > RFX_DELTA * RFX_DELTA) considerHit(best, sqrtf(sq), i, sc.mat[i].order, t, 0.0f, 0.0f, n.x, n.y, n.z);
          }
        }
      }
    }
  }
}

// ---- Scene::trace (reference Scene.cpp:73-236) ------------------------------------------------------------------------
__device__ __forceinline__ V3 traceSmall(const SmallScene & sc, V3 origin, V3 ray, int reflNumber, V3 randDir,
                                         uint32_t & nBounces, uint32_t & nShadow, uint32_t & sig)
{
  V3 mul = mk(1.0f, 1.0f, 1.0f);
  V3 pix = mk(0.0f, 0.0f, 0.0f);

  for (int refl = 0; refl < reflNumber; ++refl)
  {
    Best hit;
    hit.dist = FLT_MAX; hit.slot = -1; hit.order = 0x7FFFFFFF; hit.t = 0; hit.u = 0; hit.v = 0; hit.ax = hit.ay = hit.az = 0;
    nBounces++;
    intersectSmall(sc, origin, ray, 0u, hit);

    if (hit.slot < 0)
    {
      RFX_SIG(sig, 0xFFFF);
      float u, v;
      skyDirToUv(ray, sc.halfTileW, sc.halfTileH, u, v);
      const V3 sky = texSampleRef(sc.skyTex >= 0 ? &sc.tex[sc.skyTex] : nullptr, sc.byteLut, u, v);
      pix = mk(clamp01(pix.x + (mul.x * sky.x) * sc.env[0]), clamp01(pix.y + (mul.y * sky.y) * sc.env[1]),
               clamp01(pix.z + (mul.z * sky.z) * sc.env[2]));            // Scene.cpp:230-231
      break;
    }

    RFX_SIG(sig, hit.order + 1);
    const V3 full = vscale(ray, hit.t);
    const V3 drop = vadd(origin, full);
    const Material m = sc.mat[hit.slot];
    V3 norm, color = mk(m.r, m.g, m.b);
    if (hit.slot < SM_TRI_BIT)
      norm = mk(drop.x - hit.ax, drop.y - hit.ay, drop.z - hit.az);      // Sphere.cpp:67
    else
    {
      norm = mk(hit.ax, hit.ay, hit.az);
      if (m.tex >= 0)
      {
        const Triangle & tr = sc.tri[hit.slot - SM_TRI_BIT];
        // tuvTrans * Vector3(u, v, 0): (u*_11 + v*_12) + 0*_13 with _13 == 0, Triangle.cpp:91
        const float tx = (hit.u * tr.tuv[0] + hit.v * tr.tuv[1]) + 0.0f;
        const float ty = (hit.u * tr.tuv[2] + hit.v * tr.tuv[3]) + 0.0f;
        color = texSampleRef(&sc.tex[m.tex], sc.byteLut, tr.tu0 + tx, tr.tv0 + ty);
      }
    }
    const V3 reflect = reflectVec(full, norm);
    const float rayLen = vlen(ray);
    const float normLen = vlen(norm);
    const float reflectLen = vlen(reflect);
    V3 sumLight = mk(0.0f, 0.0f, 0.0f);
    V3 sumSpec = mk(0.0f, 0.0f, 0.0f);

    for (int li = 0; li < sc.nL; li++)
    {
      const Light L = sc.light[li];
      const V3 toLight = mk(L.ox - drop.x, L.oy - drop.y, L.oz - drop.z);
      const float facing = vdot(toLight, norm);
      if (facing > RFX_VSN)
      {
        const V3 sray = vadd(toLight, vscale(randDir, L.radius));        // Scene.cpp:129
        nShadow++;
        Best sh;
        sh.dist = FLT_MAX; sh.slot = -1; sh.order = 0x7FFFFFFF; sh.t = 0; sh.u = 0; sh.v = 0; sh.ax = sh.ay = sh.az = 0;
        intersectSmall(sc, drop, sray, 1u << hit.slot, sh);
        const bool inShadow = sh.slot >= 0;
        RFX_SIG(sig, 0x100 + 2 * li + (inShadow ? 1 : 0));

        if (!inShadow)
        {
          const float toLightLen = vlen(toLight);
          float a = toLightLen * normLen;
          const float lightDropCos = (a > RFX_VSN) ? facing / a : 0.0f;
          if (L.power > RFX_VSN)
          {
            sumLight.x = sumLight.x + (L.r * lightDropCos) * L.power;     // Scene.cpp:156
            sumLight.y = sumLight.y + (L.g * lightDropCos) * L.power;
            sumLight.z = sumLight.z + (L.b * lightDropCos) * L.power;
          }
          a = vsqlen(toLight);
          const float larsc = (a > RFX_VSN) ? 1.0f - L.radius * L.radius / a : 0.0f;   // Scene.cpp:160
          if (larsc > 0)
          {
            // dropToLight.normalized(): same length value as toLightLen (same operations), Vector3.cpp:55-64
            const V3 nl = (toLightLen > RFX_VSN) ? mk(toLight.x / toLightLen, toLight.y / toLightLen, toLight.z / toLightLen) : toLight;
            const V3 dtl = vadd(nl, vscale(randDir, 1.0f - m.reflectivity));
            a = vlen(dtl) * reflectLen;
            float rsc = (a > RFX_VSN) ? vdot(dtl, reflect) / a : 0.0f;
            rsc = clamp01(rsc + (1.0f - sqrtf(larsc)));
            if (rsc > RFX_VSN && L.radius > RFX_VSN)
            {
              const float sp = powLikePowf(rsc, 1 + 3 * m.reflectivity * toLightLen / L.radius) * m.reflectivity;   // Scene.cpp:175
              sumSpec.x = sumSpec.x + L.r * sp;
              sumSpec.y = sumSpec.y + L.g * sp;
              sumSpec.z = sumSpec.z + L.b * sp;
            }
          }
        }
      }
    }

    sumLight = mk(sc.ambient[0] * sc.ambientPower + sumLight.x, sc.ambient[1] * sc.ambientPower + sumLight.y,
                  sc.ambient[2] * sc.ambientPower + sumLight.z);         // Scene.cpp:189

    float rf = 0.8f;                                                     // metal, Scene.cpp:207
    if (m.type == 1)                                                     // dielectric, Scene.cpp:192-196
    {
      const float a = rayLen * normLen;
      const float cosA = (a > RFX_VSN) ? clamp01(((ray.x * -norm.x + ray.y * -norm.y) + ray.z * -norm.z) / a) : 0.0f;
      rf = 0.2f + 0.8f * cubeLikePowf(1.0f - cosA);
    }
    const float k = 1.0f - rf;
    const V3 fin = mk(((color.x * k) * sumLight.x + sumSpec.x) * mul.x, ((color.y * k) * sumLight.y + sumSpec.y) * mul.y,
                      ((color.z * k) * sumLight.z + sumSpec.z) * mul.z); // Scene.cpp:198-199 / 209-210
    if (m.type == 1) mul = vscale(mul, rf);                              // Scene.cpp:202
    else mul = mk(mul.x * (color.x * rf), mul.y * (color.y * rf), mul.z * (color.z * rf));   // Scene.cpp:213

    pix = mk(clamp01(pix.x + fin.x), clamp01(pix.y + fin.y), clamp01(pix.z + fin.z));

    if (mul.x < 0.01f && mul.y < 0.01f && mul.z < 0.01f) break;

    origin = drop;
    ray = vadd(normalizeVec(reflect), vscale(randDir, 1.0f - m.reflectivity));   // Scene.cpp:226
  }
  return pix;
}

// ---- K2 ----------------------------------------------------------------------------------------------------------------
constexpr int SMALL_THREADS = RFX_SMALL_THREADS;

__global__ void __launch_bounds__(SMALL_THREADS, RFX_SMALL_MINBLOCKS) k_trace_small(const __grid_constant__ SmallScene sc, const __grid_constant__ FrameParams fp,
                                                               const uint32_t * __restrict__ sampleStates, float * __restrict__ image,
                                                               uint32_t * __restrict__ argbOut, uint32_t * __restrict__ sigOut,
                                                               unsigned long long * __restrict__ counters, int tiled)
{
  uint32_t nBounces = 0, nShadow = 0;
  const V3 eye = mk(fp.eye[0], fp.eye[1], fp.eye[2]);
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;

  // ---- which pixel does this lane own, and how many Scene::trace calls does it make -------------------------------
  const bool blockMode = fp.sampleNum < 0;                 // block preview, Render.cpp:158-173
  const uint32_t blk = blockMode ? (uint32_t)(-fp.sampleNum) : 1u;
  const int sn = blockMode ? 1 : fp.sampleNum;             // grid SSAA factor, Render.cpp:174-196
  uint32_t x, y;
  bool valid;
  uint64_t firstState;                                     // index of this lane's first ranked random state
  if (blockMode)
  {
    // gid enumerates block origins in scan order starting at rank fp.firstRank
    const uint32_t bw = (fp.W + blk - 1) / blk;
    const uint64_t k = fp.firstRank + gid;
    y = (uint32_t)(k / bw) * blk; x = (uint32_t)(k % bw) * blk;
    valid = y < fp.H && ((uint64_t)y * fp.W + x) < fp.p1;
    firstState = gid;
  }
  else if (tiled)
  {
    // row-aligned slice (a whole frame or a band of rows): 8x4 pixel tile per warp
    const uint32_t tilesX = (fp.W + (RFX_TILE_W - 1u)) / RFX_TILE_W;
    const uint32_t y0 = (uint32_t)(fp.p0 / fp.W), y1 = (uint32_t)(fp.p1 / fp.W);
    const uint32_t warp = (uint32_t)(gid >> 5), lane = threadIdx.x & 31u;
    x = (warp % tilesX) * RFX_TILE_W + (lane % RFX_TILE_W);
    y = y0 + (warp / tilesX) * RFX_TILE_H + (lane / RFX_TILE_W);
    if (fp.stripWorld)
    {
      // split frame: the launch enumerates only this GPU's rows; compact row -> (own strip k, row in strip) -> frame row
      const uint32_t k = y / fp.stripRows;
      y = (k * fp.stripWorld + fp.stripRank) * fp.stripRows + (y % fp.stripRows);
    }
    valid = x < fp.W && y < y1;
    firstState = (((uint64_t)y * fp.W + x) - fp.p0) * (uint64_t)(sn * sn);
  }
  else
  {
    const uint64_t p = fp.p0 + gid;
    valid = p < fp.p1;
    y = (uint32_t)(p / fp.W); x = (uint32_t)(p % fp.W);
    firstState = gid * (uint64_t)(sn * sn);
  }

  if (valid)
  {
    const uint64_t p = (uint64_t)y * fp.W + x;
    const float rx = float(x) - fp.wHalf;
    const float ry = float(y) - fp.hHalf;
    float rndx = 0, rndy = 0;
    if (fp.jitter && !blockMode)
    {
      uint32_t s = lcgJump(fp.seedRender, (uint32_t)(2 * (p - fp.p0)));      // two draws per pixel, Render.cpp:177-178
      s = 214013u * s + 2531011u; rndx = divExact(float((int)((s >> 16) & 0x7FFFu)), 32767.0f, RFX_RCP_32767);
      s = 214013u * s + 2531011u; rndy = divExact(float((int)((s >> 16) & 0x7FFFu)), 32767.0f, RFX_RCP_32767);
    }
    V3 fin = mk(0.0f, 0.0f, 0.0f);
    uint32_t sig = 2166136261u;
    const uint32_t * st = sampleStates + firstState;
    const int nCalls = sn * sn;
    int ssx = 0, ssy = 0;
#pragma unroll 1
    for (int call = 0; call < nCalls; call++)
    {
      uint32_t s = st[call];
      V3 rd;
      rngTriple(s, rd.x, rd.y, rd.z);
      float px = rx, py = ry;
      if (!blockMode)
      {
        // (rx + float(ssx)/s) + rndx, Render.cpp:184; x / 1.0f == x exactly, so sn == 1 skips the divides
        const float offx = sn == 1 ? 0.0f : float(ssx) / float(sn);
        const float offy = sn == 1 ? 0.0f : float(ssy) / float(sn);
        px = (rx + offx) + rndx;
        py = (ry + offy) + rndy;
      }
      const V3 ray = mk((px * fp.view[0] + py * fp.view[1]) + fp.rz * fp.view[2],
                        (px * fp.view[3] + py * fp.view[4]) + fp.rz * fp.view[5],
                        (px * fp.view[6] + py * fp.view[7]) + fp.rz * fp.view[8]);
      const V3 c = traceSmall(sc, eye, ray, fp.reflNum, rd, nBounces, nShadow, sig);
      fin = blockMode ? c : vadd(fin, c);
      if (++ssy == sn) { ssy = 0; ssx++; }                  // ssx outer, ssy inner: the reference's summation order
    }
    if (sn != 1)   // finColor /= float(s*s): dividing by 1.0f is the identity (Color.cpp:50-61)
    {
      const float sq = float(sn * sn);
      fin = mk(fin.x / sq, fin.y / sq, fin.z / sq);
    }
    const uint32_t ex = min(x + blk, fp.W), ey = min(y + blk, fp.H);
    const uint32_t packed = packArgb(fin.x, fin.y, fin.z);
#pragma unroll 1
    for (uint32_t qy = y; qy < ey; qy++)
#pragma unroll 1
      for (uint32_t qx = x; qx < ex; qx++)
      {
        const uint64_t q = (uint64_t)qy * fp.W + qx;
        if (image)
        {
          float * px = image + q * 3;
          if (fp.accumulate && !blockMode) { px[0] = px[0] + fin.x; px[1] = px[1] + fin.y; px[2] = px[2] + fin.z; }
          else { px[0] = fin.x; px[1] = fin.y; px[2] = fin.z; }
        }
        if (argbOut) argbOut[q] = packed;
        if (sigOut) sigOut[q] = sig;
      }
  }

  // event counters: one striped atomic pair per warp
  __syncwarp();
  const uint32_t wb = __reduce_add_sync(0xffffffffu, nBounces);
  const uint32_t ws = __reduce_add_sync(0xffffffffu, nShadow);
  if ((threadIdx.x & 31) == 0 && counters)
  {
    const uint32_t slot = (blockIdx.x * (SMALL_THREADS / 32) + (threadIdx.x >> 5)) & 31u;
    atomicAdd(&counters[slot * 2], (unsigned long long)wb);
    atomicAdd(&counters[slot * 2 + 1], (unsigned long long)ws);
  }
}

int launchTraceSmall(const SmallScene & sc, const TraceWork & w, cudaStream_t st)
{
  const FrameParams & fp = w.fp;
  uint64_t nThreads;
  int tiled = 0;
  if (fp.sampleNum > 0)
  {
    if (fp.p0 % fp.W == 0 && fp.p1 % fp.W == 0)
    {
      tiled = 1;
      uint64_t rows = (fp.p1 - fp.p0) / fp.W;
      if (fp.stripWorld)
      {
        // rows owned by this rank: full strips plus the (possibly shorter) last strip, rounded up to whole strips
        const uint64_t nStrips = (rows + fp.stripRows - 1) / fp.stripRows;
        const uint64_t mine = nStrips > fp.stripRank ? (nStrips - fp.stripRank + fp.stripWorld - 1) / fp.stripWorld : 0;
        rows = mine * fp.stripRows;
      }
      nThreads = (uint64_t)((fp.W + RFX_TILE_W - 1) / RFX_TILE_W) * ((rows + RFX_TILE_H - 1) / RFX_TILE_H) * 32;
    }
    else
      nThreads = fp.p1 - fp.p0;
  }
  else
  {
    const uint32_t a = (uint32_t)(-fp.sampleNum);
    const uint32_t bw = (fp.W + a - 1) / a, bh = (fp.H + a - 1) / a;
    nThreads = (uint64_t)bw * bh - fp.firstRank;   // upper bound; threads past p1 exit
  }
  if (nThreads == 0) return 0;
  const uint32_t blocks = (uint32_t)((nThreads + SMALL_THREADS - 1) / SMALL_THREADS);
  k_trace_small<<<blocks, SMALL_THREADS, 0, st>>>(sc, fp, w.sampleStates, w.image, w.argbOut, w.sigOut, w.counters, tiled);
  return 1;
}

} // namespace rfx
