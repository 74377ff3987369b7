// rfx_trace_blob.cu — K2 for scenes that do not fit the constant bank (config 4: 1024 spheres): row-aligned slices, ARGB and/or
// float image out; one sample per pixel (the batch path the bench times) or grid SSAA / additive jitter (Render API, MULTI).
// Block preview, ragged slices and signature runs of such scenes stay on k_trace (rfx_kernels.cu), whose results this kernel
// reproduces bit for bit (tests/test_gpu_parity.py::test_blob_batch_kernel_*).
//
// It is the structure of the constant-bank kernel (rfx_trace_small.cu) applied to the scene blob in global memory:
//   * Scene::trace as a per-lane state machine with ONE traversal site: the query in flight is the bounce segment (closest
//     hit) or the shadow ray of light li (any hit).  k_trace inlines the traversal twice and weighs 111 KB of SASS — a fifth
//     of its stall cycles wait for instructions; this kernel is 2 600 instructions.
//   * The BVH over the spheres (built by rfx_capi.cu, boxes inflated far beyond the float error of the exact test) only
//     selects which spheres get the reference's exact test.  It is walked over PAIR NODES (both children's boxes in the parent,
//     four 16-byte loads per trip): the walk continues into the nearer hit child in a register and defers the other one to a
//     stack in shared memory (one column per thread); "while-while" order — every lane walks to its next leaf, then the warp
//     tests its leaves together.  A leaf is four contiguous sphere records (NaN-padded) tested behind one gate, like a quad of
//     the constant-bank kernel.
//   * 4x8 pixel tiles per warp on a 2-D grid, 128-bit framebuffer stores, 64 registers / 8 CTAs per SM.
// Each step was measured (profiles/README.md: 6.66 -> 4.68 ms for a 3840x2160 frame of the 1024-sphere scene at depth 8).
//
// ARITHMETIC CONTRACT: as in rfx_trace_small.cu — --fmad=false, every + - * / sqrtf is the IEEE binary32 RN operation in the
// reference's evaluation order; the expressions below are the ones of rfx_trace_small.cu / rfx_kernels.cu.
// Reference map (path:line under /root/reference/src/common): Sphere.cpp:44-85, Triangle.cpp:53-108, Plane.cpp:36-73,
// Scene.cpp:73-236, Render.cpp:136-215.
#include "rfx_kernels.h"
#include "rfx_device.cuh"

namespace rfx
{

namespace
{

#ifndef RFX_BLOB_WW
#define RFX_BLOB_WW 1        // 1: while-while walk (every lane reaches its next leaf, then the warp tests leaves together), 0: if-if
#endif
#ifndef RFX_BLOB_PAIRS
#define RFX_BLOB_PAIRS 1     // 1: walk the pair nodes (both children's boxes per trip), 0: the single-box nodes k_trace walks
#endif
#ifndef RFX_BLOB_NEAR
#define RFX_BLOB_NEAR 1      // pair nodes: nearer child first
#endif
#ifndef RFX_BLOB_MINBLOCKS
#define RFX_BLOB_MINBLOCKS 8
#endif

#ifndef RFX_BLOB_TILE_W
#define RFX_BLOB_TILE_W 4     // pixel tile of a warp: 4 x 8 (a multiple of 4: one 128-bit framebuffer store per 4 lanes)
#endif

constexpr int BLOB_THREADS = 128;
constexpr uint32_t BLOB_TILE_W = RFX_BLOB_TILE_W, BLOB_TILE_H = 32 / RFX_BLOB_TILE_W;
constexpr int BLOB_STACK = 24;          // entries per thread; the host only routes BVHs of depth <= BLOB_STACK - 2 here

struct BlobView
{
  const SceneHeader * h;
  const Light * lights;
  const float4 * spheres;
  const Triangle * tris;
  const Plane * planes;
  const Material * mats;
  const TexRef * tex;
};

__device__ __forceinline__ BlobView blobView(const unsigned char * base)
{
  BlobView v;
  v.h = reinterpret_cast<const SceneHeader *>(base);
  v.lights = reinterpret_cast<const Light *>(base + v.h->offLights);
  v.spheres = reinterpret_cast<const float4 *>(base + v.h->offSpheres);
  v.tris = reinterpret_cast<const Triangle *>(base + v.h->offTris);
  v.planes = reinterpret_cast<const Plane *>(base + v.h->offPlanes);
  v.mats = reinterpret_cast<const Material *>(base + v.h->offMats);
  v.tex = reinterpret_cast<const TexRef *>(base + v.h->offTex);
  return v;
}

struct Hit
{
  float dist;
  int idx;           // position in the sorted object arrays (spheres, triangles, planes), -1 = none
  int order;         // insertion index (closest-hit tie-break, reference Scene.cpp:98)
  float t;
  float u, v;
};

__device__ __forceinline__ void consider(Hit & best, float dist, int idx, int order, float t, float u, float v)
{
  if (dist < best.dist || (dist == best.dist && order < best.order))
  {
    best.dist = dist; best.idx = idx; best.order = order; best.t = t; best.u = u; best.v = v;
  }
}

// All objects against one ray; `skip` = object the query ignores (-1 none); anyHit = shadow query (stop at the first hit).
__device__ __forceinline__ void intersectBlob(const BlobView & sc, int * __restrict__ stack, V3 o, V3 d, int skip, bool anyHit, Hit & best)
{
  const SceneHeader & h = *sc.h;
  const float a = vsqlen(d);                                          // Sphere.cpp:50
  const float r2x = d.x * 2.0f, r2y = d.y * 2.0f, r2z = d.z * 2.0f;   // 2.0f * ray, Sphere.cpp:51
  const float a2 = 2.0f * a;                                          // Sphere.cpp:57
  // a lane that must not take any (more) sphere hit carries NaN instead of 4a (see rfx_trace_small.cu)
  float a4 = (a > RFX_VSN) ? 4.0f * a : __int_as_float(0x7FC00000);  // Sphere.cpp:53,55
  bool open = true;                                                   // false: an any-hit query that has its occluder

  // exact sphere test (Sphere.cpp:44-85) in two halves: REJECT computes b and the discriminant, TAIL (behind the gate
  // disc >= 0 && b < 0; b < 0 is implied by t > 2^-63) the root, the distance and the bookkeeping
#define RFX_BLOB_REJECT(S, B, DISC)                                                          \
  float B, DISC;                                                                             \
  {                                                                                          \
    const float vx = o.x - S.x, vy = o.y - S.y, vz = o.z - S.z;                              \
    B = (r2x * vx + r2y * vy) + r2z * vz;                                                    \
    const float c = ((vx * vx + vy * vy) + vz * vz) - S.w;                                   \
    DISC = B * B - a4 * c;                                                                   \
  }
#define RFX_BLOB_GATE(B, DISC) ((DISC) >= 0.0f && (B) < 0.0f)
#define RFX_BLOB_TAIL(IDX, B, DISC)                                                          \
  {                                                                                          \
    const float t = (-B - sqrtf(DISC)) / a2;                          /* Sphere.cpp:57 */    \
    if (t > RFX_VSN)                                                                         \
    {                                                                                        \
      const float fx = d.x * t, fy = d.y * t, fz = d.z * t;                                  \
      const float dist = sqrtf((fx * fx + fy * fy) + fz * fz);        /* Sphere.cpp:62 */    \
      if (dist > RFX_DELTA)                                                                  \
      {                                                                                      \
        const int si = (IDX);                                                                \
        if (si != skip)                                                                      \
        {                                                                                    \
          consider(best, dist, si, sc.mats[si].order, t, 0.0f, 0.0f);                        \
          if (anyHit) { a4 = __int_as_float(0x7FC00000); open = false; }                     \
        }                                                                                    \
      }                                                                                      \
    }                                                                                        \
  }
  // one leaf = 4 contiguous sphere records (NaN in unused slots: the gate stays shut), tested behind one gate region
#define RFX_BLOB_LEAF(CA)                                                                    \
  {                                                                                          \
    const float4 * ls = h.bvhLeafSph + (CA);                                                 \
    const float4 s0 = __ldg(ls), s1 = __ldg(ls + 1), s2 = __ldg(ls + 2), s3 = __ldg(ls + 3); \
    RFX_BLOB_REJECT(s0, b0, disc0)                                                           \
    RFX_BLOB_REJECT(s1, b1, disc1)                                                           \
    RFX_BLOB_REJECT(s2, b2, disc2)                                                           \
    RFX_BLOB_REJECT(s3, b3, disc3)                                                           \
    const bool g0 = RFX_BLOB_GATE(b0, disc0), g1 = RFX_BLOB_GATE(b1, disc1), g2 = RFX_BLOB_GATE(b2, disc2), g3 = RFX_BLOB_GATE(b3, disc3); \
    if (g0 | g1 | g2 | g3)                                                                   \
    {                                                                                        \
      if (g0) RFX_BLOB_TAIL(__ldg(&h.bvhPrims[(CA)]), b0, disc0)                             \
      if (g1 && a4 == a4) RFX_BLOB_TAIL(__ldg(&h.bvhPrims[(CA) + 1]), b1, disc1)             \
      if (g2 && a4 == a4) RFX_BLOB_TAIL(__ldg(&h.bvhPrims[(CA) + 2]), b2, disc2)             \
      if (g3 && a4 == a4) RFX_BLOB_TAIL(__ldg(&h.bvhPrims[(CA) + 3]), b3, disc3)             \
    }                                                                                        \
  }

  if (h.bvhNodes == nullptr)
  {
#pragma unroll 1
    for (int i = 0; i < h.nSpheres; i++)
    {
      const float4 s = __ldg(&sc.spheres[i]);
      RFX_BLOB_REJECT(s, b, disc)
      if (RFX_BLOB_GATE(b, disc)) RFX_BLOB_TAIL(i, b, disc)
    }
  }
  else
  {
    const float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
    const float lenD = sqrtf(a);
#if RFX_BLOB_PAIRS
    // pair nodes: one trip tests both children of an inner node; the walk continues into a hit child in a register and only
    // defers the second one (the farther one) to the stack.  The bottom stack entry is the end marker.
    constexpr int DONE = (int)0x80000000;
    const float4 * __restrict__ pn = h.bvhPairs;
    int cur = h.bvhRoot;
    int sp = 1;
    stack[0] = DONE;
#pragma unroll 1
    while (cur != DONE)
    {
#if RFX_BLOB_WW
#pragma unroll 1
      while (cur >= 0)
#else
      if (cur >= 0)
#endif
      {
        const float4 n0 = __ldg(pn + 4 * cur), n1 = __ldg(pn + 4 * cur + 1), n2 = __ldg(pn + 4 * cur + 2), n3 = __ldg(pn + 4 * cur + 3);
        const float ax1 = (n0.x - o.x) * ix, ax2 = (n1.x - o.x) * ix;
        const float ay1 = (n0.y - o.y) * iy, ay2 = (n1.y - o.y) * iy;
        const float az1 = (n0.z - o.z) * iz, az2 = (n1.z - o.z) * iz;
        const float bx1 = (n2.x - o.x) * ix, bx2 = (n3.x - o.x) * ix;
        const float by1 = (n2.y - o.y) * iy, by2 = (n3.y - o.y) * iy;
        const float bz1 = (n2.z - o.z) * iz, bz2 = (n3.z - o.z) * iz;
        const float tminA = fmaxf(fmaxf(fminf(ax1, ax2), fminf(ay1, ay2)), fmaxf(fminf(az1, az2), 0.0f));
        const float tmaxA = fminf(fminf(fmaxf(ax1, ax2), fmaxf(ay1, ay2)), fmaxf(az1, az2));
        const float tminB = fmaxf(fmaxf(fminf(bx1, bx2), fminf(by1, by2)), fmaxf(fminf(bz1, bz2), 0.0f));
        const float tmaxB = fminf(fminf(fmaxf(bx1, bx2), fmaxf(by1, by2)), fmaxf(bz1, bz2));
        // a box farther than the closest hit so far cannot improve it (generous slack; +inf while there is no hit / in any-hit queries)
        const float reach = best.dist * 1.001f + 1e-2f;
        const bool hitA = tminA <= tmaxA && !(tminA * lenD > reach);
        const bool hitB = tminB <= tmaxB && !(tminB * lenD > reach);
        int ra = __float_as_int(n0.w), rb = __float_as_int(n1.w);
        if (hitA && hitB)
        {
#if RFX_BLOB_NEAR
          if (tminB < tminA) { const int t = ra; ra = rb; rb = t; }
#endif
          stack[sp * BLOB_THREADS] = rb; sp++;
          cur = ra;
        }
        else if (hitA) cur = ra;
        else if (hitB) cur = rb;
        else cur = stack[(--sp) * BLOB_THREADS];
      }
#if RFX_BLOB_WW
      if (cur == DONE) break;
#else
      else
#endif
      {
        RFX_BLOB_LEAF(~cur)
        cur = open ? stack[(--sp) * BLOB_THREADS] : DONE;
      }
    }
#else
    int sp = 1;
    stack[0] = 0;
#if RFX_BLOB_WW
    // while-while: every lane walks to its next leaf, then the warp tests its leaves together
    int leaf = -1;
#pragma unroll 1
    while (open)
    {
#pragma unroll 1
      while (sp && leaf < 0)
      {
        const int ni = stack[(--sp) * BLOB_THREADS];
        const float4 lo = __ldg(&h.bvhNodes[2 * ni]), hi = __ldg(&h.bvhNodes[2 * ni + 1]);
        const float tx1 = (lo.x - o.x) * ix, tx2 = (hi.x - o.x) * ix;
        const float ty1 = (lo.y - o.y) * iy, ty2 = (hi.y - o.y) * iy;
        const float tz1 = (lo.z - o.z) * iz, tz2 = (hi.z - o.z) * iz;
        const float tmin = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fmaxf(fminf(tz1, tz2), 0.0f));
        const float tmax = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fmaxf(tz1, tz2));
        if (!(tmin <= tmax)) continue;
        if (!anyHit && tmin * lenD > best.dist * 1.001f + 1e-2f) continue;   // cannot beat the current closest hit (generous slack)
        const int ca = __float_as_int(lo.w), cb = __float_as_int(hi.w);
        if (cb < 0) leaf = ca;
        else if (sp < BLOB_STACK - 1)
        {
          stack[sp * BLOB_THREADS] = ca; sp++;
          stack[sp * BLOB_THREADS] = cb; sp++;
        }
      }
      if (leaf < 0) break;
      RFX_BLOB_LEAF(leaf)
      leaf = -1;
    }
#else
#pragma unroll 1
    while (sp && open)
    {
      const int ni = stack[(--sp) * BLOB_THREADS];
      const float4 lo = __ldg(&h.bvhNodes[2 * ni]), hi = __ldg(&h.bvhNodes[2 * ni + 1]);
      const float tx1 = (lo.x - o.x) * ix, tx2 = (hi.x - o.x) * ix;
      const float ty1 = (lo.y - o.y) * iy, ty2 = (hi.y - o.y) * iy;
      const float tz1 = (lo.z - o.z) * iz, tz2 = (hi.z - o.z) * iz;
      const float tmin = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fmaxf(fminf(tz1, tz2), 0.0f));
      const float tmax = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fmaxf(tz1, tz2));
      if (!(tmin <= tmax)) continue;
      if (!anyHit && tmin * lenD > best.dist * 1.001f + 1e-2f) continue;   // cannot beat the current closest hit (generous slack)
      const int ca = __float_as_int(lo.w), cb = __float_as_int(hi.w);
      if (cb < 0) RFX_BLOB_LEAF(ca)
      else if (sp < BLOB_STACK - 1)
      {
        stack[sp * BLOB_THREADS] = ca; sp++;
        stack[sp * BLOB_THREADS] = cb; sp++;
      }
    }
#endif
#endif
  }
#undef RFX_BLOB_LEAF
#undef RFX_BLOB_TAIL
#undef RFX_BLOB_GATE
#undef RFX_BLOB_REJECT

  // ---- triangles (Triangle.cpp:53-108): third matrix row first; t = -oz/rz > 2^-63 needs oz, rz nonzero of opposite sign
  const float rzMin = open ? RFX_VSN : __int_as_float(0x7F800000);
  bool openT = open;
#pragma unroll 1
  for (int k = 0; k < h.nTris; k++)
  {
    const int ti = h.nSpheres + k;
    const Triangle & tr = sc.tris[k];
    const float px = o.x - tr.v0[0], py = o.y - tr.v0[1], pz = o.z - tr.v0[2];
    const float oz = (px * tr.ax[6] + py * tr.ax[7]) + pz * tr.ax[8];
    const float rz = (d.x * tr.ax[6] + d.y * tr.ax[7]) + d.z * tr.ax[8];
    if ((__float_as_int(oz) ^ __float_as_int(rz)) < 0 && fabsf(rz) > rzMin && fabsf(oz) > 0.0f && ti != skip && openT)
    {
      const float t = -oz / rz;                                       // Triangle.cpp:61
      if (t > RFX_VSN)
      {
        const float ox = (px * tr.ax[0] + py * tr.ax[1]) + pz * tr.ax[2];
        const float rx = (d.x * tr.ax[0] + d.y * tr.ax[1]) + d.z * tr.ax[2];
        const float oy = (px * tr.ax[3] + py * tr.ax[4]) + pz * tr.ax[5];
        const float ry = (d.x * tr.ax[3] + d.y * tr.ax[4]) + d.z * tr.ax[5];
        const float u = ox + t * rx;                                  // Triangle.cpp:65-66
        const float v = oy + t * ry;
        if (u >= 0.0f && v >= 0.0f && u + v < 1.0f)
        {
          const float fx = d.x * t, fy = d.y * t, fz = d.z * t;
          const float sq = (fx * fx + fy * fy) + fz * fz;
          if (sq > RFX_DELTA * RFX_DELTA)
          {
            consider(best, sqrtf(sq), ti, sc.mats[ti].order, t, u, v);
            if (anyHit) openT = false;
          }
        }
      }
    }
  }

  // ---- planes (Plane.cpp:36-73)
#pragma unroll 1
  for (int k = 0; k < h.nPlanes; k++)
  {
    const int pi = h.nSpheres + h.nTris + k;
    if (pi == skip || !openT) continue;
    const Plane & pl = sc.planes[k];
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
        if (sq > RFX_DELTA * RFX_DELTA) consider(best, sqrtf(sq), pi, sc.mats[pi].order, t, 0.0f, 0.0f);
      }
    }
  }
}

// Scene::trace (Scene.cpp:73-236) as a state machine: see traceSmall in rfx_trace_small.cu, whose expressions these are
__device__ __forceinline__ V3 traceBlob(const BlobView & sc, int * __restrict__ stack, V3 origin, V3 ray, int reflNumber, V3 randDir, uint32_t & events)
{
  const SceneHeader & h = *sc.h;
  V3 mul = mk(1.0f, 1.0f, 1.0f);
  V3 pix = mk(0.0f, 0.0f, 0.0f);
  if (reflNumber <= 0) return pix;

  V3 qo = origin, qd = ray;
  bool shadowQuery = false;
  int li = 0, hidx = -1;
  V3 norm = ray, reflect = ray, color = mul, sumLight = pix, sumSpec = pix;
  float normLen = 0.0f, reflectLen = 0.0f, mrefl = 0.0f;
  float rfs = 0.0f;          // continuation weight (Scene.cpp:196 / :207), negated for metals

  for (;;)
  {
    Hit hit;
    hit.dist = FLT_MAX; hit.idx = -1; hit.order = 0x7FFFFFFF; hit.t = 0; hit.u = 0; hit.v = 0;
    intersectBlob(sc, stack, qo, qd, shadowQuery ? hidx : -1, shadowQuery, hit);

    if (!shadowQuery)
    {
      // ---- closest hit of the bounce segment, Scene.cpp:80-112
      events++;
      if (hit.idx < 0)
      {
        float u, v;
        skyDirToUv(qd, vlen(qd), h.halfTileW, h.halfTileH, u, v);
        const V3 sky = texSampleRef(h.skyTex >= 0 ? &sc.tex[h.skyTex] : nullptr, h.byteLut, u, v);
        pix = mk(clamp01(pix.x + (mul.x * sky.x) * h.env[0]), clamp01(pix.y + (mul.y * sky.y) * h.env[1]),
                 clamp01(pix.z + (mul.z * sky.z) * h.env[2]));            // Scene.cpp:230-231
        break;
      }
      const V3 full = vscale(qd, hit.t);
      qo = vadd(qo, full);                                                 // drop point
      const Material m = sc.mats[hit.idx];
      color = mk(m.r, m.g, m.b);
      mrefl = m.reflectivity;
      const bool dielectric = m.type == 1;
      hidx = hit.idx;
      if (hit.idx < h.nSpheres)
      {
        const float4 s = __ldg(&sc.spheres[hit.idx]);
        norm = mk(qo.x - s.x, qo.y - s.y, qo.z - s.z);                     // Sphere.cpp:67
      }
      else if (hit.idx < h.nSpheres + h.nTris)
      {
        const Triangle & tr = sc.tris[hit.idx - h.nSpheres];
        norm = mk(tr.n[0], tr.n[1], tr.n[2]);
        if (m.tex >= 0)
        {
          // tuvTrans * Vector3(u, v, 0): (u*_11 + v*_12) + 0*_13 with _13 == 0, Triangle.cpp:91
          const float tx = (hit.u * tr.tuv[0] + hit.v * tr.tuv[1]) + 0.0f;
          const float ty = (hit.u * tr.tuv[2] + hit.v * tr.tuv[3]) + 0.0f;
          color = texSampleRef(&sc.tex[m.tex], h.byteLut, tr.tu0 + tx, tr.tv0 + ty);
        }
      }
      else
      {
        const Plane & pl = sc.planes[hit.idx - h.nSpheres - h.nTris];
        norm = mk(pl.n[0], pl.n[1], pl.n[2]);
      }
      reflect = reflectVec(full, norm);
      normLen = vlen(norm);
      reflectLen = vlen(reflect);
      rfs = -0.8f;                                                         // metal, Scene.cpp:207
      if (dielectric)                                                      // Scene.cpp:192-196
      {
        const float a = vlen(qd) * normLen;
        const float cosA = (a > RFX_VSN) ? clamp01(((qd.x * -norm.x + qd.y * -norm.y) + qd.z * -norm.z) / a) : 0.0f;
        rfs = 0.2f + 0.8f * cubeLikePowf(1.0f - cosA);
      }
      sumLight = mk(0.0f, 0.0f, 0.0f);
      sumSpec = mk(0.0f, 0.0f, 0.0f);
      li = 0;
    }
    else
    {
      // ---- answer of the shadow query for light li, Scene.cpp:125-186
      const Light L = sc.lights[li];
      if (hit.idx < 0)
      {
        const V3 toLight = mk(L.ox - qo.x, L.oy - qo.y, L.oz - qo.z);
        const float facing = vdot(toLight, norm);
        const float toLightLen = vlen(toLight);
        float a = toLightLen * normLen;
        const float lightDropCos = (a > RFX_VSN) ? facing / a : 0.0f;
        if (L.power > RFX_VSN)
        {
          sumLight.x = sumLight.x + (L.r * lightDropCos) * L.power;       // Scene.cpp:156
          sumLight.y = sumLight.y + (L.g * lightDropCos) * L.power;
          sumLight.z = sumLight.z + (L.b * lightDropCos) * L.power;
        }
        a = vsqlen(toLight);
        const float larsc = (a > RFX_VSN) ? 1.0f - L.radius * L.radius / a : 0.0f;   // Scene.cpp:160
        if (larsc > 0)
        {
          const V3 nl = (toLightLen > RFX_VSN) ? mk(toLight.x / toLightLen, toLight.y / toLightLen, toLight.z / toLightLen) : toLight;
          const V3 dtl = vadd(nl, vscale(randDir, 1.0f - mrefl));
          a = vlen(dtl) * reflectLen;
          float rsc = (a > RFX_VSN) ? vdot(dtl, reflect) / a : 0.0f;
          rsc = clamp01(rsc + (1.0f - sqrtf(larsc)));
          if (rsc > RFX_VSN && L.radius > RFX_VSN)
          {
            const float sp = powLikePowf(rsc, 1 + 3 * mrefl * toLightLen / L.radius) * mrefl;   // Scene.cpp:175
            sumSpec.x = sumSpec.x + L.r * sp;
            sumSpec.y = sumSpec.y + L.g * sp;
            sumSpec.z = sumSpec.z + L.b * sp;
          }
        }
      }
      li++;
    }

    // ---- next light that faces the surface gets a shadow query, Scene.cpp:118-129
    bool cast = false;
    for (; li < h.nLights; li++)
    {
      const Light L = sc.lights[li];
      const V3 toLight = mk(L.ox - qo.x, L.oy - qo.y, L.oz - qo.z);
      if (vdot(toLight, norm) > RFX_VSN)
      {
        qd = vadd(toLight, vscale(randDir, L.radius));                    // Scene.cpp:129
        cast = true;
        break;
      }
    }
    if (cast)
    {
      shadowQuery = true;
      events += 0x10000u;
      continue;
    }

    // ---- all lights answered: finish the hit, Scene.cpp:189-226
    sumLight = mk(h.ambient[0] * h.ambientPower + sumLight.x, h.ambient[1] * h.ambientPower + sumLight.y,
                  h.ambient[2] * h.ambientPower + sumLight.z);           // Scene.cpp:189
    const bool dielectric = rfs > 0.0f;
    const float rf = fabsf(rfs);
    const float k = 1.0f - rf;
    const V3 fin = mk(((color.x * k) * sumLight.x + sumSpec.x) * mul.x, ((color.y * k) * sumLight.y + sumSpec.y) * mul.y,
                      ((color.z * k) * sumLight.z + sumSpec.z) * mul.z); // Scene.cpp:198-199 / 209-210
    if (dielectric) mul = vscale(mul, rf);                               // Scene.cpp:202
    else mul = mk(mul.x * (color.x * rf), mul.y * (color.y * rf), mul.z * (color.z * rf));   // Scene.cpp:213

    pix = mk(clamp01(pix.x + fin.x), clamp01(pix.y + fin.y), clamp01(pix.z + fin.z));

    if (mul.x < 0.01f && mul.y < 0.01f && mul.z < 0.01f) break;
    if ((int)(events & 0xFFFFu) >= reflNumber) break;                    // ++refl < reflNumber, Scene.cpp:80

    const V3 rn = (reflectLen > RFX_VSN) ? mk(reflect.x / reflectLen, reflect.y / reflectLen, reflect.z / reflectLen) : reflect;
    qd = vadd(rn, vscale(randDir, 1.0f - mrefl));                        // Scene.cpp:226
    shadowQuery = false;
  }
  return pix;
}

// K2 for row-aligned slices of blob scenes.  MULTI = false: one sample per pixel, no jitter (the batch path of the bench).
// MULTI = true: the same tiling for grid SSAA (Render.cpp:174-196: the thread walks its s*s samples in the reference's ssx, ssy
// order so the sum is formed in the same order) and additive jitter / accumulation (Render.cpp:177-178, :196-207).
template <bool MULTI>
__global__ void __launch_bounds__(BLOB_THREADS, RFX_BLOB_MINBLOCKS) k_trace_blob(const unsigned char * __restrict__ sceneBlob, const __grid_constant__ FrameParams fp,
                                                                const uint32_t * __restrict__ sampleStates, uint32_t * __restrict__ argbOut,
                                                                unsigned long long * __restrict__ counters, uint32_t y0, uint32_t y1,
                                                                float * __restrict__ image)
{
  __shared__ int stackMem[BLOB_STACK * BLOB_THREADS];
  const BlobView sc = blobView(sceneBlob);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t x = (blockIdx.x * (BLOB_THREADS / 32) + warp) * BLOB_TILE_W + (lane % BLOB_TILE_W);
  const uint32_t y = y0 + blockIdx.y * BLOB_TILE_H + (lane / BLOB_TILE_W);
  const bool valid = x < fp.W && y < y1;
  uint32_t nBounces = 0, nShadow = 0, packed = 0, qOut = 0;
  if (valid)
  {
    const uint32_t q = y * fp.W + x;
    const float rx = float(x) - fp.wHalf;                                // Render.cpp:154-155
    const float ry = float(y) - fp.hHalf;
    const V3 eye = mk(fp.eye[0], fp.eye[1], fp.eye[2]);
    V3 c;
    if (!MULTI)
    {
      uint32_t s = __ldg(sampleStates + (q - y0 * fp.W));
      const V3 ray = mk((rx * fp.view[0] + ry * fp.view[1]) + fp.rz * fp.view[2],
                        (rx * fp.view[3] + ry * fp.view[4]) + fp.rz * fp.view[5],
                        (rx * fp.view[6] + ry * fp.view[7]) + fp.rz * fp.view[8]);
      V3 rd;
      rngTriple(s, rd.x, rd.y, rd.z);
      uint32_t events = 0;
      c = traceBlob(sc, stackMem + threadIdx.x, eye, ray, fp.reflNum, rd, events);   // one sample: colour / 1 == colour
      nBounces = events & 0xFFFFu; nShadow = events >> 16;
    }
    else
    {
      const int sn = fp.sampleNum;
      const uint32_t rel = q - y0 * fp.W;                                // pixel index inside the slice
      float rndx = 0, rndy = 0;
      if (fp.jitter)
      {
        uint32_t s = lcgJump(fp.seedRender, 2u * rel);                   // two draws per pixel, Render.cpp:177-178
        s = 214013u * s + 2531011u; rndx = divExact(float((int)((s >> 16) & 0x7FFFu)), 32767.0f, RFX_RCP_32767);
        s = 214013u * s + 2531011u; rndy = divExact(float((int)((s >> 16) & 0x7FFFu)), 32767.0f, RFX_RCP_32767);
      }
      c = mk(0.0f, 0.0f, 0.0f);
      const uint32_t * st = sampleStates + (size_t)rel * (size_t)(sn * sn);
#pragma unroll 1
      for (int k = 0; k < sn * sn; k++)
      {
        const int ssx = k / sn, ssy = k - ssx * sn;
        uint32_t s = __ldg(st + k);
        V3 rd;
        rngTriple(s, rd.x, rd.y, rd.z);
        const float px = (rx + float(ssx) / float(sn)) + rndx;           // Render.cpp:184
        const float py = (ry + float(ssy) / float(sn)) + rndy;
        const V3 ray = mk((px * fp.view[0] + py * fp.view[1]) + fp.rz * fp.view[2],
                          (px * fp.view[3] + py * fp.view[4]) + fp.rz * fp.view[5],
                          (px * fp.view[6] + py * fp.view[7]) + fp.rz * fp.view[8]);
        uint32_t events = 0;
        const V3 one = traceBlob(sc, stackMem + threadIdx.x, eye, ray, fp.reflNum, rd, events);
        nBounces += events & 0xFFFFu; nShadow += events >> 16;
        c = vadd(c, one);
      }
      const float sq = float(sn * sn);
      if (fabsf(sq) > RFX_VSN) c = mk(c.x / sq, c.y / sq, c.z / sq);     // Color::operator/=, Color.cpp:50-61
    }
    packed = packArgb(c.x, c.y, c.z);
    qOut = q;
    if (image)                                                           // Render.cpp:196-207
    {
      float * px = image + (size_t)q * 3;
      if (fp.accumulate) { px[0] = px[0] + c.x; px[1] = px[1] + c.y; px[2] = px[2] + c.z; }
      else { px[0] = c.x; px[1] = c.y; px[2] = c.z; }
    }
  }
  if (argbOut)
  {
    // framebuffer: one 128-bit store per tile row (see k_trace_small)
    const uint32_t p1 = __shfl_down_sync(0xffffffffu, packed, 1), p2 = __shfl_down_sync(0xffffffffu, packed, 2), p3 = __shfl_down_sync(0xffffffffu, packed, 3);
    const uint32_t validMask = __ballot_sync(0xffffffffu, valid);
    if ((fp.W & 3u) == 0u && ((validMask >> (lane & ~3u)) & 0xFu) == 0xFu)
    {
      if ((lane & 3u) == 0u) *reinterpret_cast<uint4 *>(argbOut + qOut) = make_uint4(packed, p1, p2, p3);
    }
    else if (valid) argbOut[qOut] = packed;
  }

  if (counters)
  {
    const uint32_t wb = __reduce_add_sync(0xffffffffu, nBounces);
    const uint32_t ws = __reduce_add_sync(0xffffffffu, nShadow);
    if (lane == 0)
    {
      const uint32_t slot = ((blockIdx.y * gridDim.x + blockIdx.x) * (BLOB_THREADS / 32) + warp) & 31u;
      atomicAdd(&counters[slot * 2], (unsigned long long)wb);
      atomicAdd(&counters[slot * 2 + 1], (unsigned long long)ws);
    }
  }
}

} // namespace

// 1 when the work was launched on k_trace_blob, 0 when it does not qualify (the caller falls back to k_trace)
int launchTraceBlobFast(const TraceWork & w, int bvhDepth, cudaStream_t st)
{
  const FrameParams & fp = w.fp;
  if (fp.sampleNum < 1 || fp.sampleNum > 64 || w.sigOut || (!w.argbOut && !w.image) || fp.W == 0 || fp.stripWorld) return 0;
  if (fp.p0 % fp.W != 0 || fp.p1 % fp.W != 0 || (uint64_t)fp.W * fp.H >= (1ull << 32)) return 0;
  if (bvhDepth > BLOB_STACK - 2) return 0;
  if (w.argbOut && (fp.W & 3u) == 0u && (reinterpret_cast<uintptr_t>(w.argbOut) & 15u) != 0u) return 0;   // 128-bit stores need a 16-byte aligned frame
  const uint64_t rows = (fp.p1 - fp.p0) / fp.W;
  if (rows == 0 || (rows + BLOB_TILE_H - 1) / BLOB_TILE_H > 65535u) return 0;
  const uint32_t tilesX = (fp.W + BLOB_TILE_W - 1) / BLOB_TILE_W, warps = BLOB_THREADS / 32;
  const dim3 grid((tilesX + warps - 1) / warps, (uint32_t)((rows + BLOB_TILE_H - 1) / BLOB_TILE_H));
  const unsigned char * blob = reinterpret_cast<const unsigned char *>(w.sceneBlob);
  const uint32_t y0 = (uint32_t)(fp.p0 / fp.W), y1 = (uint32_t)(fp.p1 / fp.W);
  if (fp.sampleNum == 1 && !fp.jitter) k_trace_blob<false><<<grid, BLOB_THREADS, 0, st>>>(blob, fp, w.sampleStates, w.argbOut, w.counters, y0, y1, w.image);
  else k_trace_blob<true><<<grid, BLOB_THREADS, 0, st>>>(blob, fp, w.sampleStates, w.argbOut, w.counters, y0, y1, w.image);
  return 1;
}

} // namespace rfx
