// rfx_kernels.cu — hand-written sm_100a kernels for ReflaxMan's per-pixel trace-and-shade path.
//
//   K1  k_rng_table / k_rng_prefix (once), k_rng_locate / k_rng_rank   the serial rejection-sampled LCG stream, ranked from a table of its cycle
//   K2  k_trace (any scene; small scenes: rfx_trace_small.cu)   primary rays + bounded bounce loop + shadow rays + shading + textures
//   K3  k_resolve                                   imagePixel() divide + 8-bit ARGB pack
//
// ARITHMETIC CONTRACT.  This file is compiled with --fmad=false and without any fast-math flag: every + - * below
// is an IEEE-754 binary32 round-to-nearest operation that ptxas may not fuse, / is div.rn, sqrtf is sqrt.rn, and
// denormals are kept.  Expressions are parenthesised in the order the reference's overloaded C++ operators evaluate
// them, because a last-bit difference in the geometry chain flips hit decisions of grazing
// rays and checker/texel boundaries into differences of tens of LSB.  The only arithmetic allowed
// to differ from the reference's glibc build is powf (two sites, colour only).
//
// Reference map (path:line under /root/reference/src/common):
//   rngAccept, k_rng_*          Vector3.cpp:176-188, trace_math.h:34-39
//   sphere tests                Sphere.cpp:44-85
//   triangle tests              Triangle.cpp:53-108 (setup Triangle.cpp:11-21,110-120 happens on the host)
//   plane tests                 Plane.cpp:36-73
//   reflectVec / normalizeVec   trace_math.cpp:3-23, Vector3.cpp:55-64,143-151
//   texSample                   Texture.cpp:216-269, Color.cpp:9-14
//   skySample                   Skybox.cpp:39-106
//   traceSample                 Scene.cpp:73-236
//   k_trace pixel loop          Render.cpp:136-215
//   k_resolve                   Render.cpp:103-114, Color.cpp:114-117
#include "rfx_kernels.h"
#include "rfx_device.cuh"
#include <float.h>
#include <math.h>

namespace rfx
{

uint32_t lcgJumpHost(uint32_t s, uint64_t n) { return lcgJump(s, (uint32_t)n); }

// =====================================================================================================================
// K1 — ranking the rejection-sampled LCG stream (reference Vector3.cpp:176-188, trace_math.h:34-39)
//
// Scene::trace call #q of a process owns the q-th ACCEPTED draw-triple of one LCG.  Whether a triple is accepted depends only
// on the LCG state it starts from, and that LCG (a = 214013 = 1 mod 4, c odd) walks ONE cycle through all 2^32 states.  So
// the accept pattern is a fixed sequence over the cycle, independent of the seed; the seed only says where on the cycle the
// stream starts.  Write pos(s) for the number of steps from state 0 to state s.  A stream that starts at position p0 visits
// the triples that start at positions p0, p0 + 3, p0 + 6, ... (mod 2^32): all in residue class p0 mod 3 until the position
// wraps, then in the class the wrapped position falls into (2^32 = 1 mod 3: classes follow each other 0 -> 2 -> 1 -> 0).
//
//   k_rng_table   (once per context, ~5 ms, 8.4 MB)  for every class r and every block of 2048 consecutive triples of the
//                 class: how many are accepted; k_rng_prefix turns the counts into exclusive prefix sums per class
//   k_rng_locate  (one CTA per pass)  pos(seed state) by a 32-step bitwise discrete logarithm (the low k bits of an LCG
//                 mod 2^32 have period 2^k), then the stream's offset inside its class from the table and one partially
//                 evaluated block.  For a skip-only pass (rfx_skip_samples, frame sharding) it also finds the triple that
//                 holds the last rank by binary search in the table: skipping any number of samples costs one small CTA
//   k_rng_rank    one CTA per block of the class that can hold wanted ranks: its first rank comes from the table, its
//                 2048 accept tests run once, the accepted states are scattered to their ranks.  Blocks do not depend on
//                 each other (no count pass, no scan between blocks), and a GPU that owns part of a frame only evaluates
//                 the blocks that hold its ranks
// =====================================================================================================================
__constant__ uint32_t c_jumpBlock[5][16][2];   // [nibble position][nibble value] -> (A, C) of (value << 4*pos) * 3 * RNG_TRIPLES_PER_BLOCK draws
__constant__ uint32_t c_jumpPow2[32][2];       // k -> (A, C) of 2^k draws (discrete logarithm)
__constant__ uint32_t c_classStart[3];         // state at position r = 0, 1, 2
__device__ uint2 g_jumpThread[RNG_THREADS];    // t -> (A, C) of t * 3 * RNG_TRIPLES_PER_THREAD draws (global: one coalesced 8-byte load per thread;
                                               // a per-lane index into the constant bank would serialise 32 ways)

int initRngTables()
{
  static uint32_t blk[5][16][2], thr[RNG_THREADS][2], p2[32][2], cls[3];
  auto affine = [](uint64_t draws, uint32_t out[2]) {
    const uint32_t c = lcgJump(0u, (uint32_t)draws);          // f^n(0) = C_n
    out[0] = lcgJump(1u, (uint32_t)draws) - c;                // f^n(1) - C_n = A_n
    out[1] = c;
  };
  for (int pos = 0; pos < 5; pos++)
    for (uint64_t v = 0; v < 16; v++) affine((v << (4 * pos)) * 3ull * RNG_TRIPLES_PER_BLOCK, blk[pos][v]);
  for (uint64_t t = 0; t < RNG_THREADS; t++) affine(t * 3ull * RNG_TRIPLES_PER_THREAD, thr[t]);
  for (int k = 0; k < 32; k++) affine(1ull << k, p2[k]);
  for (uint32_t r = 0; r < 3; r++) cls[r] = lcgJump(0u, r);
  if (cudaMemcpyToSymbol(c_jumpBlock, blk, sizeof(blk)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(g_jumpThread, thr, sizeof(thr)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(c_jumpPow2, p2, sizeof(p2)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(c_classStart, cls, sizeof(cls)) != cudaSuccess) return 1;
  return 0;
}

// triples of class r: positions r, r + 3, ... below 2^32
__device__ __forceinline__ uint32_t rngClassTriples(uint32_t r) { return r == 0 ? 1431655766u : 1431655765u; }

// LCG state at the first triple of this thread in block j of class r (block j starts at position r + 6144 j)
__device__ __forceinline__ uint32_t rngThreadStart(uint32_t r, uint32_t j)
{
  uint32_t s = c_classStart[r];
#pragma unroll
  for (int pos = 0; pos < 5; pos++, j >>= 4)
  {
    if (j == 0) break;                                         // uniform
    const uint32_t v = j & 15u;
    s = c_jumpBlock[pos][v][0] * s + c_jumpBlock[pos][v][1];
  }
  const uint2 m = __ldg(&g_jumpThread[threadIdx.x]);
  return m.x * s + m.y;
}

// accept bits of this thread's RNG_TRIPLES_PER_THREAD triples (bit k = triple k); triples past the end of the class are cleared
__device__ __forceinline__ uint32_t rngThreadMask(uint32_t sStart, uint32_t r, uint32_t j)
{
  uint32_t s = sStart, mask = 0;
#pragma unroll
  for (int k = 0; k < RNG_TRIPLES_PER_THREAD; k++)
  {
    float x, y, z;
    if (rngTriple(s, x, y, z)) mask |= 1u << k;
  }
  const uint32_t qFirst = j * RNG_TRIPLES_PER_BLOCK + threadIdx.x * RNG_TRIPLES_PER_THREAD, nr = rngClassTriples(r);
  if (qFirst + RNG_TRIPLES_PER_THREAD > nr) mask &= qFirst >= nr ? 0u : ((1u << (nr - qFirst)) - 1u);
  return mask;
}

// exclusive prefix of v over the CTA (RNG_THREADS threads); total receives the CTA sum
__device__ __forceinline__ int rngBlockScan(int v, int & total)
{
  __shared__ int warpSums[RNG_THREADS / 32];
  int incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1)
  {
    const int t = __shfl_up_sync(0xffffffffu, incl, off);
    if ((threadIdx.x & 31) >= off) incl += t;
  }
  __syncthreads();                                             // warpSums may still be read from a previous call
  if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = incl;
  __syncthreads();
  int wbase = 0, all = 0;
#pragma unroll
  for (int w = 0; w < RNG_THREADS / 32; w++)
  {
    if (w < (int)(threadIdx.x >> 5)) wbase += warpSums[w];
    all += warpSums[w];
  }
  total = all;
  return wbase + incl - v;
}

__global__ void __launch_bounds__(RNG_THREADS) k_rng_table(uint32_t * __restrict__ counts)
{
  const uint32_t r = blockIdx.y, j = blockIdx.x;
  const uint32_t mask = rngThreadMask(rngThreadStart(r, j), r, j);
  int total;
  rngBlockScan(__popc(mask), total);
  if (threadIdx.x == 0) counts[r * RNG_CLASS_BLOCKS + j] = (uint32_t)total;
}

// one CTA per class: prefix[r][j] = accepted triples of class r in blocks < j, j = 0 .. RNG_CLASS_BLOCKS (the last entry is the class total)
__global__ void __launch_bounds__(1024) k_rng_prefix(const uint32_t * __restrict__ counts, uint32_t * __restrict__ prefix)
{
  __shared__ uint32_t warpTot[32];
  __shared__ uint32_t carryS;
  const uint32_t r = blockIdx.x;
  counts += r * RNG_CLASS_BLOCKS;
  prefix += r * (RNG_CLASS_BLOCKS + 1);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carryS = 0u;
  __syncthreads();
  for (uint32_t base = 0; base < RNG_CLASS_BLOCKS; base += 1024u)
  {
    const uint32_t idx = base + threadIdx.x;
    const uint32_t v = idx < RNG_CLASS_BLOCKS ? counts[idx] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
    {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= (uint32_t)off) incl += t;
    }
    if (lane == 31u) warpTot[warp] = incl;
    __syncthreads();
    const uint32_t carry = carryS;
    uint32_t wt = warpTot[lane], winc = wt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
    {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, off);
      if (lane >= (uint32_t)off) winc += t;
    }
    const uint32_t warpBase = __shfl_sync(0xffffffffu, winc - wt, warp);
    const uint32_t tileTotal = __shfl_sync(0xffffffffu, winc, 31);
    if (idx < RNG_CLASS_BLOCKS) prefix[idx] = carry + warpBase + (incl - v);
    __syncthreads();
    if (threadIdx.x == 0) carryS = carry + tileTotal;
    __syncthreads();
  }
  if (threadIdx.x == 0) prefix[RNG_CLASS_BLOCKS] = carryS;
}

__global__ void __launch_bounds__(RNG_THREADS) k_rng_locate(const uint32_t * __restrict__ stateIn, const uint32_t * __restrict__ prefix,
                                                            unsigned long long n, RngLocate * __restrict__ loc,
                                                            uint32_t * __restrict__ stateOut /* skip-only pass: where the stream ends; else NULL */,
                                                            int * __restrict__ status)
{
  __shared__ uint32_t shPos, shBlock, shClass;
  __shared__ unsigned long long shWant;
  if (threadIdx.x == 0)
  {
    // pos(s): bit k of the position is set exactly when the state reached with the lower bits differs from s in bit k
    const uint32_t s = *stateIn;
    uint32_t pos = 0, t = 0;
    for (int k = 0; k < 32; k++)
      if (((t ^ s) >> k) & 1u) { t = c_jumpPow2[k][0] * t + c_jumpPow2[k][1]; pos |= 1u << k; }
    shPos = pos;
  }
  __syncthreads();
  const uint32_t p0 = shPos, r0 = p0 % 3u, q0 = p0 / 3u, j0 = q0 / RNG_TRIPLES_PER_BLOCK;
  const uint32_t r1 = r0 == 0 ? 2u : r0 - 1u;                  // class after the wrap
  const uint32_t * pre0 = prefix + r0 * (RNG_CLASS_BLOCKS + 1), * pre1 = prefix + r1 * (RNG_CLASS_BLOCKS + 1);

  // accepted triples of the class before the stream's first triple: whole blocks from the table, block j0 by evaluation
  const uint32_t qFirst = j0 * RNG_TRIPLES_PER_BLOCK + threadIdx.x * RNG_TRIPLES_PER_THREAD;
  uint32_t mask = rngThreadMask(rngThreadStart(r0, j0), r0, j0);
  const uint32_t before = q0 <= qFirst ? 0u : min(q0 - qFirst, (uint32_t)RNG_TRIPLES_PER_THREAD);
  int partial;
  rngBlockScan(__popc(mask & ((1u << before) - 1u)), partial);
  const unsigned long long off1 = (unsigned long long)pre0[j0] + (unsigned long long)partial;   // accepted before the stream in class r0
  const unsigned long long seg1 = (unsigned long long)pre0[RNG_CLASS_BLOCKS] - off1;            // accepted from the stream start to the wrap
  if (threadIdx.x == 0)
  {
    loc->r0 = r0; loc->j0 = j0; loc->r1 = r1; loc->nb1 = RNG_CLASS_BLOCKS - j0; loc->off1 = off1; loc->seg1 = seg1;
  }
  if (!stateOut) return;

  // ---- skip-only pass: the triple holding rank n-1 is the want-th accepted triple of its class
  if (threadIdx.x == 0)
  {
    uint32_t cls = r0;
    unsigned long long want = off1 + n;
    const uint32_t * pre = pre0;
    if (n > seg1) { cls = r1; want = n - seg1; pre = pre1; }
    if (want > pre[RNG_CLASS_BLOCKS]) { *status = 1; want = pre[RNG_CLASS_BLOCKS]; }   // more than two classes in one pass: callers chunk below that
    uint32_t lo = 0, hi = RNG_CLASS_BLOCKS - 1;                 // smallest block with pre[block + 1] >= want
    while (lo < hi)
    {
      const uint32_t mid = (lo + hi) >> 1;
      if (pre[mid + 1] >= want) hi = mid; else lo = mid + 1;
    }
    shBlock = lo; shClass = cls; shWant = want;
  }
  __syncthreads();
  const uint32_t cls = shClass, jE = shBlock;
  const uint32_t sStart = rngThreadStart(cls, jE);
  mask = rngThreadMask(sStart, cls, jE);
  int total;
  unsigned long long cum = (unsigned long long)(prefix + cls * (RNG_CLASS_BLOCKS + 1))[jE] + (unsigned long long)rngBlockScan(__popc(mask), total);
  uint32_t s = sStart;
#pragma unroll
  for (int k = 0; k < RNG_TRIPLES_PER_THREAD; k++)
  {
    s = 214013u * s + 2531011u;
    s = 214013u * s + 2531011u;
    s = 214013u * s + 2531011u;
    if (mask & (1u << k))
    {
      cum++;
      if (cum == shWant) *stateOut = s;
    }
  }
}

__global__ void __launch_bounds__(RNG_THREADS) k_rng_rank(const RngLocate * __restrict__ loc, const uint32_t * __restrict__ prefix,
                                                          uint32_t * __restrict__ stateOut, uint32_t * __restrict__ sampleStates,
                                                          unsigned long long n, unsigned long long ownPeriod, uint32_t ownWorld, uint32_t ownRank,
                                                          int * __restrict__ status)
{
  // virtual block v of the pass -> block j of class r; ranks of the stream = table prefix + local prefix - off
  const uint32_t v = blockIdx.x, nb1 = loc->nb1;
  uint32_t r, j;
  long long off;
  if (v < nb1) { r = loc->r0; j = loc->j0 + v; off = (long long)loc->off1; }
  else         { r = loc->r1; j = v - nb1;     off = -(long long)loc->seg1; }
  const bool lastBlock = v == gridDim.x - 1;
  if (j >= RNG_CLASS_BLOCKS)
  {
    if (lastBlock && threadIdx.x == 0) *status = 1;             // a third class in one pass: callers chunk below that
    return;
  }
  const uint32_t * pre = prefix + r * (RNG_CLASS_BLOCKS + 1);
  const long long first = (long long)pre[j] - off, next = (long long)pre[j + 1] - off;   // ranks [first, next) start in this block
  if (lastBlock && next < (long long)n && threadIdx.x == 0) *status = 1;                  // the provisioned blocks do not reach rank n-1
  if (first >= (long long)n || next <= 0) return;                                         // uniform per CTA
  if (ownWorld)
  {
    // split frames: a CTA whose ranks all lie in one strip of another GPU has nothing to store (the stream-end state is
    // written by whoever holds rank n-1, so that CTA is never skipped)
    const unsigned long long lo = (unsigned long long)(first < 0 ? 0 : first), hi = (unsigned long long)(next > (long long)n ? (long long)n : next) - 1ull;
    const unsigned long long s0 = lo / ownPeriod, s1 = hi / ownPeriod;
    if (s0 == s1 && (uint32_t)(s0 % ownWorld) != ownRank && hi < n - 1) return;
  }

  const uint32_t sStart = rngThreadStart(r, j);
  const uint32_t mask = rngThreadMask(sStart, r, j);
  int total;
  long long rank = first + (long long)rngBlockScan(__popc(mask), total);
  uint32_t s = sStart;
#pragma unroll
  for (int k = 0; k < RNG_TRIPLES_PER_THREAD; k++)
  {
    const uint32_t before = s;
    s = 214013u * s + 2531011u;
    s = 214013u * s + 2531011u;
    s = 214013u * s + 2531011u;
    if (mask & (1u << k))
    {
      if (rank >= 0 && rank < (long long)n)                     // negative: triples of block j0 before the stream's first
      {
        if (!ownWorld || (uint32_t)(((unsigned long long)rank / ownPeriod) % ownWorld) == ownRank) sampleStates[rank] = before;
        if (rank == (long long)n - 1) *stateOut = s;
      }
      rank++;
    }
  }
}

uint32_t rngBlocksFor(uint64_t n)
{
  // acceptance is pi/6 = 0.5236 (1.91 triples per sample); 2n + 4096 triples leaves > 7 sigma of slack for every n; two more
  // blocks for the partial blocks at the stream start and at the end of a class
  const uint64_t triples = 2 * n + 4096;
  return (uint32_t)((triples + RNG_TRIPLES_PER_BLOCK - 1) / RNG_TRIPLES_PER_BLOCK) + 2u;
}

int launchRngTable(uint32_t * counts, uint32_t * prefix, cudaStream_t st)
{
  k_rng_table<<<dim3(RNG_CLASS_BLOCKS, 3), RNG_THREADS, 0, st>>>(counts);
  k_rng_prefix<<<3, 1024, 0, st>>>(counts, prefix);
  return 2;
}

int launchRngRank(const RngWork & w, cudaStream_t st)
{
  if (!w.sampleStates)
  {
    k_rng_locate<<<1, RNG_THREADS, 0, st>>>(w.stateIn, w.prefix, (unsigned long long)w.n, w.locate, w.stateOut, w.status);
    return 1;
  }
  k_rng_locate<<<1, RNG_THREADS, 0, st>>>(w.stateIn, w.prefix, (unsigned long long)w.n, w.locate, nullptr, w.status);
  k_rng_rank<<<w.nBlocks, RNG_THREADS, 0, st>>>(w.locate, w.prefix, w.stateOut, w.sampleStates, (unsigned long long)w.n,
                                                 (unsigned long long)w.ownPeriod, w.ownWorld, w.ownRank, w.status);
  return 2;
}

// =====================================================================================================================
// scene view over the shared-memory copy of the blob
// =====================================================================================================================
struct SceneView
{
  const SceneHeader * h;
  const Light * lights;
  const float4 * spheres;
  const Triangle * tris;
  const Plane * planes;
  const Material * mats;
  const TexRef * tex;
};

__device__ __forceinline__ SceneView makeView(const unsigned char * base)
{
  SceneView v;
  v.h = reinterpret_cast<const SceneHeader *>(base);
  v.lights = reinterpret_cast<const Light *>(base + v.h->offLights);
  v.spheres = reinterpret_cast<const float4 *>(base + v.h->offSpheres);
  v.tris = reinterpret_cast<const Triangle *>(base + v.h->offTris);
  v.planes = reinterpret_cast<const Plane *>(base + v.h->offPlanes);
  v.mats = reinterpret_cast<const Material *>(base + v.h->offMats);
  v.tex = reinterpret_cast<const TexRef *>(base + v.h->offTex);
  return v;
}

// =====================================================================================================================
// textures / skybox: thin adapters from the shared-memory scene view onto the shared samplers (rfx_device.cuh)
// =====================================================================================================================
__device__ __forceinline__ V3 texSample(const SceneView & sc, int texId, float u, float v)
{
  return texSampleRef(texId >= 0 ? &sc.tex[texId] : nullptr, sc.h->byteLut, u, v);
}
__device__ __forceinline__ V3 skySample(const SceneView & sc, V3 ray)
{
  float u, v;
  skyDirToUv(ray, vlen(ray), sc.h->halfTileW, sc.h->halfTileH, u, v);
  return texSample(sc, sc.h->skyTex, u, v);
}

// =====================================================================================================================
// intersection: all objects against one ray.  ANYHIT = shadow query (reference passes NULL outputs).
// Per-ray invariants of the sphere test (a, 2*ray, 4a, 2a) are hoisted: same operations, evaluated once.
// =====================================================================================================================
struct HitRec
{
  int idx;        // position in the sorted object arrays (spheres, triangles, planes), -1 = none
  int order;      // insertion index (tie-break)
  float dist;
  float t;
  float u, v;     // triangle barycentrics (texture lookup)
};

template <bool ANYHIT>
__device__ __forceinline__ bool intersectAll(const SceneView & sc, V3 o, V3 d, int skipIdx, HitRec & best)
{
  const SceneHeader & h = *sc.h;
  const float a = vsqlen(d);                         // Sphere.cpp:50
  const float r2x = d.x * 2.0f, r2y = d.y * 2.0f, r2z = d.z * 2.0f;   // 2.0f * ray, Sphere.cpp:51
  const float a4 = 4.0f * a, a2 = 2.0f * a;          // Sphere.cpp:53,57
  const bool aOk = a > RFX_VSN;

  // exact test of sphere i (reference Sphere.cpp:44-85); returns true when an any-hit query is satisfied
  auto testSphere = [&](int i) -> bool
  {
    const float4 s = sc.spheres[i];
    const float vx = o.x - s.x, vy = o.y - s.y, vz = o.z - s.z;
    const float b = (r2x * vx + r2y * vy) + r2z * vz;
    const float c = ((vx * vx + vy * vy) + vz * vz) - s.w;
    const float disc = b * b - a4 * c;
    if (disc >= 0.0f && aOk && (!ANYHIT || i != skipIdx))
    {
      const float t = (-b - sqrtf(disc)) / a2;
      if (t > RFX_VSN)
      {
        const float fx = d.x * t, fy = d.y * t, fz = d.z * t;
        const float dist = sqrtf((fx * fx + fy * fy) + fz * fz);
        if (dist > RFX_DELTA)
        {
          if (ANYHIT) return true;
          const int order = sc.mats[i].order;
          if (dist < best.dist || (dist == best.dist && order < best.order))
          {
            best.dist = dist; best.idx = i; best.order = order; best.t = t;
          }
        }
      }
    }
    return false;
  };

  if (h.bvhNodes == nullptr)
  {
    for (int i = 0; i < h.nSpheres; i++)
      if (testSphere(i)) return true;
  }
  else
  {
    // Bounding-volume hierarchy over the spheres (SURVEY f-3).  It only decides WHICH spheres get the exact test above;
    // boxes are inflated by a margin three orders of magnitude above the float error of that test, the slab test is
    // NaN-tolerant (fminf/fmaxf drop NaNs), and ties still go to the lowest insertion index, so the result is identical
    // to the reference's list walk (checked against brute force in tests/test_gpu_parity.py::test_bvh_equals_brute_force).
    const float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
    const float lenD = sqrtf(a);
    int stack[32];
    int sp = 0;
    stack[sp++] = 0;
    while (sp)
    {
      const int ni = stack[--sp];
      const float4 lo = __ldg(&h.bvhNodes[2 * ni]), hi = __ldg(&h.bvhNodes[2 * ni + 1]);
      const float tx1 = (lo.x - o.x) * ix, tx2 = (hi.x - o.x) * ix;
      const float ty1 = (lo.y - o.y) * iy, ty2 = (hi.y - o.y) * iy;
      const float tz1 = (lo.z - o.z) * iz, tz2 = (hi.z - o.z) * iz;
      const float tmin = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fmaxf(fminf(tz1, tz2), 0.0f));
      const float tmax = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fmaxf(tz1, tz2));
      if (!(tmin <= tmax)) continue;
      if (!ANYHIT && tmin * lenD > best.dist * 1.001f + 1e-2f) continue;   // cannot beat the current closest hit (generous slack)
      const int ca = __float_as_int(lo.w), cb = __float_as_int(hi.w);
      if (cb < 0)
      {
        for (int k = 0; k < -cb; k++)
          if (testSphere(__ldg(&h.bvhPrims[ca + k]))) return true;
      }
      else if (sp < 31)
      {
        stack[sp++] = ca;
        stack[sp++] = cb;
      }
    }
  }

  for (int k = 0; k < h.nTris; k++)
  {
    const int i = h.nSpheres + k;
    if (ANYHIT && i == skipIdx) continue;
    const Triangle & tr = sc.tris[k];
    const float px = o.x - tr.v0[0], py = o.y - tr.v0[1], pz = o.z - tr.v0[2];
    // third row first (the only part the early-outs need); rows are (x*m1 + y*m2) + z*m3, Matrix33.cpp:232-234
    const float oz = (px * tr.ax[6] + py * tr.ax[7]) + pz * tr.ax[8];
    const float rz = (d.x * tr.ax[6] + d.y * tr.ax[7]) + d.z * tr.ax[8];
    if (fabsf(rz) > RFX_VSN)
    {
      const float t = -oz / rz;
      if (t > RFX_VSN)
      {
        const float ox = (px * tr.ax[0] + py * tr.ax[1]) + pz * tr.ax[2];
        const float rx = (d.x * tr.ax[0] + d.y * tr.ax[1]) + d.z * tr.ax[2];
        const float oy = (px * tr.ax[3] + py * tr.ax[4]) + pz * tr.ax[5];
        const float ry = (d.x * tr.ax[3] + d.y * tr.ax[4]) + d.z * tr.ax[5];
        const float u = ox + t * rx;
        const float v = oy + t * ry;
        if (u >= 0.0f && v >= 0.0f && u + v < 1.0f)
        {
          const float fx = d.x * t, fy = d.y * t, fz = d.z * t;
          const float sq = (fx * fx + fy * fy) + fz * fz;
          if (sq > RFX_DELTA * RFX_DELTA)
          {
            if (ANYHIT) return true;
            const float dist = sqrtf(sq);
            const int order = sc.mats[i].order;
            if (dist < best.dist || (dist == best.dist && order < best.order))
            {
              best.dist = dist; best.idx = i; best.order = order; best.t = t; best.u = u; best.v = v;
            }
          }
        }
      }
    }
  }

  for (int k = 0; k < h.nPlanes; k++)
  {
    const int i = h.nSpheres + h.nTris + k;
    if (ANYHIT && i == skipIdx) continue;
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
        if (sq > RFX_DELTA * RFX_DELTA)
        {
          if (ANYHIT) return true;
          const float dist = sqrtf(sq);
          const int order = sc.mats[i].order;
          if (dist < best.dist || (dist == best.dist && order < best.order))
          {
            best.dist = dist; best.idx = i; best.order = order; best.t = t;
          }
        }
      }
    }
  }
  return ANYHIT ? false : best.idx >= 0;
}

// =====================================================================================================================
// Scene::trace (reference Scene.cpp:73-236)
// =====================================================================================================================
__device__ V3 traceSample(const SceneView & sc, V3 origin, V3 ray, int reflNumber, V3 randDir,
                          uint32_t & nBounces, uint32_t & nShadow, uint32_t & sig)
{
  const SceneHeader & h = *sc.h;
  V3 mul = mk(1.0f, 1.0f, 1.0f);
  V3 pix = mk(0.0f, 0.0f, 0.0f);

  for (int refl = 0; refl < reflNumber; ++refl)
  {
    HitRec hit;
    hit.idx = -1; hit.order = 0x7FFFFFFF; hit.dist = FLT_MAX; hit.t = 0; hit.u = 0; hit.v = 0;
    nBounces++;

    if (intersectAll<false>(sc, origin, ray, -1, hit))
    {
      RFX_SIG(sig, hit.order + 1);
      // outputs of the winning object's trace(): drop, norm, reflect, material
      const V3 full = vscale(ray, hit.t);
      const V3 drop = vadd(origin, full);
      const Material m = sc.mats[hit.idx];
      V3 norm, color = mk(m.r, m.g, m.b);
      if (hit.idx < h.nSpheres)
      {
        const float4 s = sc.spheres[hit.idx];
        norm = mk(drop.x - s.x, drop.y - s.y, drop.z - s.z);               // Sphere.cpp:67
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
          color = texSample(sc, m.tex, tr.tu0 + tx, tr.tv0 + ty);
        }
      }
      else
      {
        const Plane & pl = sc.planes[hit.idx - h.nSpheres - h.nTris];
        norm = mk(pl.n[0], pl.n[1], pl.n[2]);
      }
      const V3 reflect = reflectVec(full, norm);

      const float rayLen = vlen(ray);
      const float normLen = vlen(norm);
      const float reflectLen = vlen(reflect);
      V3 sumLight = mk(0.0f, 0.0f, 0.0f);
      V3 sumSpec = mk(0.0f, 0.0f, 0.0f);

      for (int li = 0; li < h.nLights; li++)
      {
        const Light L = sc.lights[li];
        const V3 toLight = mk(L.ox - drop.x, L.oy - drop.y, L.oz - drop.z);
        const float facing = vdot(toLight, norm);
        if (facing > RFX_VSN)
        {
          const V3 sray = vadd(toLight, vscale(randDir, L.radius));        // Scene.cpp:129
          nShadow++;
          HitRec dummy;
          const bool inShadow = intersectAll<true>(sc, drop, sray, hit.idx, dummy);
          RFX_SIG(sig, 0x100 + 2 * li + (inShadow ? 1 : 0));

          if (!inShadow)
          {
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
              const V3 dtl = vadd(normalizeVec(toLight), vscale(randDir, 1.0f - m.reflectivity));
              a = vlen(dtl) * reflectLen;
              float rsc = (a > RFX_VSN) ? vdot(dtl, reflect) / a : 0.0f;
              rsc = clamp01(rsc + (1.0f - sqrtf(larsc)));
              if (rsc > RFX_VSN)
              {
                if (L.radius > RFX_VSN)
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
      }

      sumLight = mk(h.ambient[0] * h.ambientPower + sumLight.x, h.ambient[1] * h.ambientPower + sumLight.y,
                    h.ambient[2] * h.ambientPower + sumLight.z);           // Scene.cpp:189

      V3 fin;
      if (m.type == 1)   // dielectric, Scene.cpp:192-203
      {
        const float a = rayLen * normLen;
        const float cosA = (a > RFX_VSN) ? clamp01(((ray.x * -norm.x + ray.y * -norm.y) + ray.z * -norm.z) / a) : 0.0f;
        const float rf = 0.2f + 0.8f * cubeLikePowf(1.0f - cosA);
        const float k = 1.0f - rf;
        fin = mk(((color.x * k) * sumLight.x + sumSpec.x) * mul.x, ((color.y * k) * sumLight.y + sumSpec.y) * mul.y,
                 ((color.z * k) * sumLight.z + sumSpec.z) * mul.z);
        mul = vscale(mul, rf);
      }
      else               // metal, Scene.cpp:204-214
      {
        const float rf = 0.8f;
        const float k = 1.0f - rf;
        fin = mk(((color.x * k) * sumLight.x + sumSpec.x) * mul.x, ((color.y * k) * sumLight.y + sumSpec.y) * mul.y,
                 ((color.z * k) * sumLight.z + sumSpec.z) * mul.z);
        mul = mk(mul.x * (color.x * rf), mul.y * (color.y * rf), mul.z * (color.z * rf));
      }

      pix = mk(clamp01(pix.x + fin.x), clamp01(pix.y + fin.y), clamp01(pix.z + fin.z));

      if (mul.x < 0.01f && mul.y < 0.01f && mul.z < 0.01f) break;

      origin = drop;
      ray = vadd(normalizeVec(reflect), vscale(randDir, 1.0f - m.reflectivity));   // Scene.cpp:226
    }
    else
    {
      RFX_SIG(sig, 0xFFFF);
      const V3 sky = skySample(sc, ray);
      pix = mk(clamp01(pix.x + (mul.x * sky.x) * h.env[0]), clamp01(pix.y + (mul.y * sky.y) * h.env[1]),
               clamp01(pix.z + (mul.z * sky.z) * h.env[2]));              // Scene.cpp:230-231
      break;
    }
  }
  return pix;
}

// =====================================================================================================================
// K2: Render::renderNext slice (reference Render.cpp:136-215).  One thread per pixel of the slice (grid SSAA: the
// thread walks its s*s samples in the reference's ssx, ssy order so the sum is formed in the same order), or one
// thread per block origin in block-preview mode.
// =====================================================================================================================
#ifndef RFX_BIG_THREADS
#define RFX_BIG_THREADS 128
#endif
#ifndef RFX_BIG_SMEM_SCENE
#define RFX_BIG_SMEM_SCENE 0
#endif
#ifndef RFX_BIG_MINBLOCKS
#define RFX_BIG_MINBLOCKS 8      // 64 registers: the BVH walk is latency-bound, resident warps matter more than spills (profiles/README.md)
#endif
constexpr int TRACE_THREADS = RFX_BIG_THREADS;

__global__ void __launch_bounds__(TRACE_THREADS, RFX_BIG_MINBLOCKS) k_trace(const unsigned char * __restrict__ sceneBlob, uint32_t sceneBytes,
                                                         const __grid_constant__ FrameParams fp,
                                                         const uint32_t * __restrict__ sampleStates, float * __restrict__ image,
                                                         uint32_t * __restrict__ argbOut, uint32_t * __restrict__ sigOut,
                                                         unsigned long long * __restrict__ counters, int tiled)
{
#if RFX_BIG_SMEM_SCENE
  extern __shared__ uint4 smemBlob[];
  {
    const uint4 * src = reinterpret_cast<const uint4 *>(sceneBlob);
    for (uint32_t i = threadIdx.x; i < sceneBytes / 16; i += blockDim.x) smemBlob[i] = src[i];
  }
  __syncthreads();
  const SceneView sc = makeView(reinterpret_cast<const unsigned char *>(smemBlob));
#else
  // the blob is read in place: 1024 spheres are 16 KB of float4 that stay in L1 (copying 60 KB into every CTA's shared memory
  // cost more than the loads it saved, profiles/README.md)
  const SceneView sc = makeView(sceneBlob);
#endif

  uint32_t nBounces = 0, nShadow = 0;
  const V3 eye = mk(fp.eye[0], fp.eye[1], fp.eye[2]);
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;

  if (fp.sampleNum > 0)
  {
    // row-aligned slices (whole frames, bands): warps own 4x8 pixel tiles, so the rays of a warp walk the same BVH nodes;
    // any other slice: 32 consecutive pixels of the scan order
    uint64_t p = fp.p0 + gid;
    bool inside = p < fp.p1;
    if (tiled)
    {
      const uint32_t tilesX = (fp.W + 3u) / 4u;
      const uint32_t warp = (uint32_t)(gid >> 5), lane = threadIdx.x & 31u;
      const uint32_t x = (warp % tilesX) * 4u + (lane & 3u);
      const uint32_t y = (uint32_t)(fp.p0 / fp.W) + (warp / tilesX) * 8u + (lane >> 2);
      p = (uint64_t)y * fp.W + x;
      inside = x < fp.W && p < fp.p1;
    }
    if (inside)
    {
      const uint32_t y = (uint32_t)(p / fp.W), x = (uint32_t)(p % fp.W);
      const int sn = fp.sampleNum;
      const float rx = float(x) - fp.wHalf;
      const float ry = float(y) - fp.hHalf;
      float rndx = 0, rndy = 0;
      if (fp.jitter)
      {
        uint32_t s = lcgJump(fp.seedRender, (uint32_t)(2 * (p - fp.p0)));   // two draws per pixel, Render.cpp:177-178
        s = 214013u * s + 2531011u; rndx = divExact(float((int)((s >> 16) & 0x7FFFu)), 32767.0f, RFX_RCP_32767);
        s = 214013u * s + 2531011u; rndy = divExact(float((int)((s >> 16) & 0x7FFFu)), 32767.0f, RFX_RCP_32767);
      }
      V3 fin = mk(0.0f, 0.0f, 0.0f);
      uint32_t sig = 2166136261u;
      const uint32_t * st = sampleStates + (p - fp.p0) * (uint64_t)(sn * sn);
      for (int ssx = 0; ssx < sn; ssx++)
        for (int ssy = 0; ssy < sn; ssy++)
        {
          uint32_t s = *st++;
          V3 rd;
          rngTriple(s, rd.x, rd.y, rd.z);
          const float px = (rx + float(ssx) / float(sn)) + rndx;           // Render.cpp:184
          const float py = (ry + float(ssy) / float(sn)) + rndy;
          const V3 ray = mk((px * fp.view[0] + py * fp.view[1]) + fp.rz * fp.view[2],
                            (px * fp.view[3] + py * fp.view[4]) + fp.rz * fp.view[5],
                            (px * fp.view[6] + py * fp.view[7]) + fp.rz * fp.view[8]);
          const V3 c = traceSample(sc, eye, ray, fp.reflNum, rd, nBounces, nShadow, sig);
          fin = vadd(fin, c);
        }
      const float sq = float(sn * sn);
      if (fabsf(sq) > RFX_VSN) fin = mk(fin.x / sq, fin.y / sq, fin.z / sq);   // Color::operator/=, Color.cpp:50-61
      if (image)
      {
        float * px = image + p * 3;
        if (fp.accumulate) { px[0] = px[0] + fin.x; px[1] = px[1] + fin.y; px[2] = px[2] + fin.z; }
        else { px[0] = fin.x; px[1] = fin.y; px[2] = fin.z; }
      }
      if (argbOut) argbOut[p] = packArgb(fin.x, fin.y, fin.z);
      if (sigOut) sigOut[p] = sig;
    }
  }
  else
  {
    // block preview, Render.cpp:158-173: gid enumerates block origins in scan order starting at rank fp.firstRank
    const uint32_t a = (uint32_t)(-fp.sampleNum);
    const uint32_t bw = (fp.W + a - 1) / a;
    const uint64_t k = fp.firstRank + gid;
    const uint32_t y = (uint32_t)(k / bw) * a, x = (uint32_t)(k % bw) * a;
    const uint64_t p = (uint64_t)y * fp.W + x;
    if (y < fp.H && p < fp.p1)
    {
      uint32_t s = sampleStates[gid];
      V3 rd;
      rngTriple(s, rd.x, rd.y, rd.z);
      const float px = float(x) - fp.wHalf, py = float(y) - fp.hHalf;
      const V3 ray = mk((px * fp.view[0] + py * fp.view[1]) + fp.rz * fp.view[2],
                        (px * fp.view[3] + py * fp.view[4]) + fp.rz * fp.view[5],
                        (px * fp.view[6] + py * fp.view[7]) + fp.rz * fp.view[8]);
      uint32_t sig = 2166136261u;
      const V3 c = traceSample(sc, eye, ray, fp.reflNum, rd, nBounces, nShadow, sig);
      const uint32_t ex = min(x + a, fp.W), ey = min(y + a, fp.H);
      for (uint32_t qy = y; qy < ey; qy++)
        for (uint32_t qx = x; qx < ex; qx++)
        {
          const uint64_t q = (uint64_t)qy * fp.W + qx;
          if (image) { image[q * 3] = c.x; image[q * 3 + 1] = c.y; image[q * 3 + 2] = c.z; }
          if (argbOut) argbOut[q] = packArgb(c.x, c.y, c.z);
          if (sigOut) sigOut[q] = sig;
        }
    }
  }

  // event counters: one striped atomic pair per warp
  const uint32_t wb = __reduce_add_sync(0xffffffffu, nBounces);
  const uint32_t ws = __reduce_add_sync(0xffffffffu, nShadow);
  if ((threadIdx.x & 31) == 0 && counters)
  {
    const uint32_t slot = (blockIdx.x * (TRACE_THREADS / 32) + (threadIdx.x >> 5)) & 31u;
    atomicAdd(&counters[slot * 2], (unsigned long long)wb);
    atomicAdd(&counters[slot * 2 + 1], (unsigned long long)ws);
  }
}

int launchTrace(const TraceWork & w, cudaStream_t st)
{
  const FrameParams & fp = w.fp;
  uint64_t nThreads;
  int tiled = 0;
  if (fp.sampleNum > 0)
  {
    nThreads = fp.p1 - fp.p0;
    if (fp.p0 % fp.W == 0 && fp.p1 % fp.W == 0)
    {
      tiled = 1;
      const uint64_t rows = (fp.p1 - fp.p0) / fp.W;
      nThreads = (uint64_t)((fp.W + 3u) / 4u) * ((rows + 7u) / 8u) * 32u;
    }
  }
  else
  {
    // block origins in [p0, p1): the host computed firstRank; count = originsBefore(p1) - firstRank is passed via p1 bound
    const uint32_t a = (uint32_t)(-fp.sampleNum);
    const uint32_t bw = (fp.W + a - 1) / a, bh = (fp.H + a - 1) / a;
    nThreads = (uint64_t)bw * bh - fp.firstRank;   // upper bound; threads past p1 exit
  }
  if (nThreads == 0) return 0;
  const uint32_t smem = RFX_BIG_SMEM_SCENE ? ((w.sceneBytes + 15u) & ~15u) : 0u;
  // the opt-in is per device and cheap: set it on every such launch (a process may drive several GPUs)
  if (smem > 48 * 1024 && cudaFuncSetAttribute(k_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  const uint32_t blocks = (uint32_t)((nThreads + TRACE_THREADS - 1) / TRACE_THREADS);
  k_trace<<<blocks, TRACE_THREADS, smem, st>>>(reinterpret_cast<const unsigned char *>(w.sceneBlob), smem, fp, w.sampleStates,
                                               w.image, w.argbOut, w.sigOut, w.counters, tiled);
  return 1;
}

// =====================================================================================================================
// Scene::trace for an explicit ray list (rfx_trace_rays)
// =====================================================================================================================
__global__ void __launch_bounds__(TRACE_THREADS) k_trace_rays(const unsigned char * __restrict__ sceneBlob, uint32_t sceneBytes, int n,
                                                              const float * __restrict__ origins, const float * __restrict__ rays, int reflNum,
                                                              const uint32_t * __restrict__ sampleStates, float * __restrict__ rgbOut,
                                                              unsigned long long * __restrict__ counters)
{
  extern __shared__ uint4 smemBlob[];
  {
    const uint4 * src = reinterpret_cast<const uint4 *>(sceneBlob);
    for (uint32_t i = threadIdx.x; i < sceneBytes / 16; i += blockDim.x) smemBlob[i] = src[i];
  }
  __syncthreads();
  const SceneView sc = makeView(reinterpret_cast<const unsigned char *>(smemBlob));
  uint32_t nBounces = 0, nShadow = 0, sig = 2166136261u;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    uint32_t s = sampleStates[i];
    V3 rd;
    rngTriple(s, rd.x, rd.y, rd.z);
    const V3 c = traceSample(sc, mk(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]), mk(rays[3 * i], rays[3 * i + 1], rays[3 * i + 2]),
                             reflNum, rd, nBounces, nShadow, sig);
    rgbOut[3 * i] = c.x; rgbOut[3 * i + 1] = c.y; rgbOut[3 * i + 2] = c.z;
  }
  const uint32_t wb = __reduce_add_sync(0xffffffffu, nBounces);
  const uint32_t ws = __reduce_add_sync(0xffffffffu, nShadow);
  if ((threadIdx.x & 31) == 0 && counters)
  {
    const uint32_t slot = (blockIdx.x * (TRACE_THREADS / 32) + (threadIdx.x >> 5)) & 31u;
    atomicAdd(&counters[slot * 2], (unsigned long long)wb);
    atomicAdd(&counters[slot * 2 + 1], (unsigned long long)ws);
  }
}

int launchTraceRays(const void * sceneBlob, uint32_t sceneBytes, int n, const float * origins, const float * rays, int reflNum,
                    const uint32_t * sampleStates, float * rgbOut, unsigned long long * counters, cudaStream_t st)
{
  if (n <= 0) return 0;
  const uint32_t smem = (sceneBytes + 15u) & ~15u;
  // the opt-in is per device and cheap: set it on every such launch (a process may drive several GPUs); a failure is
  // left as the sticky error the caller's cudaGetLastError() reports
  if (smem > 48 * 1024 && cudaFuncSetAttribute(k_trace_rays, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  k_trace_rays<<<(n + TRACE_THREADS - 1) / TRACE_THREADS, TRACE_THREADS, smem, st>>>(reinterpret_cast<const unsigned char *>(sceneBlob), smem, n,
                                                                                      origins, rays, reflNum, sampleStates, rgbOut, counters);
  return 1;
}

// =====================================================================================================================
// K3: imagePixel() + argb() for the whole image (reference Render.cpp:103-114, Color.cpp:114-117)
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_resolve(const float * __restrict__ image, unsigned long long nPixels, int additiveCounter,
                                                 float * __restrict__ rgbfOut, uint32_t * __restrict__ argbOut)
{
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const float div = float(additiveCounter);
  for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < nPixels; p += stride)
  {
    float r = image[p * 3], g = image[p * 3 + 1], b = image[p * 3 + 2];
    if (additiveCounter > 1 && fabsf(div) > RFX_VSN) { r = r / div; g = g / div; b = b / div; }
    if (rgbfOut) { rgbfOut[p * 3] = r; rgbfOut[p * 3 + 1] = g; rgbfOut[p * 3 + 2] = b; }
    if (argbOut) argbOut[p] = packArgb(r, g, b);
  }
}

int launchResolve(const float * image, uint64_t nPixels, int additiveCounter, float * rgbfOut, uint32_t * argbOut, cudaStream_t st)
{
  if (!nPixels) return 0;
  const uint32_t blocks = (uint32_t)((nPixels + 255) / 256 < 148ull * 16 ? (nPixels + 255) / 256 : 148ull * 16);
  k_resolve<<<blocks, 256, 0, st>>>(image, (unsigned long long)nPixels, additiveCounter, rgbfOut, argbOut);
  return 1;
}

__global__ void __launch_bounds__(256) k_clear(float * __restrict__ p, unsigned long long n)
{
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = 0.0f;
}

int launchClear(float * image, uint64_t nFloats, cudaStream_t st)
{
  if (!nFloats) return 0;
  const uint32_t blocks = (uint32_t)((nFloats + 255) / 256 < 148ull * 16 ? (nFloats + 255) / 256 : 148ull * 16);
  k_clear<<<blocks, 256, 0, st>>>(image, (unsigned long long)nFloats);
  return 1;
}

} // namespace rfx
