// rfx_kernels.cu — K1 and K3 of ReflaxMan's per-pixel trace-and-shade path on sm_100a (K2, the trace kernels, live in
// rfx_trace_small.cu — scenes in the constant bank — and rfx_trace_blob.cu — any scene).
//
//   K1  k_rng_table / k_rng_prefix (once), k_rng_locate / k_rng_rank   the serial rejection-sampled LCG stream, ranked from a table of its cycle
//   K3  k_resolve, k_clear                          imagePixel() divide + 8-bit ARGB pack; setImageSize's zero fill
//
// ARITHMETIC CONTRACT.  This file is compiled with --fmad=false and without any fast-math flag: every + - * below
// is an IEEE-754 binary32 round-to-nearest operation that ptxas may not fuse, / is div.rn, sqrtf is sqrt.rn, and
// denormals are kept.
//
// Reference map (path:line under /root/reference/src/common):
//   rngAccept, k_rng_*          Vector3.cpp:176-188, trace_math.h:34-39
//   k_resolve                   Render.cpp:103-114, Color.cpp:114-117
//   k_clear                     Render.cpp:67-71
#include "rfx_kernels.h"
#include "rfx_device.cuh"
#include <float.h>
#include <math.h>
#include <algorithm>

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
    if (rngAccept(s)) mask |= 1u << k;
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

// Self-test: the integer accept test with its guard band (rngAccept) against the reference's float expression (rngTriple) on
// EVERY triple of the LCG's cycle — all three residue classes of start positions, 4.3e9 triples: proof by exhaustion that K1's
// decisions are the reference's.  out[0] = triples whose decisions differ (must be 0), out[1] = triples inside the guard band.
__global__ void __launch_bounds__(RNG_THREADS) k_rng_selftest(unsigned long long * __restrict__ out)
{
  const uint32_t r = blockIdx.y, j = blockIdx.x;
  uint32_t s1 = rngThreadStart(r, j), s2 = s1;
  uint32_t differ = 0, band = 0;
#pragma unroll
  for (int k = 0; k < RNG_TRIPLES_PER_THREAD; k++)
  {
    uint32_t t = s2;
    t = 214013u * t + 2531011u; const int a = (int)((t >> 16) & 0x7FFFu);
    t = 214013u * t + 2531011u; const int b = (int)((t >> 16) & 0x7FFFu);
    t = 214013u * t + 2531011u; const int c = (int)((t >> 16) & 0x7FFFu);
    const int da = 2 * a - 32767, db = 2 * b - 32767, dc = 2 * c - 32767;
    const uint32_t S = (uint32_t)(da * da) + (uint32_t)(db * db) + (uint32_t)(dc * dc), R2 = 32767u * 32767u;
    band += (S >= R2 - 8192u && S <= R2 + 8192u) ? 1u : 0u;
    float x, y, z;
    const bool viaFloat = rngTriple(s2, x, y, z);
    const bool viaInt = rngAccept(s1);
    differ += (viaFloat != viaInt || s1 != s2) ? 1u : 0u;
  }
  differ = __reduce_add_sync(0xffffffffu, differ);
  band = __reduce_add_sync(0xffffffffu, band);
  if ((threadIdx.x & 31u) == 0u)
  {
    if (differ) atomicAdd(&out[0], (unsigned long long)differ);
    if (band) atomicAdd(&out[1], (unsigned long long)band);
  }
}

int launchRngSelftest(unsigned long long * out, cudaStream_t st)
{
  k_rng_selftest<<<dim3(RNG_CLASS_BLOCKS, 3), RNG_THREADS, 0, st>>>(out);
  return 1;
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

// virtual block v of a pass -> block j of class r, and the ranks [first, next) of the stream that start in it
struct RngBlock { uint32_t r, j; long long first, next; bool valid; };
__device__ __forceinline__ RngBlock rngVirtualBlock(uint32_t v, uint32_t r0, uint32_t j0, uint32_t r1, uint32_t nb1, unsigned long long off1,
                                                    unsigned long long seg1, const uint32_t * __restrict__ prefix)
{
  RngBlock b;
  long long off;
  if (v < nb1) { b.r = r0; b.j = j0 + v; off = (long long)off1; }
  else         { b.r = r1; b.j = v - nb1; off = -(long long)seg1; }
  b.valid = b.j < RNG_CLASS_BLOCKS;
  b.first = 0; b.next = 0;
  if (b.valid)
  {
    const uint32_t * pre = prefix + b.r * (RNG_CLASS_BLOCKS + 1);
    b.first = (long long)pre[b.j] - off; b.next = (long long)pre[b.j + 1] - off;
  }
  return b;
}

// split frames: ranks are dealt to the GPUs in strips of ownPeriod ranks, strip s to GPU s % ownWorld.  Does a block whose ranks are
// [first, next) (clipped to [0, n)) hold a rank of ownRank's — or rank n-1, whose block writes the stream-end state?
__device__ __forceinline__ bool rngBlockOwned(long long first, long long next, unsigned long long n, unsigned long long ownPeriod, uint32_t ownWorld, uint32_t ownRank)
{
  const unsigned long long lo = (unsigned long long)(first < 0 ? 0 : first), hi = (unsigned long long)(next > (long long)n ? (long long)n : next) - 1ull;
  if (hi >= n - 1) return true;
  const unsigned long long s0 = lo / ownPeriod, s1 = hi / ownPeriod;
  if (s1 - s0 + 1 >= ownWorld) return true;
  const uint32_t ahead = (ownRank + ownWorld - (uint32_t)(s0 % ownWorld)) % ownWorld;   // strips from s0 to the next one ownRank owns
  return (unsigned long long)ahead <= s1 - s0;
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
    loc->ownCount = 0;
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

// Split frame: lists the virtual blocks that hold ranks of this GPU's strips (one thread per block), so that k_rng_rank is launched over
// them only — an empty CTA still costs its launch and two dependent loads: 28 000 of them were most of the ranking time of an 8-way split.
__global__ void __launch_bounds__(RNG_THREADS) k_rng_own_list(RngLocate * __restrict__ loc, const uint32_t * __restrict__ prefix, unsigned long long n,
                                                              uint32_t nBlocks, unsigned long long ownPeriod, uint32_t ownWorld, uint32_t ownRank,
                                                              uint32_t * __restrict__ ownList, uint32_t ownListCap, int * __restrict__ status)
{
  const uint32_t v = blockIdx.x * RNG_THREADS + threadIdx.x;
  bool owned = false;
  if (v < nBlocks)
  {
    const RngBlock b = rngVirtualBlock(v, loc->r0, loc->j0, loc->r1, loc->nb1, loc->off1, loc->seg1, prefix);
    if (!b.valid) { if (v == nBlocks - 1) *status = 1; }                                   // a third class in one pass: callers chunk below that
    else
    {
      if (v == nBlocks - 1 && b.next < (long long)n) *status = 1;                          // the provisioned blocks do not reach rank n-1
      owned = b.first < (long long)n && b.next > 0 && b.next > b.first && rngBlockOwned(b.first, b.next, n, ownPeriod, ownWorld, ownRank);
    }
  }
  const uint32_t m = __ballot_sync(0xffffffffu, owned);
  if (m)
  {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&loc->ownCount, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (owned)
    {
      const uint32_t slot = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
      if (slot < ownListCap) ownList[slot] = v; else *status = 1;
    }
  }
}

__global__ void __launch_bounds__(RNG_THREADS) k_rng_rank(const RngLocate * __restrict__ loc, const uint32_t * __restrict__ prefix,
                                                          uint32_t * __restrict__ stateOut, uint32_t * __restrict__ sampleStates,
                                                          unsigned long long n, unsigned long long ownPeriod, uint32_t ownWorld, uint32_t ownRank,
                                                          int * __restrict__ status, uint32_t nBlocks, const uint32_t * __restrict__ ownList)
{
  // virtual block v of the pass -> block j of class r; ranks of the stream = table prefix + local prefix - off
  uint32_t v = blockIdx.x;
  if (ownList)
  {
    if (blockIdx.x >= min(loc->ownCount, gridDim.x)) return;     // the grid is the host's upper bound of the list
    v = ownList[blockIdx.x];
  }
  const RngBlock b = rngVirtualBlock(v, loc->r0, loc->j0, loc->r1, loc->nb1, loc->off1, loc->seg1, prefix);
  const bool lastBlock = v == nBlocks - 1;
  if (!b.valid)
  {
    if (lastBlock && threadIdx.x == 0) *status = 1;             // a third class in one pass: callers chunk below that
    return;
  }
  const uint32_t r = b.r, j = b.j;
  const long long first = b.first, next = b.next;               // ranks [first, next) start in this block
  if (lastBlock && next < (long long)n && threadIdx.x == 0) *status = 1;                  // the provisioned blocks do not reach rank n-1
  if (first >= (long long)n || next <= 0) return;                                         // uniform per CTA
  // split frames without a list (no scratch given): a CTA whose ranks all lie in strips of other GPUs has nothing to store
  if (ownWorld && !ownList && !rngBlockOwned(first, next, n, ownPeriod, ownWorld, ownRank)) return;

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

uint32_t rngOwnBlocksBound(uint64_t n, uint64_t ownPeriod, uint32_t ownWorld, uint32_t nBlocks)
{
  // a strip of ownPeriod ranks overlaps at most ownPeriod / 900 + 3 blocks (a block of 2048 triples holds 1072 +- 23 accepted ones:
  // 900 is 7 sigma below), this GPU owns every ownWorld-th strip, and the block of rank n-1 is always listed
  const uint64_t strips = (n + ownPeriod - 1) / ownPeriod, mine = (strips + ownWorld - 1) / ownWorld;
  const uint64_t bound = mine * (ownPeriod / 900 + 3) + 4;
  return (uint32_t)std::min<uint64_t>(bound, nBlocks);
}

int launchRngRank(const RngWork & w, cudaStream_t st)
{
  if (!w.sampleStates)
  {
    k_rng_locate<<<1, RNG_THREADS, 0, st>>>(w.stateIn, w.prefix, (unsigned long long)w.n, w.locate, w.stateOut, w.status);
    return 1;
  }
  const uint32_t cap = w.ownWorld ? rngOwnBlocksBound(w.n, w.ownPeriod, w.ownWorld, w.nBlocks) : 0u;
  const bool listed = w.ownWorld && w.ownList && w.ownListCap >= cap;
  k_rng_locate<<<1, RNG_THREADS, 0, st>>>(w.stateIn, w.prefix, (unsigned long long)w.n, w.locate, nullptr, w.status);
  if (listed)
    k_rng_own_list<<<(w.nBlocks + RNG_THREADS - 1) / RNG_THREADS, RNG_THREADS, 0, st>>>(w.locate, w.prefix, (unsigned long long)w.n, w.nBlocks,
                                                                                        (unsigned long long)w.ownPeriod, w.ownWorld, w.ownRank, w.ownList, cap, w.status);
  k_rng_rank<<<listed ? cap : w.nBlocks, RNG_THREADS, 0, st>>>(w.locate, w.prefix, w.stateOut, w.sampleStates, (unsigned long long)w.n,
                                                                (unsigned long long)w.ownPeriod, w.ownWorld, w.ownRank, w.status, w.nBlocks,
                                                                listed ? w.ownList : nullptr);
  return listed ? 3 : 2;
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
