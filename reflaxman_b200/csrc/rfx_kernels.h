// rfx_kernels.h — launch interface between the C-ABI host layer (rfx_capi.cu) and the sm_100a kernels (rfx_kernels.cu: K1 and K3;
// rfx_trace_small.cu, rfx_trace_blob.cu: K2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rfx_types.h"

namespace rfx
{

// ---- K1: randDir stream ranking -------------------------------------------------------------------------------
// The reference draws one Vector3::randomInsideSphere per Scene::trace call from a single serial LCG with rejection
// sampling (reference Vector3.cpp:176-188, trace_math.h:34-39).  Trace call #p therefore owns the p-th ACCEPTED
// draw-triple of the stream.  K1 reproduces that in parallel from a table of the accept counts of the LCG's whole
// 2^32-state cycle (see the K1 header comment in rfx_kernels.cu).
constexpr int RNG_THREADS = 256;
constexpr int RNG_TRIPLES_PER_THREAD = 8;
constexpr int RNG_TRIPLES_PER_BLOCK = RNG_THREADS * RNG_TRIPLES_PER_THREAD;
constexpr uint32_t RNG_CLASS_BLOCKS = 699051;   // ceil(1431655766 / 2048): blocks of 2048 triples per residue class of start positions

struct RngLocate                // written by k_rng_locate, read by k_rng_rank: where on the cycle this pass starts
{
  uint32_t r0, j0;              // residue class of the start position and the block of that class holding it
  uint32_t r1, nb1;             // class after the wrap; virtual blocks [0, nb1) are blocks j0.. of class r0, the rest blocks 0.. of r1
  unsigned long long off1;      // accepted triples of class r0 before the stream's first triple
  unsigned long long seg1;      // accepted triples from the stream's first triple to the wrap
  uint32_t ownCount;            // split frames: number of virtual blocks in the list of blocks that hold ranks this GPU owns
};

struct RngWork
{
  const uint32_t * stateIn;     // device: LCG state before the first triple
  uint32_t * stateOut;          // device: LCG state after the triple that holds rank n-1
  const uint32_t * prefix;      // device [3][RNG_CLASS_BLOCKS + 1]: exclusive prefix sums of the per-block accept counts of the cycle
  RngLocate * locate;           // device scratch
  uint32_t * sampleStates;      // device out [n] (NULL: skip-only — one small CTA whatever n is)
  int * status;                 // device: set to 1 when the provisioned blocks do not hold n accepted triples
  uint64_t n;                   // accepted triples wanted
  uint32_t nBlocks;
  // optional scatter filter (split frames): only ranks r with (r / ownPeriod) % ownWorld == ownRank are stored (ownWorld = 0: all)
  uint64_t ownPeriod = 0;
  uint32_t ownWorld = 0, ownRank = 0;
  uint32_t * ownList = nullptr; // device scratch [ownListCap]: with ownWorld, k_rng_locate lists the virtual blocks that hold ranks this GPU owns and
  uint32_t ownListCap = 0;      // k_rng_rank is launched over that list only (an 8-GPU split 8K frame: 4 100 CTAs instead of 32 400, 74 -> 12 us)
};
// upper bound of the blocks that hold ranks of one GPU's strips (what ownListCap must be at least)
uint32_t rngOwnBlocksBound(uint64_t n, uint64_t ownPeriod, uint32_t ownWorld, uint32_t nBlocks);
// uploads K1's jump-ahead tables to the current device (once per context); 0 = ok
int initRngTables();
// fills the cycle's accept-count table: counts [3][RNG_CLASS_BLOCKS] scratch, prefix [3][RNG_CLASS_BLOCKS + 1]; returns kernels launched
int launchRngTable(uint32_t * counts, uint32_t * prefix, cudaStream_t st);
// K1 self-test over the LCG's whole cycle: out[0] = triples on which the integer accept test and the reference's float expression
// disagree (must be 0), out[1] = triples inside the guard band (decided by the float expression); out must be zeroed
int launchRngSelftest(unsigned long long * out, cudaStream_t st);
// number of blocks that over-provisions n accepted triples (acceptance pi/6 = 0.5236)
uint32_t rngBlocksFor(uint64_t n);
// enqueue locate (+ rank); returns the number of kernels launched
int launchRngRank(const RngWork & w, cudaStream_t st);

// ---- K2: trace + shade ----------------------------------------------------------------------------------------
// Cost-ordered tile scheduling of the fast kernel.  Path lengths differ by 20x between tiles (sky: one segment; between two
// mirror spheres: the full reflection depth), and the hardware starts CTAs in index order, so the long tiles of the image
// centre start late and the GPU drains for ~17 % of the kernel while they finish.  Every CTA therefore files itself, at its
// end, into one of four cost classes (longest path among its pixels); the next launch over the same grid starts the
// expensive classes first.  Results do not depend on the order; the first launch over a grid runs in index order.
constexpr int TILE_CLASSES = 4;
struct TileOrder
{
  const uint32_t * inLists = nullptr;   // [TILE_CLASSES][capacity] packed (tileRow << 16 | tileColumnGroup); NULL = index order
  const uint32_t * inCounts = nullptr;  // [TILE_CLASSES], sums to the grid size
  uint32_t * outLists = nullptr;        // [TILE_CLASSES][capacity], NULL = do not record
  uint32_t * outCounts = nullptr;       // [TILE_CLASSES], zero when the launch starts (the previous launch cleared it)
  uint32_t * zeroCounts = nullptr;      // [TILE_CLASSES] the NEXT launch records into: this launch clears it (no memset between frames)
  uint32_t capacity = 0;
};

// Blob scenes: candidate spheres of the FIRST query of a path (origin = eye) per screen cell of 2^shift x 2^shift pixels, rebuilt by the
// host for every camera from the primary-ray screen bounds of the spheres; cellStart == NULL: none (the query walks the hierarchy)
struct EyeGrid
{
  const uint32_t * cellStart;   // [nx * ny + 1]
  const float4 * itemSphere;    // (cx, cy, cz, r^2) per (cell, sphere)
  const int * itemIndex;        // position in the sorted sphere array
  int nx, ny, shift;
};

struct TraceWork
{
  const void * sceneBlob;       // device scene blob (SceneHeader first)
  uint32_t sceneBytes;
  FrameParams fp;
  const uint32_t * sampleStates;// K1 output for this slice (one per Scene::trace call, call order)
  float * image;                // device W*H*3 float image (row 0 = bottom), may be NULL when argbOut is set
  uint32_t * argbOut;           // optional direct ARGB target (non-additive, whole-pixel results)
  uint32_t * sigOut;            // optional per-pixel hit-path signature
  unsigned long long * counters;// device [32][2] striped {bounces, shadowRays}
  TileOrder order;              // fast constant-bank kernel only
  bool lightGrids = false;      // blob scenes: the header carries candidate grids for some light (picks the kernel instantiation)
  EyeGrid eyeGrid = { nullptr, nullptr, nullptr, 0, 0, 0 };
};
// blob scenes (rfx_trace_blob.cu).  launchTraceBlobFast: row-aligned slices (one sample per pixel, grid SSAA, additive jitter; ARGB
// and/or float image out); one-sample ARGB slices run as a wavefront pair when queue scratch is given (persistentCtas: grid of
// the queue-driven kernel in 128-thread CTAs).  Returns the
// number of kernels launched, 0 when the work does not qualify (the caller uses launchTraceBlobAny).  BVH depth is bounded by the host
// (BLOB_MAX_BVH_DEPTH): the traversal stack holds 24 entries per thread.
constexpr int BLOB_MAX_BVH_DEPTH = 22;
uint64_t blobWavePixels(const TraceWork & w);
#ifndef RFX_BLOB_WAVE_MIN_DEPTH
#define RFX_BLOB_WAVE_MIN_DEPTH 3
#endif
constexpr int BLOB_WAVE_MIN_DEPTH = RFX_BLOB_WAVE_MIN_DEPTH;       // shallower reflection limits render with the single tile kernel (the wavefront gains from depth 3 on: 2.73 against 2.77 ms, profiles/r2_grid)
// queueRecords: blobWavePixels(w) records of 64 bytes, queueCounters: 2 words
int launchTraceBlobFast(const TraceWork & w, cudaStream_t st, void * queueRecords = nullptr, uint32_t * queueCounters = nullptr, uint32_t persistentCtas = 0,
                        int firstSegments = 2,    // segments the tile kernel renders before it queues a path (0: no wavefront)
                        uint32_t bvhFloat4 = 0);  // SceneHeader::bvhFloat4 of the blob (0: no hierarchy): small hierarchies are copied to shared memory
int launchTraceBlobAny(const TraceWork & w, cudaStream_t st);     // every renderNext mode: ragged slices, block preview, signatures
// Scene::trace for a list of rays (rfx_trace_rays): rays[i] uses sampleStates[i]
int launchTraceBlobRays(const void * sceneBlob, int n, const float * origins, const float * rays, int reflNum,
                        const uint32_t * sampleStates, float * rgbOut, unsigned long long * counters, cudaStream_t st);
// small scenes: constant-bank resident; *fastGrid (optional) receives the CTA count of the fast kernel's grid, 0 when the
// general kernel ran (the caller keeps tile-order history only for fast launches)
int launchTraceSmall(const SmallScene & sc, const TraceWork & w, cudaStream_t st, uint32_t * fastGrid = nullptr);
// CTA count the fast kernel would use for this work, 0 if the work does not qualify for it
uint32_t fastGridSize(const TraceWork & w);
// per launch of the fast kernel: pixel rectangles outside of which no primary ray can hit sphere i (rect[i]) / triangle k
// (rect[SMALL_MAX_SPHERES + k]); rfx_trace_small.cu has the derivation
struct PrimaryCull
{
  int4 rect[SMALL_MAX_SPHERES + SMALL_MAX_TRIS];   // x0, x1, y0, y1 (inclusive, frame pixels); x0 > x1: no pixel
};
PrimaryCull makePrimaryCull(const SmallScene & sc, const FrameParams & fp);
// the same bounds one sphere at a time (blob scenes bin them into the screen grid of the first query, EyeGrid)
struct PrimaryCamera { double c[3][3], eye[3], rz, wHalf, hHalf, W, H; bool ok; };   // ok = false: no bounds (camera not orthonormal): whole image
PrimaryCamera makePrimaryCamera(const FrameParams & fp);
int4 primarySphereBounds(const PrimaryCamera & cam, const float4 & sphere);   // x0, x1, y0, y1 inclusive; x0 > x1: no pixel

// ---- K3: resolve (imagePixel + argb) ---------------------------------------------------------------------------
int launchResolve(const float * image, uint64_t nPixels, int additiveCounter, float * rgbfOut, uint32_t * argbOut, cudaStream_t st);
int launchClear(float * image, uint64_t nFloats, cudaStream_t st);

// host mirror of the LCG jump (n draws ahead)
uint32_t lcgJumpHost(uint32_t s, uint64_t n);

} // namespace rfx
