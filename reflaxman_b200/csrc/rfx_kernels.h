// rfx_kernels.h — launch interface between the C-ABI host layer (rfx_capi.cu) and the sm_100a kernels (rfx_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rfx_types.h"

namespace rfx
{

// ---- K1: randDir stream ranking -------------------------------------------------------------------------------
// The reference draws one Vector3::randomInsideSphere per Scene::trace call from a single serial LCG with rejection
// sampling (reference Vector3.cpp:176-188, trace_math.h:34-39).  Trace call #p therefore owns the p-th ACCEPTED
// draw-triple of the stream.  K1 reproduces that in parallel: LCG jump-ahead to every triple, accept flag, exclusive
// scan, scatter of the LCG state that precedes each accepted triple.
constexpr int RNG_THREADS = 256;
constexpr int RNG_TRIPLES_PER_THREAD = 8;
constexpr int RNG_TRIPLES_PER_BLOCK = RNG_THREADS * RNG_TRIPLES_PER_THREAD;

struct RngWork
{
  const uint32_t * stateIn;     // device: LCG state before the first triple
  uint32_t * stateOut;          // device: LCG state after the triple that holds rank n-1
  uint32_t * blockCounts;       // device scratch [nBlocks]
  uint32_t * blockOffsets;      // device scratch [nBlocks]
  uint8_t * acceptMasks;        // device scratch [nBlocks * RNG_THREADS]: 8 accept bits per thread, count -> scatter
  uint32_t * sampleStates;      // device out [n] (may be NULL: skip-only)
  int * status;                 // device: set to 1 when fewer than n triples were accepted in nBlocks blocks
  uint64_t n;                   // accepted triples wanted
  uint32_t nBlocks;
  // optional scatter filter (split frames): only ranks r with (r / ownPeriod) % ownWorld == ownRank are stored (ownWorld = 0: all)
  uint64_t ownPeriod = 0;
  uint32_t ownWorld = 0, ownRank = 0;
};
// uploads K1's jump-ahead tables to the current device (once per context); 0 = ok
int initRngTables();
// number of blocks that over-provisions n accepted triples (acceptance pi/6 = 0.5236)
uint32_t rngBlocksFor(uint64_t n);
// enqueue count + scan + scatter; returns the number of kernels launched
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
  uint32_t * outCounts = nullptr;       // [TILE_CLASSES], zeroed before the launch
  uint32_t capacity = 0;
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
};
int launchTrace(const TraceWork & w, cudaStream_t st);                              // any scene: shared-memory resident blob
// small scenes: constant-bank resident; *fastGrid (optional) receives the CTA count of the fast kernel's grid, 0 when the
// general kernel ran (the caller keeps tile-order history only for fast launches)
int launchTraceSmall(const SmallScene & sc, const TraceWork & w, cudaStream_t st, uint32_t * fastGrid = nullptr);
// CTA count the fast kernel would use for this work, 0 if the work does not qualify for it
uint32_t fastGridSize(const TraceWork & w);

// Scene::trace for a list of rays (rfx_trace_rays): rays[i] uses sampleStates[i]
int launchTraceRays(const void * sceneBlob, uint32_t sceneBytes, int n, const float * origins, const float * rays, int reflNum,
                    const uint32_t * sampleStates, float * rgbOut, unsigned long long * counters, cudaStream_t st);

// ---- K3: resolve (imagePixel + argb) ---------------------------------------------------------------------------
int launchResolve(const float * image, uint64_t nPixels, int additiveCounter, float * rgbfOut, uint32_t * argbOut, cudaStream_t st);
int launchClear(float * image, uint64_t nFloats, cudaStream_t st);

// host mirror of the LCG jump (n draws ahead)
uint32_t lcgJumpHost(uint32_t s, uint64_t n);

} // namespace rfx
