// pow_check.cu — evidence for the one arithmetic deviation from the reference's glibc build: powLikePowf
// (reflaxman_b200/csrc/rfx_device.cuh) against RN_float(pow(double, double)) of CUDA's libdevice (< 1 ulp of binary64), on
// 1.24e9 random arguments per mode drawn from the ranges Scene.cpp:175 produces, plus the special cases.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -Ireflaxman_b200/csrc -o tools/pow_check tools/pow_check.cu
// Run:   tools/pow_check > profiles/pow_check_r2.jsonl
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include "rfx_device.cuh"

using namespace rfx;

__device__ __forceinline__ uint32_t mix(uint32_t & s) { s = s * 1664525u + 1013904223u; uint32_t x = s; x ^= x >> 15; x *= 0x2c1b3c6du; x ^= x >> 12; return x; }
__device__ __forceinline__ float unit(uint32_t & s) { return (float)(mix(s) >> 8) * (1.0f / 16777216.0f); }

// mode 0: x uniform in (0, 1], y uniform in [1, 121)   (1 + 3 * refl * len / radius for the scene's lights)
// mode 1: x = 1 - 2^-k * u (close to 1: the bright core of the lobe), y in [1, 4097)
// mode 2: x log-uniform in [2^-63, 1], y in [1, 9)
__global__ void __launch_bounds__(256) k_check(int mode, int iters, unsigned long long * out)
{
  uint32_t s = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + 12345u + (uint32_t)mode * 97u;
  unsigned long long differ = 0, far = 0;
  for (int i = 0; i < iters; i++)
  {
    float x, y;
    if (mode == 0) { x = 1.0f - unit(s); y = 1.0f + 120.0f * unit(s); }
    else if (mode == 1) { x = 1.0f - ldexpf(unit(s), -(int)(mix(s) % 20u)); y = 1.0f + 4096.0f * unit(s); if (!(x > 0.0f)) x = 1.0f; }
    else { x = ldexpf(0.5f + 0.5f * unit(s), -(int)(mix(s) % 63u)); y = 1.0f + 8.0f * unit(s); }
    const float got = powLikePowf(x, y);
    const float want = (float)pow((double)x, (double)y);
    if (__float_as_uint(got) != __float_as_uint(want))
    {
      differ++;
      const int d = abs((int)__float_as_uint(got) - (int)__float_as_uint(want));
      if (d > 1) far++;
    }
  }
  atomicAdd(&out[0], differ);
  atomicAdd(&out[1], far);
}

__global__ void k_special(float * out)
{
  const float inf = __int_as_float(0x7F800000);
  out[0] = powLikePowf(1.0f, inf);            // powf: 1
  out[1] = powLikePowf(1.0f, 5.0f);           // 1
  out[2] = powLikePowf(0.5f, inf);            // 0
  out[3] = powLikePowf(0.5f, 140.0f);         // 2^-140: a float denormal, exactly representable
  out[4] = powLikePowf(1.0842022e-19f, 1.0f); // x itself
  out[5] = powLikePowf(0.99999994f, 1.0f);    // x itself
  out[6] = powLikePowf(0.25f, 0.5f);          // 0.5
}

int main()
{
  unsigned long long * d;
  cudaMalloc(&d, 16);
  const int blocks = 148 * 32, iters = 1024;
  for (int mode = 0; mode < 3; mode++)
  {
    cudaMemset(d, 0, 16);
    k_check<<<blocks, 256>>>(mode, iters, d);
    unsigned long long h[2];
    if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    printf("{\"mode\": %d, \"samples\": %llu, \"differ_from_rn_float_of_double_pow\": %llu, \"differ_by_more_than_1ulp\": %llu}\n",
           mode, (unsigned long long)blocks * 256ull * iters, h[0], h[1]);
  }
  float * f;
  cudaMalloc(&f, 32);
  k_special<<<1, 1>>>(f);
  float h[7];
  cudaMemcpy(h, f, sizeof(h), cudaMemcpyDeviceToHost);
  const float want[7] = { 1.0f, 1.0f, 0.0f, ldexpf(1.0f, -140), 1.0842022e-19f, 0.99999994f, 0.5f };
  int ok = 1;
  for (int i = 0; i < 7; i++) ok &= (h[i] == want[i]);
  printf("{\"special_cases\": [%a, %a, %a, %a, %a, %a, %a], \"all_as_powf\": %s}\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], ok ? "true" : "false");
  return ok ? 0 : 1;
}
