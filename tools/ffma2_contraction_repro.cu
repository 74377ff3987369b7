// ffma2_contraction_repro.cu — minimal repro: ptxas 12.9 contracts a DEPENDENT mul.rn.f32x2 -> add.rn.f32x2 pair into one FFMA2
// (a single rounding), also under --fmad=false, although both PTX instructions carry the .rn modifier that forbids contraction for
// their scalar forms.  That breaks this repo's arithmetic contract (every product and every sum of the reference is rounded
// separately), so packed f32x2 arithmetic cannot be used where a sum consumes a packed product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -cubin -o /tmp/r.cubin tools/ffma2_contraction_repro.cu
//   cuobjdump -sass /tmp/r.cubin | grep -E "FMUL2|FADD2|FFMA2"
//     k_dependent:   FMUL2 + FFMA2      <- the second product was fused into the sum (expected: 2 x FMUL2 + FADD2)
//     k_independent: FMUL2 + FADD2      <- no data dependence, nothing to fuse
//   ptxas -O0 keeps FMUL2, FMUL2, FADD2 in k_dependent; -O1 .. -O3 with --fmad false all emit FFMA2.
//
// And there would be nothing to gain if it worked: tools/microbench_fp32.cu runs FMUL2 and FADD2 on independent chains (the real
// un-fused packed mix) at 1.97 warp-instructions per clock per SM = 36.6 TFLOP/s, the same flop rate as scalar FMUL + FADD at
// 3.84 per clock (35.8 TFLOP/s).  Packed operations halve the issue slots, not the time on the FP32 pipe (profiles/microbench_r2.jsonl;
// round 1's "73 TFLOP/s un-fused mix" row was this very contraction: its SASS holds 512 FFMA2 and no FMUL2 / FADD2).
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

__global__ void k_dependent(const unsigned long long * in, unsigned long long * out)
{
  const unsigned long long a = in[threadIdx.x], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64], d = in[threadIdx.x + 96];
  out[threadIdx.x] = add2(mul2(a, b), mul2(c, d));      // (a*b) + (c*d): three roundings per half in PTX, two in the SASS
}

__global__ void k_independent(const unsigned long long * in, unsigned long long * out)
{
  const unsigned long long a = in[threadIdx.x], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64], d = in[threadIdx.x + 96];
  out[threadIdx.x] = mul2(a, b);
  out[threadIdx.x + 32] = add2(c, d);
}
