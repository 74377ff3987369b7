// microbench_fp32.cu — measures the FP32 pipe ceilings the trace kernel lives under on this B200:
// FFMA, un-fused FMUL+FADD mix, packed f32x2 forms (sm_100 add/mul/fma.rn.f32x2), IEEE div.rn and sqrt.rn.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_fp32 tools/microbench_fp32.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int MODE> __global__ void __launch_bounds__(256) k(float * out, float a, float b)
{
  float x[ILP];
  unsigned long long p[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { x[i] = a + threadIdx.x * 1e-6f + i; p[i] = ((unsigned long long)__float_as_uint(x[i]) << 32) | __float_as_uint(x[i] + 1.f); }
  unsigned long long pa = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
  unsigned long long pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
  for (int it = 0; it < ITERS; it++)
  {
#pragma unroll
    for (int i = 0; i < ILP; i++)
    {
      if (MODE == 0) x[i] = __fmaf_rn(x[i], a, b);                       // FFMA
      if (MODE == 1) { x[i] = __fmul_rn(x[i], a); x[i] = __fadd_rn(x[i], b); }   // FMUL + FADD
      if (MODE == 2) x[i] = __fadd_rn(x[i], b);                          // FADD
      if (MODE == 3) x[i] = __fmul_rn(x[i], a);                          // FMUL
      if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
      if (MODE == 5) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); }
      if (MODE == 6) x[i] = __fdiv_rn(x[i], a);                          // IEEE divide
      if (MODE == 7) x[i] = __fsqrt_rn(x[i]) + b;                        // IEEE sqrt (+ FADD)
      if (MODE == 8) x[i] = __fdividef(x[i], a);                         // approx divide
      if (MODE == 9) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
      if (MODE == 10) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
      if (MODE == 11) { x[i] = __fmul_rn(x[i], a); p[i] = p[i] * 6364136223846793005ull + 1442695040888963407ull; }   // FMUL + 64-bit IMAD mix
      if (MODE == 12) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa)); x[i] = __fadd_rn(x[i], b); }   // packed mul + scalar add
      // MODE 5 above is NOT an un-fused mix: ptxas contracts the dependent mul.rn.f32x2 -> add.rn.f32x2 pair into FFMA2 (cuobjdump shows
      // 512 FFMA2 and no FMUL2/FADD2 in k<5>), also with --fmad=false.  The modes below keep the two packed ops on INDEPENDENT chains
      // (even i: multiply chain, odd i: add chain), which ptxas cannot contract: that is the real FMUL2 + FADD2 mix.
      if (MODE == 13) { if (i & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); else asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa)); }
      if (MODE == 14) { if (i & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); else x[i] = __fmul_rn(x[i], a); }   // scalar mul + packed add
      if (MODE == 15) { if (i & 1) x[i] = __fadd_rn(x[i], b); else asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa)); }   // packed mul + scalar add, independent
      if (MODE == 16) { if (i & 1) x[i] = __fadd_rn(x[i], b); else x[i] = __fmul_rn(x[i], a); }                                          // scalar mul + scalar add, independent
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
  if (s == 12345.678f) out[0] = s;
}

template <int MODE> void run(const char * name, double opsPerIter, double flopPerOp, int sms, int blocksPerSm)
{
  float * d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = sms * blocksPerSm;
  k<MODE><<<grid, 256>>>(d, 1.0000001f, 1e-7f);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++)
  {
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, 1.0000001f, 1e-7f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double threadOps = (double)grid * 256 * ITERS * ILP * opsPerIter;
  printf("{\"bench\": \"%s\", \"ms\": %.4f, \"Gops_per_s\": %.1f, \"TFLOP_per_s\": %.2f, \"warp_instr_per_clk_per_sm_at_1965MHz\": %.3f}\n",
         name, best, threadOps / best / 1e6, threadOps * flopPerOp / best / 1e9, threadOps / 32 / (best * 1e-3) / sms / 1.965e9);
  cudaFree(d);
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, khz);
  const int sms = p.multiProcessorCount;
  run<0>("ffma", 1, 2, sms, 8);
  run<1>("fmul+fadd", 2, 1, sms, 8);
  run<2>("fadd", 1, 1, sms, 8);
  run<3>("fmul", 1, 1, sms, 8);
  run<4>("ffma2 (f32x2)", 1, 4, sms, 8);
  run<5>("dependent mul.rn.f32x2 -> add.rn.f32x2 (ptxas emits FFMA2: NOT un-fused)", 2, 2, sms, 8);
  run<6>("div.rn", 1, 1, sms, 8);
  run<7>("sqrt.rn+fadd", 1, 1, sms, 8);
  run<8>("div.approx", 1, 1, sms, 8);
  run<9>("fmul2 only (f32x2)", 1, 2, sms, 8);
  run<10>("fadd2 only (f32x2)", 1, 2, sms, 8);
  run<12>("fmul2 + scalar fadd", 2, 1.5, sms, 8);
  run<13>("fmul2 | fadd2 on independent chains (true un-fused packed mix)", 1, 2, sms, 8);
  run<14>("fmul | fadd2 on independent chains", 1, 1.5, sms, 8);
  run<15>("fmul2 | fadd on independent chains", 1, 1.5, sms, 8);
  run<16>("fmul | fadd on independent chains", 1, 1, sms, 8);
  return 0;
}
