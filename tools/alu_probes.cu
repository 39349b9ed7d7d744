// Instruction-throughput probes for the min-plus inner loop (no common sub-expressions between bodies).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b) { float d; asm("fma.rn.f32 %0, %1, 0f3F800000, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

// VAR: 0 FADD+FMNMX | 1 2xFADD+FMNMX3 | 2 2xFFMAimm+FMNMX3 | 3 FADD2+FMNMX3 | 4 FADD only | 5 FFMAimm only | 6 FMNMX only | 7 FMNMX3 only | 8 FADD2 only
template <int VAR>
__global__ void __launch_bounds__(256, 2) probe(float* out, int iters, float seed) {
  float acc[64];
  float a0[8], a1[8], b0[8], b1[8];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 1e30f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) { a0[i] = seed + threadIdx.x * 0.001f + i; a1[i] = seed * 1.7f + i * 0.3f; b0[i] = seed * 2 + i * 3 + blockIdx.x; b1[i] = seed * 0.9f + i * 1.1f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float& x = acc[r * 8 + c];
        if (VAR == 0) { x = fminf(x, a0[r] + b0[c]); x = fminf(x, a1[r] + b1[c]); }
        else if (VAR == 1) x = fmin3(x, a0[r] + b0[c], a1[r] + b1[c]);
        else if (VAR == 2) x = fmin3(x, fma1(a0[r], b0[c]), fma1(a1[r], b1[c]));
        else if (VAR == 3) { float2 s = __fadd2_rn(make_float2(a0[r], a1[r]), make_float2(b0[c], b1[c])); x = fmin3(x, s.x, s.y); }
        else if (VAR == 4) { x = x + a0[r]; x = x + b1[c]; }
        else if (VAR == 5) { x = fma1(x, a0[r]); x = fma1(x, b1[c]); }
        else if (VAR == 6) { x = fminf(x, a0[r]); x = fminf(x, b1[c]); }
        else if (VAR == 7) { x = fmin3(x, a0[r], b1[c]); }
        else if (VAR == 8) { float2 s = __fadd2_rn(make_float2(x, a1[r]), make_float2(b0[c], b1[c])); x = s.x + 0.0f * s.y; }
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) { a0[i] += 1.0f; a1[i] += 0.25f; b0[i] += 0.5f; b1[i] += 0.125f; }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;
}

template <int VAR> void run(const char* name, double per_body, int sms, int iters, float* d) {
  probe<VAR><<<sms * 2, 256>>>(d, 16, 1.0f); CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) { CK(cudaEventRecord(e0)); probe<VAR><<<sms * 2, 256>>>(d, iters, 1.0f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  double bodies = (double)sms * 2 * 256 * (double)iters * 64.0;
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double per_s = bodies * per_body / (best * 1e-3);
  printf("{\"probe\": \"%s\", \"units_per_s\": %.4e, \"units_per_clk_per_sm_at_max_clock\": %.2f}\n", name, per_s, per_s / (sms * (double)clk * 1e3));
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 20000;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  float* d; CK(cudaMalloc(&d, 4));
  int sms = p.multiProcessorCount;
  run<0>("relax: FADD+FMNMX (2 relax per body)", 2, sms, iters, d);
  run<1>("relax: 2xFADD+FMNMX3 (2 relax per body)", 2, sms, iters, d);
  run<2>("relax: 2xFFMAimm+FMNMX3 (2 relax per body)", 2, sms, iters, d);
  run<3>("relax: FADD2+FMNMX3 (2 relax per body)", 2, sms, iters, d);
  run<4>("instr: FADD (2 per body)", 2, sms, iters, d);
  run<5>("instr: FFMA imm (2 per body)", 2, sms, iters, d);
  run<6>("instr: FMNMX (2 per body)", 2, sms, iters, d);
  run<7>("instr: FMNMX3 (1 per body)", 1, sms, iters, d);
  run<8>("instr: FADD2 (1 per body, +1 FFMA)", 1, sms, iters, d);
  return 0;
}
