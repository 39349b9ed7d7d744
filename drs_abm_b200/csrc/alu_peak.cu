// Register-resident add+min stream: the ALU roofline denominator for the
// Floyd-Warshall tiles (SURVEY.md section 8d asks the build to measure one).
// Same shape as the FW inner loop (64 independent accumulators per thread, 8+8
// operands) but operands never leave registers, so the result is the issue-rate
// ceiling of the add+min instruction pair on this GPU at its running clock.
#include "drs_common.cuh"

namespace {

template <int VAR>
__global__ void __launch_bounds__(256, 2) alu_stream_kernel(float* out, int iters, float seed) {
  float accf[64]; int acci[64];
  float af[8], bf[8]; int ai[8], bi[8];
#pragma unroll
  for (int i = 0; i < 64; ++i) { accf[i] = 1e30f; acci[i] = DRS_I32_INF; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    af[i] = seed + threadIdx.x * 0.001f + i; bf[i] = seed * 2 + i * 3 + blockIdx.x;
    ai[i] = (int)af[i] * 1000; bi[i] = (int)bf[i] * 1000;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (VAR == 0) {          // FADD + FMNMX
            accf[r * 8 + c] = fminf(accf[r * 8 + c], af[r] + bf[c]);
          } else if (VAR == 1) {   // VIADDMNMX
            acci[r * 8 + c] = __viaddmin_s32(ai[r], bi[c], acci[r * 8 + c]);
          } else {                 // 2 x FADD + FMNMX3 (two relaxations per body)
            if (kk & 1) continue;
            float d, s0 = af[r] + bf[c], s1 = af[(r + 1) & 7] + bf[(c + 3) & 7];
            asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(accf[r * 8 + c]), "f"(s0), "f"(s1));
            accf[r * 8 + c] = d;
          }
        }
#pragma unroll
      for (int i = 0; i < 8; ++i) { af[i] += 1.0f; bf[i] += 0.5f; ai[i] += 3; bi[i] += 1; }
    }
  }
  float s = 0; int si = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) { s += accf[i]; si += acci[i]; }
  if (s + (float)si == 123.456f) out[0] = s;   // never true; keeps the accumulators live
}

template <int VAR> int run_variant(int iters, double* relax_per_s) {
  int dev = 0, sms = 0;
  DRS_CUDA(cudaGetDevice(&dev));
  DRS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* d_out = nullptr;
  DRS_CUDA(cudaMalloc((void**)&d_out, 4));
  cudaEvent_t e0, e1;
  DRS_CUDA(cudaEventCreate(&e0));
  DRS_CUDA(cudaEventCreate(&e1));
  dim3 grid(sms * 2), block(256);
  alu_stream_kernel<VAR><<<grid, block>>>(d_out, 64, 1.0f);
  DRS_CUDA(cudaDeviceSynchronize());
  double best_ms = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    DRS_CUDA(cudaEventRecord(e0));
    alu_stream_kernel<VAR><<<grid, block>>>(d_out, iters, 1.0f);
    DRS_CUDA(cudaEventRecord(e1));
    DRS_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    DRS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best_ms) best_ms = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
  *relax_per_s = (double)grid.x * block.x * (double)iters * 4.0 * 64.0 / (best_ms * 1e-3);
  return DRS_OK;
}

}  // namespace

extern "C" int drs_alu_peak(int32_t variant, int32_t iters, double* relax_per_s) {
  DRS_REQUIRE(relax_per_s && iters > 0, DRS_ERR_ARG, "drs_alu_peak: bad argument");
  if (variant == 0) return run_variant<0>(iters, relax_per_s);
  if (variant == 1) return run_variant<1>(iters, relax_per_s);
  if (variant == 2) return run_variant<2>(iters, relax_per_s);
  DRS_REQUIRE(false, DRS_ERR_ARG, "drs_alu_peak: unknown variant %d", variant);
}
