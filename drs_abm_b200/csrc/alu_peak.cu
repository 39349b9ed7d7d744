// Register-resident add+min stream: the ALU roofline denominator for the
// Floyd-Warshall tiles (SURVEY.md section 8d asks the build to measure one).
// Same shape as the FW inner loop (64 independent accumulators per thread, 8+8
// operands per k) but operands never leave registers, so the result is the ceiling
// of the add+min relaxation on this GPU at its running clock.  Measured on B200: all
// instruction mixes (2xFADD+FMNMX3, FADD2+FMNMX3, 2xFFMA-imm+FMNMX3, VIADDMNMX) saturate
// at ~60-63 relaxations/clk/SM although FADD alone issues at ~109 and FMNMX3 alone at
// ~61 warp-lanes/clk/SM: the pair is limited by register-operand bandwidth (7 operand
// reads per two relaxations), not by either pipe or by issue slots.
#include "drs_common.cuh"

namespace {

// Every body uses its own operand pair (a0[r], b0[c]) / (a1[r], b1[c]); no two bodies share a sum, so
// the compiler cannot merge adds (an earlier probe that reused sums across bodies over-reported by 1.4x).
template <int VAR>
__global__ void __launch_bounds__(256, 2) alu_stream_kernel(float* out, int iters, float seed) {
  float accf[64]; int acci[64];
  float a0[8], a1[8], b0[8], b1[8]; int ai[8], bi[8], ai1[8], bi1[8];
#pragma unroll
  for (int i = 0; i < 64; ++i) { accf[i] = 1e30f + i; acci[i] = DRS_I32_INF - i; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a0[i] = seed + threadIdx.x * 0.001f + i; a1[i] = seed * 1.7f + i * 0.3f;
    b0[i] = seed * 2 + i * 3 + blockIdx.x; b1[i] = seed * 0.9f + i * 1.1f;
    ai[i] = (int)a0[i] * 1000; bi[i] = (int)b0[i] * 1000; ai1[i] = (int)a1[i] * 777; bi1[i] = (int)b1[i] * 333;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {   // 2 x 64 bodies x 2 relaxations = 4 x 64 relaxations per iteration
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (VAR == 0) {          // 2 x FADD + FMNMX3
            float d, s0 = a0[r] + b0[c], s1 = a1[r] + b1[c];
            asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(accf[r * 8 + c]), "f"(s0), "f"(s1));
            accf[r * 8 + c] = d;
          } else if (VAR == 1) {   // 2 x VIADDMNMX
            acci[r * 8 + c] = __viaddmin_s32(ai1[r], bi1[c], __viaddmin_s32(ai[r], bi[c], acci[r * 8 + c]));
          } else if (VAR == 3) {   // 2 x VIADDMNMX.U16x2: two 16-bit relaxations per lane and instruction
            acci[r * 8 + c] = (int)__viaddmin_u16x2((unsigned)ai1[r], (unsigned)bi1[c],
                                                    __viaddmin_u16x2((unsigned)ai[r], (unsigned)bi[c], (unsigned)acci[r * 8 + c]));
          } else {                 // FADD2 (packed fp32x2 add, sm_100) + FMNMX3
            const float2 s = __fadd2_rn(make_float2(a0[r], a1[r]), make_float2(b0[c], b1[c]));
            float d;
            asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(accf[r * 8 + c]), "f"(s.x), "f"(s.y));
            accf[r * 8 + c] = d;
          }
        }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a0[i] += 1.0f; a1[i] += 0.25f; b0[i] += 0.5f; b1[i] += 0.125f; ai[i] += 3; bi[i] += 1; ai1[i] += 5; bi1[i] += 7;
      }
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
  drs_count_alloc();
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
  *relax_per_s = (double)grid.x * block.x * (double)iters * 4.0 * 64.0 * (VAR == 3 ? 2.0 : 1.0) / (best_ms * 1e-3);
  return DRS_OK;
}

}  // namespace

extern "C" int drs_alu_peak(int32_t variant, int32_t iters, double* relax_per_s) {
  DRS_REQUIRE(relax_per_s && iters > 0, DRS_ERR_ARG, "drs_alu_peak: bad argument");
  if (variant == 0) return run_variant<0>(iters, relax_per_s);
  if (variant == 1) return run_variant<1>(iters, relax_per_s);
  if (variant == 2) return run_variant<2>(iters, relax_per_s);
  if (variant == 3) return run_variant<3>(iters, relax_per_s);
  DRS_REQUIRE(false, DRS_ERR_ARG, "drs_alu_peak: unknown variant %d", variant);
}
