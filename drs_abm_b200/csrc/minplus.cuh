// Min-plus 128x128 register-tile product shared by the blocked Floyd-Warshall (fw_kernels.cu) and the
// separator-partitioned build (sep_apsp.cu).
//
// One CTA = 256 threads, an 8x8 accumulator tile per thread, A (128 x K) and B (K x 128) streamed through a 3-stage
// cp.async ring in 32-wide k chunks.  fp32 runs on packed adds: one FADD2 + one 3-input FMNMX3 per two relaxations;
// int32 on the DPX add-min (VIADDMNMX).  The work is ALU-bound (kernel K1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "drs_b200.h"

namespace mp {

constexpr int TILE = DRS_FW_TILE;  // 128
constexpr int KC = 32;             // k chunk streamed per pipeline stage
constexpr int STAGES = 3;
constexpr int APAD = KC + 4;       // padded row of the A chunk (keeps 16 B alignment)
constexpr int NTHREADS = 256;

template <typename T> struct Semiring;
template <> struct Semiring<float> {
  static constexpr bool packed = true;    // sm_100 packed fp32x2 add (FADD2) in the tile kernel
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
  // acc = min(acc, a0+b0, a1+b1): two FADD + one 3-input FMNMX3 (sm_100+)
  static __device__ __forceinline__ float relax2(float acc, float a0, float b0, float a1, float b1) {
    float d;
    float s0 = a0 + b0, s1 = a1 + b1;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(acc), "f"(s0), "f"(s1));
    return d;
  }
  static __device__ __forceinline__ float relax1(float acc, float a, float b) { return fminf(acc, a + b); }
};
template <> struct Semiring<int> {
  static constexpr bool packed = false;
  static __device__ __forceinline__ int inf() { return DRS_I32_INF; }
  // VIADDMNMX: add and min in one DPX instruction
  static __device__ __forceinline__ int relax2(int acc, int a0, int b0, int a1, int b1) {
    return __viaddmin_s32(a1, b1, __viaddmin_s32(a0, b0, acc));
  }
  static __device__ __forceinline__ int relax1(int acc, int a, int b) { return __viaddmin_s32(a, b, acc); }
};

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<int> { using type = int4; };
template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<int> { using type = int2; };

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

constexpr size_t kTileSmemBytes = sizeof(float) * STAGES * (TILE * APAD + KC * TILE);

// Tile coordinates of accumulator element acc[r][c] of thread (tx = tid & 15, ty = tid >> 4).
// fp32 (packed): rows ty + 16 r, the four column pairs 32 j + 2 tx (keeps the k-pair-interleaved B loads conflict-free);
// int32: rows ty + 16 r, columns 4 tx .. and 64 + 4 tx ..
template <typename T, int ROWS = TILE> __device__ __forceinline__ int acc_row(int r, int ty) { return ty + (ROWS / 8) * r; }
template <typename T> __device__ __forceinline__ int acc_col(int c, int tx) {
  if (Semiring<T>::packed) return 32 * (c >> 1) + 2 * tx + (c & 1);
  return (c < 4) ? tx * 4 + c : 64 + tx * 4 + (c - 4);
}

template <typename T> __device__ __forceinline__ void acc_fill(T (&acc)[8][8], T v) {
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = v;
}

// Whole 128x128 tile at C (16-byte aligned, ldc a multiple of 4) into the accumulators.
template <typename T> __device__ __forceinline__ void acc_load(T (&acc)[8][8], const T* C, int64_t ldc, int tid) {
  using V4 = typename Vec4<T>::type;
  using V2 = typename Vec2<T>::type;
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const T* crow = C + (int64_t)(ty + 16 * r) * ldc;
    if (Semiring<T>::packed) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const V2 v = *reinterpret_cast<const V2*>(crow + 32 * j + tx * 2);
        acc[r][2 * j] = v.x; acc[r][2 * j + 1] = v.y;
      }
    } else {
      V4 lo = *reinterpret_cast<const V4*>(crow + tx * 4);
      V4 hi = *reinterpret_cast<const V4*>(crow + 64 + tx * 4);
      acc[r][0] = lo.x; acc[r][1] = lo.y; acc[r][2] = lo.z; acc[r][3] = lo.w;
      acc[r][4] = hi.x; acc[r][5] = hi.y; acc[r][6] = hi.z; acc[r][7] = hi.w;
    }
  }
}

// Accumulators to the tile at C; rows >= rows and columns outside [col_lo, cols) are not written (edge tiles store
// element-wise).
template <typename T, int ROWS = TILE>
__device__ __forceinline__ void acc_store(const T (&acc)[8][8], T* C, int64_t ldc, int rows, int cols, int tid, int col_lo = 0) {
  using V4 = typename Vec4<T>::type;
  using V2 = typename Vec2<T>::type;
  const int tx = tid & 15, ty = tid >> 4;
  if (rows >= ROWS && cols >= TILE && col_lo <= 0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      T* crow = C + (int64_t)(ty + (ROWS / 8) * r) * ldc;
      if (Semiring<T>::packed) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          V2 v; v.x = acc[r][2 * j]; v.y = acc[r][2 * j + 1];
          *reinterpret_cast<V2*>(crow + 32 * j + tx * 2) = v;
        }
      } else {
        V4 lo, hi;
        lo.x = acc[r][0]; lo.y = acc[r][1]; lo.z = acc[r][2]; lo.w = acc[r][3];
        hi.x = acc[r][4]; hi.y = acc[r][5]; hi.z = acc[r][6]; hi.w = acc[r][7];
        *reinterpret_cast<V4*>(crow + tx * 4) = lo;
        *reinterpret_cast<V4*>(crow + 64 + tx * 4) = hi;
      }
    }
    return;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int row = ty + (ROWS / 8) * r;
    if (row >= rows) continue;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int col = acc_col<T>(c, tx);
      if (col < cols && col >= col_lo) C[(int64_t)row * ldc + col] = acc[r][c];
    }
  }
}

// The transposed tile: element (row, col) of the accumulators goes to Ct[col * ldc + row] for col in [col_lo, col_hi) and
// row < rows, staged through shared memory (TILE x (TILE + 1) elements of smem_raw, free once tile_accumulate has
// returned) so that the global stores run along rows of Ct.
// staged row v (a column of the tile, ROWS values, pitch ROWS + 4 so that rows stay 16-byte aligned) is rotated by 4 tx(v)
// elements (tx = the lane that owns column v): the lanes of a staging store spread over the banks, and a row segment of
// four stays one aligned 128-bit shared load
template <typename T> __device__ __forceinline__ int stage_rot(int v) { return Semiring<T>::packed ? 4 * ((v >> 1) & 15) : 4 * ((v >> 2) & 15); }
template <typename T, int ROWS = TILE>
__device__ __forceinline__ void acc_store_transposed(const T (&acc)[8][8], unsigned char* smem_raw, T* Ct, int64_t ldc, int col_lo, int col_hi,
                                                     int rows, int tid) {
  using V4 = typename Vec4<T>::type;
  constexpr int NT = 2 * ROWS, PITCH = ROWS + 4, LPR = ROWS / 4, RPW = 32 / LPR;
  T* s = reinterpret_cast<T*>(smem_raw);
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) s[acc_col<T>(c, tx) * PITCH + ((acc_row<T, ROWS>(r, ty) + 4 * tx) & (ROWS - 1))] = acc[r][c];
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int v = warp * RPW + lane / LPR; v < TILE; v += (NT / 32) * RPW) {
    if (v < col_lo || v >= col_hi) continue;
    T* crow = Ct + (int64_t)v * ldc;
    const int u = (lane % LPR) * 4;
    const V4 q4 = *reinterpret_cast<const V4*>(s + v * PITCH + ((u + stage_rot<T>(v)) & (ROWS - 1)));
    if (u + 3 < rows) *reinterpret_cast<V4*>(crow + u) = q4;     // whole 16-byte segments of a transposed row
    else {
      if (u < rows) crow[u] = q4.x;
      if (u + 1 < rows) crow[u + 1] = q4.y;
      if (u + 2 < rows) crow[u + 2] = q4.z;
    }
  }
}

// Phase 1 of the blocked Floyd-Warshall: closes one 128x128 diagonal tile D (row-major, leading dimension ld, 16-byte
// aligned) in place.  One CTA of DIAG_THREADS threads; the tile lives in registers (4x8 per thread: rows 4 ty .., columns
// 8 tx ..) and only the pivot row and the pivot column of the next step pass through shared memory (double-buffered: one
// barrier per step).  Row k and column k do not change in step k (the diagonal is 0 and the weights are positive), so
// publishing them one step ahead is exact.  The step loop is unrolled by 8 so that the published register is a
// compile-time choice.  `store_peers(offset, lo, hi)`: the closed tile also goes to the peers' buffers (fused broadcast
// of the pivot step).
constexpr int DIAG_THREADS = 512;
template <typename T, typename PeerFn>
__device__ __forceinline__ void fw_close_diag_tile(T* D, int64_t ld, int tid, PeerFn store_peers) {
  using V4 = typename Vec4<T>::type;
  __shared__ __align__(16) T rowb[2][TILE];
  __shared__ __align__(16) T colb[2][TILE];
  const int tx = tid & 15, ty = tid >> 4;   // ty in [0, 32)
  T c[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const V4 lo = *reinterpret_cast<const V4*>(D + (int64_t)(ty * 4 + r) * ld + tx * 8);
    const V4 hi = *reinterpret_cast<const V4*>(D + (int64_t)(ty * 4 + r) * ld + tx * 8 + 4);
    c[r][0] = lo.x; c[r][1] = lo.y; c[r][2] = lo.z; c[r][3] = lo.w;
    c[r][4] = hi.x; c[r][5] = hi.y; c[r][6] = hi.z; c[r][7] = hi.w;
  }
  if (ty == 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) rowb[0][tx * 8 + q] = c[0][q];
  }
  if (tx == 0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) colb[0][ty * 4 + r] = c[r][0];
  }
  __syncthreads();
#pragma unroll 1
  for (int kb = 0; kb < TILE / 8; ++kb) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int k = kb * 8 + kk, cur = kk & 1;
      T bk[8], a[4];
      {
        const V4 b0 = *reinterpret_cast<const V4*>(&rowb[cur][tx * 8]), b1 = *reinterpret_cast<const V4*>(&rowb[cur][tx * 8 + 4]);
        const V4 a0 = *reinterpret_cast<const V4*>(&colb[cur][ty * 4]);
        bk[0] = b0.x; bk[1] = b0.y; bk[2] = b0.z; bk[3] = b0.w; bk[4] = b1.x; bk[5] = b1.y; bk[6] = b1.z; bk[7] = b1.w;
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 8; ++q) c[r][q] = Semiring<T>::relax1(c[r][q], a[r], bk[q]);
      // publish row k + 1 and column k + 1 (static register choice: (kk + 1) & 3 and (kk + 1) & 7)
      const int kn = k + 1;
      if (kn < TILE) {
        if ((kn >> 2) == ty) {
#pragma unroll
          for (int q = 0; q < 8; ++q) rowb[cur ^ 1][tx * 8 + q] = c[(kk + 1) & 3][q];
        }
        if ((kn >> 3) == tx) {
#pragma unroll
          for (int r = 0; r < 4; ++r) colb[cur ^ 1][ty * 4 + r] = c[r][(kk + 1) & 7];
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    V4 lo, hi;
    lo.x = c[r][0]; lo.y = c[r][1]; lo.z = c[r][2]; lo.w = c[r][3];
    hi.x = c[r][4]; hi.y = c[r][5]; hi.z = c[r][6]; hi.w = c[r][7];
    const int64_t off = (int64_t)(ty * 4 + r) * ld + tx * 8;
    *reinterpret_cast<V4*>(D + off) = lo;
    *reinterpret_cast<V4*>(D + off + 4) = hi;
    store_peers(off, lo, hi);
  }
}

// acc = min(acc, A (x) B) over kcols inner indices (a multiple of 16), streamed in 32-wide chunks; a last half chunk is
// loaded whole (the 16 extra k columns of A / rows of B must be readable) and half of it used.  A: 128 rows x kcols
// (16-byte aligned rows), B: kcols rows x 128 columns (16-byte aligned rows for int32, 4-byte for fp32).  smem_raw: kTileSmemBytes of dynamic shared memory.
// Returns with every copy drained and every thread past its last shared-memory read, so the ring can be refilled at once
// (the multi-pass tiles of the separator build call it once per pass).
// ROWS = rows of the tile (128: 256 threads, or 64: 128 threads and half-height tiles, more and smaller CTAs per SM), NSTG =
// stages of the ring (3, or 2 so that four 64-row CTAs fit an SM).
template <int ROWS, int NSTG> struct TileCfg {
  static constexpr int NT = 2 * ROWS;
  static constexpr size_t ring_bytes = sizeof(float) * NSTG * (ROWS * APAD + KC * TILE);
  static constexpr size_t stage_bytes = sizeof(float) * TILE * (ROWS + 4);   // acc_store_transposed
  static constexpr size_t smem_bytes = ring_bytes > stage_bytes ? ring_bytes : stage_bytes;
};
template <typename T, int ROWS = TILE, int NSTG = STAGES>
__device__ __forceinline__ void tile_accumulate(T (&acc)[8][8], const T* __restrict__ A, int64_t lda, const T* __restrict__ B,
                                                int64_t ldb, int kcols, unsigned char* smem_raw, int tid) {
  using V4 = typename Vec4<T>::type;
  using V2 = typename Vec2<T>::type;
  constexpr int NT = 2 * ROWS, RG = ROWS / 8;
  T(*As)[ROWS][APAD] = reinterpret_cast<T(*)[ROWS][APAD]>(smem_raw);
  T(*Bs)[KC][TILE] = reinterpret_cast<T(*)[KC][TILE]>(smem_raw + sizeof(T) * NSTG * ROWS * APAD);
  const int tx = tid & 15, ty = tid >> 4;
  constexpr bool kPacked = Semiring<T>::packed;
  if (kcols <= 0) return;
  const int nchunk = (kcols + KC - 1) / KC;
  const bool half_tail = (kcols & (KC - 1)) != 0;

  auto load_chunk = [&](int kc, int stage) {
    const int k0 = kc * KC;
#pragma unroll
    for (int q = 0; q < ROWS * 8 / NT; ++q) {   // A chunk: ROWS rows x 32 k
      const int v = tid + NT * q;
      const int row = v >> 3, cv = v & 7;
      cp_async16(&As[stage][row][cv * 4], A + (int64_t)row * lda + k0 + cv * 4);
    }
    if (kPacked) {
      // fp32: B is stored k-pair interleaved, Bs2[k/2][column] = (b[k][column], b[k+1][column]), so that one
      // 8-byte shared load yields the operand pair of a packed add (FADD2) -- 4-byte cp.async, coalesced per row
      float* Bf = reinterpret_cast<float*>(&Bs[stage][0][0]);
#pragma unroll
      for (int q = 0; q < KC * TILE / NT; ++q) {
        const int e = tid + NT * q;
        const int row = e >> 7, col = e & 127;
        cp_async4(Bf + ((row >> 1) * TILE + col) * 2 + (row & 1), B + (int64_t)(k0 + row) * ldb + col);
      }
    } else {
#pragma unroll
      for (int q = 0; q < KC * TILE / 4 / NT; ++q) {   // B chunk: 32 k x 128 columns
        const int v = tid + NT * q;
        const int row = v >> 5, cv = v & 31;
        cp_async16(&Bs[stage][row][cv * 4], B + (int64_t)(k0 + row) * ldb + cv * 4);
      }
    }
  };

  // NSTG == 3: two chunks in flight ahead of the one being used, the ring refilled without a second barrier (the FW tiles:
  // four chunks).  NSTG == 2 (short tiles, two or three chunks): both stages are requested at entry -- the latency of the
  // first loads is what the CTA waits for -- and a third chunk reuses its stage after a second barrier.
  constexpr int AHEAD = NSTG == 2 ? 2 : NSTG - 1;
#pragma unroll
  for (int q = 0; q < AHEAD; ++q) {
    if (q < nchunk) load_chunk(q, q);
    cp_async_commit();
  }

#pragma unroll 1
  for (int kc = 0; kc < nchunk; ++kc) {
    const int stage = kc % NSTG;
    cp_async_wait<AHEAD - 1>();
    __syncthreads();
    if (NSTG != 2) {
      if (kc + NSTG - 1 < nchunk) load_chunk(kc + NSTG - 1, (kc + NSTG - 1) % NSTG);
      cp_async_commit();
    }
    const bool full = !(half_tail && kc == nchunk - 1);
    if (kPacked) {
      const float2* Bp = reinterpret_cast<const float2*>(&Bs[stage][0][0]);
#pragma unroll
      for (int kk = 0; kk < KC; kk += 2) {
        if (kk == KC / 2 && !full) break;
        float2 a[8], b[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const float2*>(&As[stage][ty + RG * r][kk]);
        const float2* brow = Bp + (kk >> 1) * TILE;
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // columns 32j + 2tx, +1: 16 lanes x 16 B contiguous per load
          const float4 q = *reinterpret_cast<const float4*>(brow + 32 * j + tx * 2);
          b[2 * j] = make_float2(q.x, q.y); b[2 * j + 1] = make_float2(q.z, q.w);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float2 sum = __fadd2_rn(a[r], b[c]);
            float d;
            asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"((float)acc[r][c]), "f"(sum.x), "f"(sum.y));
            acc[r][c] = (T)d;
          }
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < KC; kk += 2) {   // two k per step: acc = min3(acc, a0+b0, a1+b1)
        if (kk == KC / 2 && !full) break;
        V2 a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const V2*>(&As[stage][ty + RG * r][kk]);
        const V4 b0l = *reinterpret_cast<const V4*>(&Bs[stage][kk][tx * 4]);
        const V4 b0h = *reinterpret_cast<const V4*>(&Bs[stage][kk][64 + tx * 4]);
        const V4 b1l = *reinterpret_cast<const V4*>(&Bs[stage][kk + 1][tx * 4]);
        const V4 b1h = *reinterpret_cast<const V4*>(&Bs[stage][kk + 1][64 + tx * 4]);
        const T b0[8] = {b0l.x, b0l.y, b0l.z, b0l.w, b0h.x, b0h.y, b0h.z, b0h.w};
        const T b1[8] = {b1l.x, b1l.y, b1l.z, b1l.w, b1h.x, b1h.y, b1h.z, b1h.w};
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[r][c] = Semiring<T>::relax2(acc[r][c], a[r].x, b0[c], a[r].y, b1[c]);
      }
    }
    if (NSTG == 2) {
      if (kc + 2 < nchunk) {   // every thread is done reading this stage before it is refilled
        __syncthreads();
        load_chunk(kc + 2, stage);
      }
      cp_async_commit();
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

}  // namespace mp

// ---------------------------------------------------------------------------------------------------------------------------
// 16-bit form of the tile product for exact small-integer maps: VIADDMNMX.U16x2 relaxes two 16-bit accumulators per lane and
// instruction (measured on B200: 123 relaxations/clk/SM against 62 for the 32-bit forms, drs_alu_peak variant 3).
// Values are distances below U16_INF = 0x7fff; U16_INF and above mean "unreachable": the sum of two operands never wraps,
// and a sum with an unreachable operand stays >= U16_INF, so it never beats a finite accumulator.
// A: 128 rows x kcols 32-bit words, each the 16-bit value in BOTH halves; B: kcols rows x 128 columns of 16-bit values.
// Thread (tx = tid & 15, ty = tid >> 4) owns rows ty + 16 r and the eight columns 8 tx ..; acc2[r][p] = columns
// 8 tx + 2 p (low half) and 8 tx + 2 p + 1 (high half).
namespace mp {

constexpr unsigned U16_INF = 0x7fffu;
constexpr int BS16 = TILE / 2;   // 32-bit words per B row in shared memory
constexpr int STAGES16 = 2;      // two chunks in flight
// ROWS = rows of the output tile (128, or 64: half-height tiles, twice as many and half as large CTAs per SM, which overlap
// each other's load, compute and store phases more finely); a tile is computed by 2 ROWS threads; thread (tx = tid & 15,
// ty = tid >> 4) owns rows ty + (ROWS / 8) r and the eight columns 8 tx ..
template <int ROWS> struct Tile16 {
  static constexpr int NT = 2 * ROWS, RG = ROWS / 8, PITCH = ROWS + 4;
  static constexpr size_t ring_bytes = sizeof(uint32_t) * STAGES16 * (ROWS * APAD + KC * BS16);
  static constexpr size_t stage_bytes = sizeof(float) * TILE * PITCH;   // the transposed tile staged for the mirrored store
  static constexpr size_t smem_bytes = ring_bytes > stage_bytes ? ring_bytes : stage_bytes;
};
constexpr size_t kTile16SmemBytes = Tile16<TILE>::smem_bytes;

template <int ROWS>
__device__ __forceinline__ void tile_accumulate_u16(uint32_t (&acc2)[8][4], const uint32_t* __restrict__ A, int64_t lda,
                                                    const uint16_t* __restrict__ B, int64_t ldb, int kcols, unsigned char* smem_raw, int tid) {
  constexpr int NT = Tile16<ROWS>::NT, RG = Tile16<ROWS>::RG;
  uint32_t(*As)[ROWS][APAD] = reinterpret_cast<uint32_t(*)[ROWS][APAD]>(smem_raw);
  uint32_t(*Bs)[KC][BS16] = reinterpret_cast<uint32_t(*)[KC][BS16]>(smem_raw + sizeof(uint32_t) * STAGES16 * ROWS * APAD);
  const int tx = tid & 15, ty = tid >> 4;
  if (kcols <= 0) return;
  const int nchunk = (kcols + KC - 1) / KC;
  const bool half_tail = (kcols & (KC - 1)) != 0;

  auto load_chunk = [&](int kc, int stage) {
    const int k0 = kc * KC;
#pragma unroll
    for (int q = 0; q < ROWS * 8 / NT; ++q) {   // A chunk: ROWS rows x 32 k words
      const int v = tid + NT * q;
      const int row = v >> 3, cv = v & 7;
      cp_async16(&As[stage][row][cv * 4], A + (int64_t)row * lda + k0 + cv * 4);
    }
#pragma unroll
    for (int q = 0; q < KC * 16 / NT; ++q) {    // B chunk: 32 k x 128 columns of 16 bits = 16 x 16 bytes per row
      const int v = tid + NT * q;
      const int row = v >> 4, cv = v & 15;
      cp_async16(&Bs[stage][row][cv * 4], B + (int64_t)(k0 + row) * ldb + cv * 8);
    }
  };

  // both stages are free at entry: the first two chunks are requested at once (a tile has two or three chunks, and the
  // latency of its first loads is what its CTA waits for)
  load_chunk(0, 0);
  cp_async_commit();
  if (nchunk > 1) load_chunk(1, 1);
  cp_async_commit();

#pragma unroll 1
  for (int kc = 0; kc < nchunk; ++kc) {
    const int stage = kc % STAGES16;
    cp_async_wait<1>();            // everything but the most recent group: chunk kc has landed
    __syncthreads();
    const bool full = !(half_tail && kc == nchunk - 1);
#pragma unroll
    for (int kk = 0; kk < KC; kk += 2) {
      if (kk == KC / 2 && !full) break;
      uint2 a[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const uint2*>(&As[stage][ty + RG * r][kk]);
      const uint4 b0 = *reinterpret_cast<const uint4*>(&Bs[stage][kk][tx * 4]);
      const uint4 b1 = *reinterpret_cast<const uint4*>(&Bs[stage][kk + 1][tx * 4]);
      const uint32_t bb0[4] = {b0.x, b0.y, b0.z, b0.w}, bb1[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int p = 0; p < 4; ++p)
          acc2[r][p] = __viaddmin_u16x2(a[r].y, bb1[p], __viaddmin_u16x2(a[r].x, bb0[p], acc2[r][p]));
    }
    if (kc + STAGES16 < nchunk) {  // a third or later chunk reuses this stage once every thread is done reading it
      __syncthreads();
      load_chunk(kc + STAGES16, stage);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();
  __syncthreads();
}

__device__ __forceinline__ float u16_to_dist(unsigned v) { return v >= U16_INF ? __int_as_float(0x7f800000) : (float)v; }
__device__ __forceinline__ unsigned dist_to_u16(float d) { return d < (float)U16_INF ? (unsigned)d : U16_INF; }

// packed accumulators as fp32 distances into the tile at C: rows >= rows and columns outside [col_lo, cols) are not written
template <int ROWS>
__device__ __forceinline__ void acc16_store(const uint32_t (&acc2)[8][4], float* C, int64_t ldc, int rows, int cols, int tid, int col_lo) {
  const int tx = tid & 15, ty = tid >> 4;
  const bool whole = rows >= ROWS && cols >= TILE && col_lo <= 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int row = ty + Tile16<ROWS>::RG * r;
    float v[8];
#pragma unroll
    for (int p = 0; p < 4; ++p) { v[2 * p] = u16_to_dist(acc2[r][p] & 0xffffu); v[2 * p + 1] = u16_to_dist(acc2[r][p] >> 16); }
    float* crow = C + (int64_t)row * ldc + tx * 8;
    if (whole) {
      *reinterpret_cast<float4*>(crow) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(crow + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else if (row < rows) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int col = tx * 8 + q;
        if (col < cols && col >= col_lo) crow[q] = v[q];
      }
    }
  }
}

// the transposed tile as fp32 (see acc_store_transposed).  Staged row v (a column of the tile, ROWS values) is rotated by
// 4 (v / 8) elements: the 16 lanes of a staging store then fall into 8 banks instead of 4, and a row segment of four is
// still one aligned 128-bit shared load followed by one 128-bit global store.
template <int ROWS>
__device__ __forceinline__ void acc16_store_transposed(const uint32_t (&acc2)[8][4], unsigned char* smem_raw, float* Ct, int64_t ldc,
                                                       int col_lo, int col_hi, int rows, int tid) {
  constexpr int NT = Tile16<ROWS>::NT, RG = Tile16<ROWS>::RG, PITCH = Tile16<ROWS>::PITCH;
  constexpr int LPR = ROWS / 4, RPW = 32 / LPR;   // lanes per staged row, staged rows per warp instruction
  float* s = reinterpret_cast<float*>(smem_raw);
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int pos = (ty + RG * r + 4 * tx) & (ROWS - 1);
      s[(tx * 8 + 2 * p) * PITCH + pos] = u16_to_dist(acc2[r][p] & 0xffffu);
      s[(tx * 8 + 2 * p + 1) * PITCH + pos] = u16_to_dist(acc2[r][p] >> 16);
    }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int v = warp * RPW + lane / LPR; v < TILE; v += (NT / 32) * RPW) {
    if (v < col_lo || v >= col_hi) continue;
    float* crow = Ct + (int64_t)v * ldc;
    const int u = (lane % LPR) * 4;
    const float4 q4 = *reinterpret_cast<const float4*>(s + v * PITCH + ((u + 4 * (v >> 3)) & (ROWS - 1)));
    if (u + 3 < rows) *reinterpret_cast<float4*>(crow + u) = q4;
    else {
      if (u < rows) crow[u] = q4.x;
      if (u + 1 < rows) crow[u + 1] = q4.y;
      if (u + 2 < rows) crow[u + 2] = q4.z;
    }
  }
}

}  // namespace mp
