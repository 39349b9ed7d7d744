// Batched single-source shortest paths over the device CSR (K2 in SURVEY.md section 2):
// delta-stepping in its near/far form, one CTA per source, many sources in flight.
//
// Serves (a) networks whose N x N matrix does not fit in HBM (BASELINE config C5: rows
// for a batch of sources only) and (b) the all-pairs build of SPARSE road networks,
// where N Dijkstra-like searches do ~N*E*c edge relaxations instead of the N^3 of
// Floyd-Warshall.  It replaces the same igraph Dijkstra calls as the FW path
// (lib/RoutingEngine.py:87,199,212; lib/optimUtils.py:528,531,659).
//
// Per source: the distance row lives directly in the output matrix (global memory,
// L2-resident while the CTA works on it).  A vertex whose tentative distance improves
// is appended to the NEAR queue (shared memory, spilling to global) when the new distance
// is below the current threshold theta, otherwise to the FAR pile (global); when NEAR
// drains, theta jumps to the bucket of the smallest valid FAR entry and the pile is split
// again.  Entries carry the distance they were pushed with, so stale ones (a better
// distance arrived later) are skipped.  Relaxations read the target first and issue an
// atomicMin on the row only when they improve it (non-negative floats order like their
// bit patterns); the adjacency of a vertex is one 32-byte ELL record of four
// (target, weight) slots, read with two 16-byte loads.
#include <algorithm>
#include <cstring>
#include <limits>
#include <vector>

#include "drs_common.cuh"

namespace {


template <typename T> struct Dist;
template <> struct Dist<float> {
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
  static __device__ __forceinline__ float weight(int bits) { return __int_as_float(bits); }
  static __device__ __forceinline__ int key(float d) { return __float_as_int(d); }
  static __device__ __forceinline__ float unkey(int k) { return __int_as_float(k); }
  // bucket boundary above dmin; clamped to the next float above dmin so that the smallest FAR entry always moves to
  // NEAR (rounding of floorf(dmin/delta)*delta, or dmin/delta >= 2^24, could otherwise leave theta <= dmin: no progress)
  static __device__ __forceinline__ float next_theta(float dmin, float delta) {
    return fmaxf((floorf(dmin / delta) + 1.0f) * delta, __int_as_float(__float_as_int(dmin) + 1));
  }
};
template <> struct Dist<int> {
  static __device__ __forceinline__ int inf() { return DRS_I32_INF; }
  static __device__ __forceinline__ int weight(int bits) { return (int)__int_as_float(bits); }
  static __device__ __forceinline__ int key(int d) { return d; }
  static __device__ __forceinline__ int unkey(int k) { return k; }
  static __device__ __forceinline__ int next_theta(int dmin, int delta) { return max((dmin / delta + 1) * delta, dmin + 1); }
};

#ifndef DRS_SSSP_MIN_THREADS_PER_SM
#define DRS_SSSP_MIN_THREADS_PER_SM 1152   // resident threads per SM the register allocation must allow (measured best)
#endif
struct __align__(8) Entry { int v; int key; };   // v = vertex | (source slot << 24); key = distance bits at push time
constexpr int SSSP_D = 4;            // arcs per entry handled in registers (road networks: degree <= 4 almost everywhere)
constexpr int SLOT_SHIFT = 24;
constexpr int VERTEX_MASK = (1 << SLOT_SHIFT) - 1;

#ifndef DRS_SSSP_NQ
#define DRS_SSSP_NQ 512
#endif
#ifndef DRS_SSSP_DEFAULT_SMEM
#define DRS_SSSP_DEFAULT_SMEM 0   // which variant drs_sssp_batch_dev picks when DRS_SSSP_IMPL is unset
#endif
#ifndef DRS_SSSP_DEFAULT_ARC
#define DRS_SSSP_DEFAULT_ARC 0    // 1: the lane-per-arc kernel with the row in L2 is the default
#endif
#ifndef DRS_SSSP_ARC_SHAPE
#define DRS_SSSP_ARC_SHAPE 12812  // CTA size * 100 + CTAs per SM
#endif
#ifndef DRS_SSSP_SMEM_DELTA_MULT
#define DRS_SSSP_SMEM_DELTA_MULT 24.0
#endif
constexpr int SSSP_NQ = DRS_SSSP_NQ;  // NEAR-queue entries held in shared memory per buffer; the rest spills to global (a large carve-out would cost the L1 that caches the adjacency)

// The 32-byte adjacency record of a vertex with ONE 256-bit load (LDG.E.ENL2.256, sm_100): a warp whose lanes hold
// 32 different vertices then costs the L1TEX pipe 32 wavefronts instead of the 64 of two 128-bit loads -- and both
// kernels below are bound by exactly those wavefronts (profiles/r02_sssp_step_profile.md).
__device__ __forceinline__ void ell_load(const int4* __restrict__ ell, int v, int4& r0, int4& r1) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w), "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w)
               : "l"(ell + 2 * (int64_t)v));
}

// One CTA advances SSSP_S sources in lock step (S = 1 by default): their wavefronts pass the same distance
// buckets at the same time, so one near queue / far pile / theta serves all of them.
//
// Dependent memory round trips per NEAR iteration: queue entry (shared memory) -> {stale check of the row,
// the vertex's 4-slot adjacency record} (both need only the vertex id) -> target distances -> atomicMins ->
// pushes (shared memory).  The adjacency is an ELL record of 4 (target, weight) slots per vertex (32 bytes, two
// 16-byte loads); slot 3 == (-2, a) says "more arcs follow in the CSR from index a".
template <typename T, int SSSP_THREADS, int SSSP_S>
__global__ void __launch_bounds__(SSSP_THREADS, DRS_SSSP_MIN_THREADS_PER_SM / SSSP_THREADS) sssp_batch_kernel(
    int n, const int32_t* __restrict__ out_ptr, const int2* __restrict__ arcs, const int4* __restrict__ ell, T* dist, int64_t ld,
    int64_t row_stride, const int32_t* __restrict__ sources, int nsrc, Entry* queues, int cap, T delta, int* overflow) {
  __shared__ int s_near, s_next, s_far, s_keep, s_min, s_group;
  __shared__ Entry s_q[2][SSSP_NQ];
  int* next_group = overflow + 1;   // dynamic work distribution: searches differ in length (corner vs centre sources)
  Entry* gq[2] = {queues + (int64_t)blockIdx.x * 4 * cap, queues + (int64_t)blockIdx.x * 4 * cap + cap};
  Entry* far = gq[1] + cap;
  Entry* far2 = far + cap;
  int sel = 0;                       // s_q[sel] / gq[sel] is the current NEAR queue
  const int tid = threadIdx.x;
  auto q_put = [&](int which, int pos, Entry e) {
    if (pos < SSSP_NQ) s_q[which][pos] = e;
    else if (pos - SSSP_NQ < cap) gq[which][pos - SSSP_NQ] = e;
    else *overflow = 1;
  };
  auto q_get = [&](int which, int pos) { return pos < SSSP_NQ ? s_q[which][pos] : gq[which][pos - SSSP_NQ]; };
  while (true) {
    if (tid == 0) s_group = atomicAdd(next_group, 1);
    __syncthreads();
    const int k0 = s_group * SSSP_S;
    if (k0 >= nsrc) break;
    const int ns = min(SSSP_S, nsrc - k0);
    T* rows = dist + (int64_t)k0 * row_stride;
    for (int sl = 0; sl < ns; ++sl)
      for (int64_t t = tid; t < ld; t += SSSP_THREADS) rows[(int64_t)sl * row_stride + t] = Dist<T>::inf();
    __syncthreads();
    if (tid < ns) {
      const int src = sources[k0 + tid];
      rows[(int64_t)tid * row_stride + src] = (T)0;
      s_q[sel][tid] = Entry{src | (tid << SLOT_SHIFT), Dist<T>::key((T)0)};
    }
    if (tid == 0) { s_near = ns; s_next = 0; s_far = 0; }
    T theta = delta;
    __syncthreads();
    while (true) {
      // ---- drain the NEAR queue (light phase of the current bucket)
      while (true) {
        const int nn = s_near;
        if (nn == 0) break;
        for (int e = tid; e < nn; e += SSSP_THREADS) {
          const Entry en = q_get(sel, e);
          const int v = en.v & VERTEX_MASK, slot = en.v >> SLOT_SHIFT;
          T* row = rows + (int64_t)slot * row_stride;
          const int kcur = Dist<T>::key(__ldcg(&row[v]));   // L2 read: the atomics live there
          int4 r0, r1;
          ell_load(ell, v, r0, r1);
          if (kcur != en.key) continue;                      // stale: a shorter distance arrived later
          const int2 arc[SSSP_D] = {make_int2(r0.x, r0.y), make_int2(r0.z, r0.w), make_int2(r1.x, r1.y), make_int2(r1.z, r1.w)};
          int nk[SSSP_D], old[SSSP_D];
          // plain L2 reads first: most arcs point back at settled vertices and need no atomic at all
#pragma unroll
          for (int j = 0; j < SSSP_D; ++j)
            if (arc[j].x >= 0) {
              nk[j] = Dist<T>::key(Dist<T>::unkey(en.key) + Dist<T>::weight(arc[j].y));
              old[j] = Dist<T>::key(__ldcg(&row[arc[j].x]));
            }
#pragma unroll
          for (int j = 0; j < SSSP_D; ++j)
            if (arc[j].x >= 0 && nk[j] < old[j]) old[j] = atomicMin(reinterpret_cast<int*>(row) + arc[j].x, nk[j]);
#pragma unroll
          for (int j = 0; j < SSSP_D; ++j)
            if (arc[j].x >= 0 && nk[j] < old[j]) {
              const Entry ne{arc[j].x | (slot << SLOT_SHIFT), nk[j]};
              if (Dist<T>::unkey(nk[j]) < theta) {
                q_put(sel ^ 1, atomicAdd(&s_next, 1), ne);
              } else {
                const int pos = atomicAdd(&s_far, 1);
                if (pos < cap) far[pos] = ne; else *overflow = 1;
              }
            }
          if (arc[SSSP_D - 1].x == -2) {                      // high-degree vertex: the remaining arcs live in the CSR
            const int a1 = out_ptr[v + 1];
            for (int a = arc[SSSP_D - 1].y; a < a1; ++a) {
              const int2 ar = arcs[a];
              const int k2 = Dist<T>::key(Dist<T>::unkey(en.key) + Dist<T>::weight(ar.y));
              int o2 = Dist<T>::key(__ldcg(&row[ar.x]));
              if (k2 < o2) o2 = atomicMin(reinterpret_cast<int*>(row) + ar.x, k2);
              if (k2 < o2) {
                const Entry ne{ar.x | (slot << SLOT_SHIFT), k2};
                if (Dist<T>::unkey(k2) < theta) {
                  q_put(sel ^ 1, atomicAdd(&s_next, 1), ne);
                } else {
                  const int pos = atomicAdd(&s_far, 1);
                  if (pos < cap) far[pos] = ne; else *overflow = 1;
                }
              }
            }
          }
        }
        __syncthreads();
        if (tid == 0) { s_near = min(s_next, SSSP_NQ + cap); s_next = 0; }
        sel ^= 1;
        __syncthreads();
      }
      // ---- NEAR is empty: advance theta to the bucket of the smallest valid FAR entry (any source)
      const int nf = min(s_far, cap);
      if (nf == 0) break;
      if (tid == 0) { s_min = 0x7fffffff; s_keep = 0; }
      __syncthreads();
      int lmin = 0x7fffffff;
      for (int e = tid; e < nf; e += SSSP_THREADS) {
        const Entry en = far[e];
        const T* row = rows + (int64_t)(en.v >> SLOT_SHIFT) * row_stride;
        if (Dist<T>::key(__ldcg(&row[en.v & VERTEX_MASK])) == en.key) lmin = min(lmin, en.key);
      }
      for (int off = 16; off > 0; off >>= 1) lmin = min(lmin, __shfl_down_sync(0xffffffffu, lmin, off));
      if ((tid & 31) == 0 && lmin != 0x7fffffff) atomicMin(&s_min, lmin);
      __syncthreads();
      const int kmin = s_min;
      if (kmin == 0x7fffffff) break;   // only stale entries left
      theta = Dist<T>::next_theta(Dist<T>::unkey(kmin), delta);
      for (int e = tid; e < nf; e += SSSP_THREADS) {
        const Entry en = far[e];
        const T* row = rows + (int64_t)(en.v >> SLOT_SHIFT) * row_stride;
        if (Dist<T>::key(__ldcg(&row[en.v & VERTEX_MASK])) != en.key) continue;
        if (Dist<T>::unkey(en.key) < theta) q_put(sel, atomicAdd(&s_near, 1), en);
        else far2[atomicAdd(&s_keep, 1)] = en;
      }
      __syncthreads();
      if (tid == 0) { s_far = s_keep; }
      Entry* tmp = far; far = far2; far2 = tmp;
      __syncthreads();
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------------------
// Shared-memory-resident variant: the distance row of the source being searched lives in the CTA's shared
// memory (4*ld bytes: 200 704 B for the 50k-node map, which fits the 227 KB a B200 CTA can own), so the stale
// check, the target reads and the atomicMins of every relaxation are shared-memory operations (~30 cycles)
// instead of L2 round trips (~250 cycles each, three dependent ones per wavefront step in the kernel above),
// and the finished row leaves the SM once, as one coalesced 16-byte-per-lane store (DRAM sees the matrix
// exactly once).  What stays in global memory: the ELL adjacency record of each visited vertex (read-only,
// L2/L1), the FAR pile, and the overflow of the NEAR queues.
//
// One barrier per wavefront step: the NEAR queues alternate between two buffers, their fill counts rotate
// through three words (step r reads count r%3, appends to (r+1)%3 and clears (r+2)%3), and appends are
// warp-aggregated (one shared-memory atomicAdd per warp and arc slot).
#ifdef DRS_SSSP_PROFILE
// cycle totals of thread 0 of every CTA: [0] steps, [1] whole wavefront steps, [2] queue entries of those steps, [6] far phases,
// [7] row write-out, [8] whole kernel
__device__ unsigned long long g_sssp_prof[16];
#define SSSP_PROF_T(x) const long long x = clock64()
#define SSSP_PROF_ADD(i, v) do { if (threadIdx.x == 0) prof[i] += (unsigned long long)(v); } while (0)
#else
#define SSSP_PROF_T(x)
#define SSSP_PROF_ADD(i, v)
#endif

// FAR = false compiles the one-bucket form (Bellman-Ford-Moore wavefronts: every improved vertex goes to the next
// NEAR queue, no FAR pile, no theta), which is what an "infinite" bucket width selects.
// SROW = false is the same search with the row left in the output matrix (global memory, L2), i.e. the structure of
// sssp_batch_kernel rebuilt on this kernel's leaner step -- one lane per (entry, arc), one barrier per step, one counter
// atomic per warp -- so that MORE searches fit an SM (MINB CTAs of THREADS): the measured lever of this workload is the
// number of independent searches in flight per SM (profiles/r02_sssp_step_profile.md).
template <typename T, int THREADS, bool FAR, bool SROW, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) sssp_smem_kernel(
    int n, const int32_t* __restrict__ out_ptr, const int2* __restrict__ arcs, const int4* __restrict__ ell, T* dist, int64_t ld,
    int64_t row_stride, const int32_t* __restrict__ sources, int nsrc, Entry* queues, int cap, T delta, int* overflow, int qn) {
  extern __shared__ __align__(16) unsigned char sssp_smem[];
  T* row = reinterpret_cast<T*>(sssp_smem);
  int* irow = reinterpret_cast<int*>(sssp_smem);
  Entry* s_q = reinterpret_cast<Entry*>(sssp_smem + (SROW ? sizeof(T) * ld : 0));   // two NEAR buffers of qn entries
  auto rd = [&](int x) { return SROW ? irow[x] : __ldcg(irow + x); };   // L2 read when the row is global: the atomics live there
  __shared__ int s_cnt[3], s_far, s_keep, s_min, s_group;
  int* next_group = overflow + 1;
  Entry* gq = queues + (int64_t)blockIdx.x * 4 * cap;                  // their overflow: two buffers of cap entries
  Entry* far = gq + 2 * (int64_t)cap;
  Entry* far2 = far + cap;
  const int tid = threadIdx.x, lane = tid & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  auto q_put = [&](int which, int pos, Entry e) {
    if (pos < qn) s_q[which * qn + pos] = e;
    else if (pos - qn < cap) gq[(int64_t)which * cap + pos - qn] = e;
    else *overflow = 1;
  };
  auto q_get = [&](int which, int pos) { return pos < qn ? s_q[which * qn + pos] : gq[(int64_t)which * cap + pos - qn]; };
  const T inf = Dist<T>::inf();
#ifdef DRS_SSSP_PROFILE
  unsigned long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long prof_t0 = clock64();
#endif
  // the row starts as +inf everywhere; after every search it is written out and reset in the same sweep
  if (SROW)
    for (int64_t t = tid; t < ld; t += THREADS) row[t] = inf;
  while (true) {
    __syncthreads();
    if (tid == 0) s_group = atomicAdd(next_group, 1);
    __syncthreads();
    const int k = s_group;
    if (k >= nsrc) break;
    if (!SROW) {
      row = dist + (int64_t)k * row_stride;
      irow = reinterpret_cast<int*>(row);
      for (int64_t t = tid; t < ld; t += THREADS) row[t] = inf;
      __syncthreads();
    }
    if (tid == 0) {
      const int src = sources[k];
      row[src] = (T)0;
      s_q[0] = Entry{src, Dist<T>::key((T)0)};
      s_cnt[0] = 1; s_cnt[1] = 0; s_cnt[2] = 0; s_far = 0;
    }
    T theta = delta;
    int buf = 0, c_rd = 0;             // queue buffer being read; index of its count word
    while (true) {
      // ---- wavefront steps of the current bucket
      while (true) {
        SSSP_PROF_T(pa);
        __syncthreads();
        const int nn = min(s_cnt[c_rd], qn + cap);
        const int c_wr = c_rd == 2 ? 0 : c_rd + 1, c_clr = c_wr == 2 ? 0 : c_wr + 1;
        if (tid == 0) s_cnt[c_clr] = 0;
        if (nn == 0) break;
        // One lane per (queue entry, arc slot): the wavefront step is one warp's dependent chain (entry -> adjacency,
        // stale check -> target -> atomicMin -> append), so the chain is kept short (one arc, ~50 instructions) and
        // spread over all warps instead of four arcs unrolled in a quarter of them (ncu: 200 instructions per warp and
        // step, IPC 1.2, two thirds of the warps idle at the barrier -- profiles/r02_sssp_step_profile.md).
        const int2* ell2 = reinterpret_cast<const int2*>(ell);
        for (int base = tid - lane; base < 4 * nn; base += THREADS) {
          SSSP_PROF_T(p0);
          const int e = (base + lane) >> 2, j = lane & 3;
          Entry en{0, 0};
          if (e < nn) en = q_get(buf, e);
          int2 arc = make_int2(-1, 0);
          bool live = false;
          if (e < nn) {
            arc = __ldg(&ell2[4 * (int64_t)en.v + j]);
            live = rd(en.v) == en.key;                   // stale otherwise: a shorter distance arrived later
          }
#ifdef DRS_SSSP_PROFILE
          if (arc.x == -12345 && live) prof[9] += 1;   // consume the loads before the time stamp
#endif
          SSSP_PROF_T(p1);
          SSSP_PROF_ADD(3, p1 - p0);
          const int nk = Dist<T>::key(Dist<T>::unkey(en.key) + Dist<T>::weight(arc.y));
          bool push = live && arc.x >= 0 && nk < rd(max(arc.x, 0));
#ifdef DRS_SSSP_PROFILE
          if (push && nk == -12345) prof[9] += 1;
#endif
          SSSP_PROF_T(p2);
          SSSP_PROF_ADD(4, p2 - p1);
          if (push) push = nk < atomicMin(irow + arc.x, nk);
#ifdef DRS_SSSP_PROFILE
          if (push && nk == -12345) prof[9] += 1;
#endif
          SSSP_PROF_T(p3);
          SSSP_PROF_ADD(5, p3 - p2);
          const bool nearp = push && (!FAR || Dist<T>::unkey(nk) < theta);
          const unsigned mn = __ballot_sync(0xffffffffu, nearp);
          if (mn) {
            int b0 = 0;
            if (lane == 0) b0 = atomicAdd(&s_cnt[c_wr], __popc(mn));
            b0 = __shfl_sync(0xffffffffu, b0, 0);
            if (nearp) q_put(buf ^ 1, b0 + __popc(mn & lt_mask), Entry{arc.x, nk});
          }
          if (FAR) {
            const unsigned mf = __ballot_sync(0xffffffffu, push && !nearp);
            if (mf) {
              int b0 = 0;
              if (lane == 0) b0 = atomicAdd(&s_far, __popc(mf));
              b0 = __shfl_sync(0xffffffffu, b0, 0);
              if (push && !nearp) {
                const int pos = b0 + __popc(mf & lt_mask);
                if (pos < cap) far[pos] = Entry{arc.x, nk}; else *overflow = 1;
              }
            }
          }
          SSSP_PROF_T(p4);
          SSSP_PROF_ADD(9, p4 - p3);
          if (live && j == SSSP_D - 1 && arc.x == -2) {      // high-degree vertex: the remaining arcs live in the CSR
            const int a1 = out_ptr[en.v + 1];
            for (int a = arc.y; a < a1; ++a) {
              const int2 ar = arcs[a];
              const int k2 = Dist<T>::key(Dist<T>::unkey(en.key) + Dist<T>::weight(ar.y));
              if (k2 < rd(ar.x) && k2 < atomicMin(irow + ar.x, k2)) {
                const Entry ne{ar.x, k2};
                if (!FAR || Dist<T>::unkey(k2) < theta) {
                  q_put(buf ^ 1, atomicAdd(&s_cnt[c_wr], 1), ne);
                } else {
                  const int pos = atomicAdd(&s_far, 1);
                  if (pos < cap) far[pos] = ne; else *overflow = 1;
                }
              }
            }
          }
        }
        buf ^= 1;
        c_rd = c_wr;
        SSSP_PROF_T(pb);
        SSSP_PROF_ADD(0, 1);
        SSSP_PROF_ADD(1, pb - pa);
        SSSP_PROF_ADD(2, nn);
      }
      if (!FAR) break;
      // ---- NEAR is empty (its count word c_rd reads 0): advance theta to the bucket of the smallest valid FAR entry
      SSSP_PROF_T(pf0);
      const int nf = min(s_far, cap);
      if (nf == 0) break;
      if (tid == 0) { s_min = 0x7fffffff; s_keep = 0; }
      __syncthreads();
      int lmin = 0x7fffffff;
      for (int e = tid; e < nf; e += THREADS) {
        const Entry en = far[e];
        if (rd(en.v) == en.key) lmin = min(lmin, en.key);
      }
      for (int off = 16; off > 0; off >>= 1) lmin = min(lmin, __shfl_down_sync(0xffffffffu, lmin, off));
      if (lane == 0 && lmin != 0x7fffffff) atomicMin(&s_min, lmin);
      __syncthreads();
      const int kmin = s_min;
      if (kmin == 0x7fffffff) break;   // only stale entries left
      theta = Dist<T>::next_theta(Dist<T>::unkey(kmin), delta);
      for (int e = tid; e < nf; e += THREADS) {
        const Entry en = far[e];
        if (rd(en.v) != en.key) continue;
        if (Dist<T>::unkey(en.key) < theta) q_put(buf, atomicAdd(&s_cnt[c_rd], 1), en);
        else far2[atomicAdd(&s_keep, 1)] = en;
      }
      __syncthreads();
      if (tid == 0) s_far = s_keep;
      Entry* tmp = far; far = far2; far2 = tmp;
      SSSP_PROF_T(pf1);
      SSSP_PROF_ADD(6, pf1 - pf0);
    }
    // ---- the row is final: one coalesced sweep writes it out and resets the shared copy
    SSSP_PROF_T(pw0);
    __syncthreads();
    if (SROW) {
      T* out = dist + (int64_t)k * row_stride;
      const int4 inf4 = make_int4(Dist<T>::key(inf), Dist<T>::key(inf), Dist<T>::key(inf), Dist<T>::key(inf));
      int4* r4 = reinterpret_cast<int4*>(row);
      int4* o4 = reinterpret_cast<int4*>(out);
      for (int t = tid; t < (int)(ld / 4); t += THREADS) {
        __stcs(&o4[t], r4[t]);       // streaming store: the row is not read again by this kernel
        r4[t] = inf4;
      }
    }
    SSSP_PROF_T(pw1);
    SSSP_PROF_ADD(7, pw1 - pw0);
  }
#ifdef DRS_SSSP_PROFILE
  if (tid == 0) {
    prof[8] = (unsigned long long)(clock64() - prof_t0);
    for (int i = 0; i < 10; ++i) atomicAdd(&g_sssp_prof[i], prof[i]);
  }
#endif
}

// rows >= n of the padded matrix: 0 on the diagonal, inf elsewhere (what drs_fw_init writes)
template <typename T>
__global__ void pad_rows_kernel(T* local, int64_t ld, int row0, int r_begin, int r_end) {
  const int64_t total = (int64_t)(r_end - r_begin) * ld;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = r_begin + (int)(idx / ld), j = (int)(idx % ld);
    local[(int64_t)i * ld + j] = (row0 + i == j) ? (T)0 : Dist<T>::inf();
  }
}

// ELL adjacency: 4 (target, weight) slots per vertex; unused slots (-1, 0); slot 3 = (-2, a) when more arcs follow at CSR index a
__global__ void build_ell_kernel(int n, const int32_t* out_ptr, const int32_t* col, const float* w, int2* ell) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int a0 = out_ptr[v], deg = out_ptr[v + 1] - a0;
  for (int j = 0; j < SSSP_D; ++j) {
    int2 rec = make_int2(-1, 0);
    if (j < deg) rec = make_int2(col[a0 + j], __float_as_int(w[a0 + j]));
    if (j == SSSP_D - 1 && deg > SSSP_D) rec = make_int2(-2, a0 + SSSP_D - 1);
    ell[(int64_t)v * SSSP_D + j] = rec;
  }
}

__global__ void pack_arcs_kernel(int narcs, const int32_t* col, const float* w, int2* arcs) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < narcs) arcs[a] = make_int2(col[a], __float_as_int(w[a]));
}

int sssp_env(const char* name, int dflt, int a, int b, int c) {
  const char* e = getenv(name);
  const int t = e ? atoi(e) : dflt;
  return (t == a || t == b || t == c || t == 8) ? t : dflt;
}
// tuning knobs, defaults chosen by measurement on B200 (tools/gpu_check_sssp2.py)
int sssp_threads() { static int t = sssp_env("DRS_SSSP_THREADS", 128, 64, 128, 256); return t; }
int sssp_slots() { static int u = sssp_env("DRS_SSSP_S", 1, 1, 2, 4) ; return u; }   // sources per CTA (also 8 below)

template <typename T, int TH, int U>
int sssp_launch(bool query, int* per_sm, int grid, cudaStream_t st, int n, const int32_t* out_ptr, const int2* arcs, const int4* ell,
                T* dist, int64_t ld, int64_t row_stride, const int32_t* sources, int nsrc, Entry* queues, int cap, T delta, int* overflow) {
  if (query) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, sssp_batch_kernel<T, TH, U>, TH, 0);
    return DRS_OK;
  }
  sssp_batch_kernel<T, TH, U><<<grid, TH, 0, st>>>(n, out_ptr, arcs, ell, dist, ld, row_stride, sources, nsrc, queues, cap, delta, overflow);
  return DRS_OK;
}

template <typename T, typename... A>
int sssp_dispatch(A... args) {
  const int th = sssp_threads(), u = sssp_slots();
#define DRS_SSSP_CASE(TH, U) if (th == TH && u == U) return sssp_launch<T, TH, U>(args...);
  DRS_SSSP_CASE(64, 1) DRS_SSSP_CASE(64, 2) DRS_SSSP_CASE(64, 4) DRS_SSSP_CASE(64, 8)
  DRS_SSSP_CASE(128, 1) DRS_SSSP_CASE(128, 2) DRS_SSSP_CASE(128, 4) DRS_SSSP_CASE(128, 8)
  DRS_SSSP_CASE(256, 1) DRS_SSSP_CASE(256, 2) DRS_SSSP_CASE(256, 4) DRS_SSSP_CASE(256, 8)
#undef DRS_SSSP_CASE
  return DRS_ERR_ARG;
}

template <typename T>
int sssp_grid(int sms, int nsrc) {
  int per_sm = 1;
  sssp_dispatch<T>(true, &per_sm, 0, (cudaStream_t) nullptr, 0, (const int32_t*)nullptr, (const int2*)nullptr, (const int4*)nullptr, (T*)nullptr,
                   (int64_t)0, (int64_t)0, (const int32_t*)nullptr, 0, (Entry*)nullptr, 0, (T)0, (int*)nullptr);
  const int groups = (nsrc + sssp_slots() - 1) / sssp_slots();
  // CTAs in flight per SM: every CTA keeps one distance row (4n bytes) hot, and the rows of all resident CTAs
  // should stay in the 126 MB L2 (the kernel runs at >70 % of L2 throughput); DRS_SSSP_CTAS_PER_SM overrides
  static int cap_per_sm = -1;
  if (cap_per_sm < 0) { const char* e = getenv("DRS_SSSP_CTAS_PER_SM"); cap_per_sm = e ? atoi(e) : 0; }
  if (cap_per_sm > 0) per_sm = std::min(per_sm, cap_per_sm);
  return std::max(1, std::min(groups, sms * std::max(per_sm, 1)));
}

template <typename T>
int sssp_run(const drs_graph* g, int nsrc, const int32_t* sources_dev, T* dist, int64_t ld, int64_t row_stride, double delta,
             cudaStream_t st, const int2* arcs, const int4* ell, Entry* queues, int cap, int grid, int* overflow) {
  T d = (T)delta;
  if (!(d > (T)0)) d = (T)1;
  int dummy = 0;
  sssp_dispatch<T>(false, &dummy, grid, st, g->n, (const int32_t*)g->d_out_ptr, arcs, ell, dist, ld, row_stride, sources_dev, nsrc, queues,
                   cap, d, overflow);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}


// ---- launch of the shared-memory-resident variant
int sssp_env_int(const char* name, int dflt) {   // read at every call: the probes change these between builds
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

constexpr int SSSP_SMEM_LIMIT = 227 * 1024 - 256;   // dynamic shared memory a CTA may own on sm_100, minus the kernel's static words

// NEAR-queue entries per buffer that fit beside the row (0: the row does not fit -> L2 variant)
int sssp_smem_qn(int64_t ld, size_t elem) {
  const int64_t avail = (int64_t)SSSP_SMEM_LIMIT - (int64_t)elem * ld;
  if (avail < (int64_t)(2 * 64 * sizeof(Entry))) return 0;
  int qn = (int)std::min<int64_t>(avail / (2 * sizeof(Entry)), 2048);
  const int want = sssp_env_int("DRS_SSSP_QN", 0);
  if (want >= 64) qn = std::min(qn, want);
  return qn / 32 * 32;
}

int sssp_smem_threads(int n) {
  const int t = sssp_env_int("DRS_SSSP_SMEM_THREADS", 0);
  if (t == 128 || t == 256 || t == 512 || t == 1024) return t;
  return n <= 2048 ? 128 : (n <= 16384 ? 256 : 1024);
}

template <typename T, int TH, bool FAR, bool SROW, int MINB>
int sssp_smem_go(int sms, cudaStream_t st, int n, const int32_t* out_ptr, const int2* arcs, const int4* ell, T* dist, int64_t ld,
                 int64_t row_stride, const int32_t* sources, int nsrc, Entry* queues, int cap, T delta, int* overflow, int qn, int* grid_out) {
  const size_t smem = (SROW ? sizeof(T) * (size_t)ld : 0) + 2 * (size_t)qn * sizeof(Entry);
  auto kern = sssp_smem_kernel<T, TH, FAR, SROW, MINB>;
  DRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  DRS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TH, smem));
  const int grid = std::max(1, std::min(nsrc, sms * std::max(per_sm, 1)));
  if (grid_out) { *grid_out = grid; return DRS_OK; }
  kern<<<grid, TH, smem, st>>>(n, out_ptr, arcs, ell, dist, ld, row_stride, sources, nsrc, queues, cap, delta, overflow, qn);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}

template <typename T, typename... A>
int sssp_smem_dispatch(int threads, bool use_far, A... args) {
#define DRS_SMEM_CASE(TH) case TH: return use_far ? sssp_smem_go<T, TH, true, true, 1>(args...) : sssp_smem_go<T, TH, false, true, 1>(args...);
  switch (threads) {
    DRS_SMEM_CASE(128) DRS_SMEM_CASE(256) DRS_SMEM_CASE(512)
    default: return use_far ? sssp_smem_go<T, 1024, true, true, 1>(args...) : sssp_smem_go<T, 1024, false, true, 1>(args...);
  }
#undef DRS_SMEM_CASE
}

// the row-in-L2 form of the same kernel: `shape` = CTA size * 100 + CTAs per SM the register allocation must allow
template <typename T, typename... A>
int sssp_arc_dispatch(int shape, A... args) {
  switch (shape) {
    case 6432: return sssp_smem_go<T, 64, true, false, 32>(args...);
    case 6424: return sssp_smem_go<T, 64, true, false, 24>(args...);
    case 12816: return sssp_smem_go<T, 128, true, false, 16>(args...);
    case 12812: return sssp_smem_go<T, 128, true, false, 12>(args...);
    case 12809: return sssp_smem_go<T, 128, true, false, 9>(args...);
    case 25608: return sssp_smem_go<T, 256, true, false, 8>(args...);
    default: return sssp_smem_go<T, 128, true, false, 12>(args...);
  }
}

}  // namespace

// Distances from `nsrc` sources into rows 0..nsrc-1 of dist_dev (row k <- sources[k]).
extern "C" int drs_sssp_batch_dev(const drs_graph* g, int32_t mode, int32_t nsrc, const int32_t* sources_host, void* dist_dev,
                                  int64_t ld, double delta, void* stream) {
  DRS_REQUIRE(g && dist_dev && (nsrc == 0 || sources_host), DRS_ERR_ARG, "drs_sssp_batch_dev: null argument");
  DRS_REQUIRE(ld >= g->n && nsrc >= 0, DRS_ERR_ARG, "drs_sssp_batch_dev: ld (%lld) must be >= n (%d)", (long long)ld, g->n);
  DRS_REQUIRE(g->n <= VERTEX_MASK, DRS_ERR_ARG, "drs_sssp_batch_dev: at most %d vertices (queue entries pack the source slot above bit 24)", VERTEX_MASK);
  { int rcm = drs_check_mode(g, mode, "drs_sssp_batch_dev"); if (rcm) return rcm; }
  for (int k = 0; k < nsrc; ++k)
    DRS_REQUIRE(sources_host[k] >= 0 && sources_host[k] < g->n, DRS_ERR_RANGE, "drs_sssp_batch_dev: source %d out of range", sources_host[k]);
  if (nsrc == 0) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  cudaStream_t st = (cudaStream_t)stream;
  drs_graph* gm = const_cast<drs_graph*>(g);   // scratch buffers are cached on the handle
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g->device);
  // variant: the row in shared memory when it fits (DRS_SSSP_IMPL=l2 forces the L2-resident kernel, =smem insists)
  const char* impl_env = getenv("DRS_SSSP_IMPL");
  const int qn = sssp_smem_qn(ld, sizeof(float));
  // default by measurement (profiles/r02_sssp_step_profile.md): DRS_SSSP_DEFAULT_SMEM
  const bool want_smem = impl_env ? !strcmp(impl_env, "smem") : (DRS_SSSP_DEFAULT_SMEM != 0);
  const bool want_arc = impl_env ? !strcmp(impl_env, "arc") : (DRS_SSSP_DEFAULT_ARC != 0);
  const int arc_shape = sssp_env_int("DRS_SSSP_ARC_SHAPE", DRS_SSSP_ARC_SHAPE);
  const int arc_qn = 256;
  bool use_smem = want_smem && qn > 0 && ld % 4 == 0 && ((uintptr_t)dist_dev & 15) == 0;
  DRS_REQUIRE(use_smem || !(impl_env && !strcmp(impl_env, "smem")), DRS_ERR_ARG,
              "drs_sssp_batch_dev: DRS_SSSP_IMPL=smem but a row of %lld elements does not fit in shared memory", (long long)ld);
  const int smem_threads = sssp_smem_threads(g->n);
  if (delta <= 0) {
    // default bucket width, in mean arc weights (measured on grid cities): 6 for the L2-resident rows, where every
    // re-relaxation is an L2 round trip; wider for the shared-memory rows, where re-relaxations are cheap and the
    // number of wavefront steps (barriers) is what costs.  DRS_SSSP_DELTA_MULT overrides (0 = one bucket: Bellman-Ford-Moore)
    double s = 0;
    for (double w : g->h_tt) s += w;
    const char* e = getenv("DRS_SSSP_DELTA_MULT");
    const double mult = e ? atof(e) : (use_smem ? DRS_SSSP_SMEM_DELTA_MULT : 6.0);
    delta = mult > 0 ? (g->m ? mult * s / g->m : 1.0) : 3.0e38;
  }
  const bool use_far = delta < 1.0e30;   // one bucket: the Bellman-Ford-Moore form without a FAR pile
  const float big_f = 3.0e38f;
  int grid = 0;
  if (use_smem) {
    int rc0 = mode == DRS_MODE_F32
                  ? sssp_smem_dispatch<float>(smem_threads, true, sms, (cudaStream_t) nullptr, 0, (const int32_t*)nullptr, (const int2*)nullptr,
                                              (const int4*)nullptr, (float*)nullptr, ld, ld, (const int32_t*)nullptr, nsrc, (Entry*)nullptr, 0,
                                              0.0f, (int*)nullptr, qn, &grid)
                  : sssp_smem_dispatch<int>(smem_threads, true, sms, (cudaStream_t) nullptr, 0, (const int32_t*)nullptr, (const int2*)nullptr,
                                            (const int4*)nullptr, (int*)nullptr, ld, ld, (const int32_t*)nullptr, nsrc, (Entry*)nullptr, 0, 0,
                                            (int*)nullptr, qn, &grid);
    if (rc0) return rc0;
  } else if (want_arc) {
    int rc0 = mode == DRS_MODE_F32
                  ? sssp_arc_dispatch<float>(arc_shape, sms, (cudaStream_t) nullptr, 0, (const int32_t*)nullptr, (const int2*)nullptr,
                                             (const int4*)nullptr, (float*)nullptr, ld, ld, (const int32_t*)nullptr, nsrc, (Entry*)nullptr, 0,
                                             0.0f, (int*)nullptr, arc_qn, &grid)
                  : sssp_arc_dispatch<int>(arc_shape, sms, (cudaStream_t) nullptr, 0, (const int32_t*)nullptr, (const int2*)nullptr,
                                           (const int4*)nullptr, (int*)nullptr, ld, ld, (const int32_t*)nullptr, nsrc, (Entry*)nullptr, 0, 0,
                                           (int*)nullptr, arc_qn, &grid);
    if (rc0) return rc0;
  } else {
    grid = mode == DRS_MODE_F32 ? sssp_grid<float>(sms, nsrc) : sssp_grid<int>(sms, nsrc);
  }
  // packed arcs (narcs records) followed by the ELL records (4 per vertex)
  if (!gm->sssp_arcs) drs_count_alloc();
  if (!gm->sssp_arcs) DRS_CUDA(cudaMalloc(&gm->sssp_arcs, sizeof(int2) * ((size_t)std::max(g->narcs, 1) + 4 + (size_t)SSSP_D * g->n)));
  int2* d_ell = (int2*)gm->sssp_arcs + ((std::max(g->narcs, 1) + 3) & ~3);   // 32-byte aligned: records are read with one 256-bit load
  if (!gm->sssp_arcs_valid) {
    if (g->narcs) pack_arcs_kernel<<<(g->narcs + 255) / 256, 256, 0, st>>>(g->narcs, g->d_out_col, g->d_out_w, (int2*)gm->sssp_arcs);
    build_ell_kernel<<<(g->n + 255) / 256, 256, 0, st>>>(g->n, g->d_out_ptr, g->d_out_col, g->d_out_w, d_ell);
    drs_count_launch(2);
    DRS_CUDA(cudaGetLastError());
    gm->sssp_arcs_valid = true;
  }
  if (gm->sssp_sources_cap < nsrc) {
    if (gm->sssp_sources) cudaFree(gm->sssp_sources);
    gm->sssp_sources = nullptr; gm->sssp_sources_cap = 0;
    drs_count_alloc();
    DRS_CUDA(cudaMalloc((void**)&gm->sssp_sources, sizeof(int32_t) * (size_t)nsrc + 2 * sizeof(int)));
    gm->sssp_sources_cap = nsrc;
  }
  int32_t* d_sources = gm->sssp_sources;
  int* d_overflow = reinterpret_cast<int*>(d_sources + gm->sssp_sources_cap);
  // queue capacity per CTA: frontiers of road networks are a few thousand entries; overflow is detected and retried
  // (an eighth of the vertices: the FAR pile of a grid city peaks at a few thousand entries; 4 * cap * 8 B per CTA, ~1 300 CTAs)
  int cap = std::max(4096, std::min(g->n / 8 + 1, 65536)) * sssp_slots();
  int rc = DRS_OK;
  for (int attempt = 0; attempt < 3 && rc == DRS_OK; ++attempt) {
    const size_t need = sizeof(Entry) * 4 * (size_t)cap * grid;
    if (gm->sssp_queue_bytes < need) {
      if (gm->sssp_queues) cudaFree(gm->sssp_queues);
      gm->sssp_queues = nullptr; gm->sssp_queue_bytes = 0;
      drs_count_alloc();
      if (cudaMalloc(&gm->sssp_queues, need) != cudaSuccess) {
        DRS_REQUIRE(false, DRS_ERR_NOMEM, "drs_sssp_batch_dev: cannot allocate %zu bytes of work queues", need);
      }
      gm->sssp_queue_bytes = need;
    }
    cudaMemsetAsync(d_overflow, 0, 2 * sizeof(int), st);   // overflow flag + work counter
    cudaMemcpyAsync(d_sources, sources_host, sizeof(int32_t) * nsrc, cudaMemcpyHostToDevice, st);
    if (use_smem && mode == DRS_MODE_F32)
      rc = sssp_smem_dispatch<float>(smem_threads, use_far, sms, st, g->n, (const int32_t*)g->d_out_ptr, (const int2*)gm->sssp_arcs, (const int4*)d_ell,
                                     (float*)dist_dev, ld, ld, (const int32_t*)d_sources, nsrc, (Entry*)gm->sssp_queues, cap,
                                     (float)std::min<double>(delta, big_f), d_overflow, qn, (int*)nullptr);
    else if (use_smem)
      rc = sssp_smem_dispatch<int>(smem_threads, use_far, sms, st, g->n, (const int32_t*)g->d_out_ptr, (const int2*)gm->sssp_arcs, (const int4*)d_ell,
                                   (int*)dist_dev, ld, ld, (const int32_t*)d_sources, nsrc, (Entry*)gm->sssp_queues, cap,
                                   (int)std::max(1.0, std::min<double>(delta, (double)DRS_I32_INF)), d_overflow, qn, (int*)nullptr);
    else if (want_arc && mode == DRS_MODE_F32)
      rc = sssp_arc_dispatch<float>(arc_shape, sms, st, g->n, (const int32_t*)g->d_out_ptr, (const int2*)gm->sssp_arcs, (const int4*)d_ell,
                                    (float*)dist_dev, ld, ld, (const int32_t*)d_sources, nsrc, (Entry*)gm->sssp_queues, cap,
                                    (float)std::min<double>(delta, big_f), d_overflow, arc_qn, (int*)nullptr);
    else if (want_arc)
      rc = sssp_arc_dispatch<int>(arc_shape, sms, st, g->n, (const int32_t*)g->d_out_ptr, (const int2*)gm->sssp_arcs, (const int4*)d_ell,
                                  (int*)dist_dev, ld, ld, (const int32_t*)d_sources, nsrc, (Entry*)gm->sssp_queues, cap,
                                  (int)std::max(1.0, std::min<double>(delta, (double)DRS_I32_INF)), d_overflow, arc_qn, (int*)nullptr);
    else if (mode == DRS_MODE_F32)
      rc = sssp_run<float>(g, nsrc, d_sources, (float*)dist_dev, ld, ld, delta, st, (const int2*)gm->sssp_arcs, (const int4*)d_ell,
                           (Entry*)gm->sssp_queues, cap, grid, d_overflow);
    else
      rc = sssp_run<int>(g, nsrc, d_sources, (int*)dist_dev, ld, ld, delta, st, (const int2*)gm->sssp_arcs, (const int4*)d_ell,
                         (Entry*)gm->sssp_queues, cap, grid, d_overflow);
    if (rc) break;
    int ovf = 0;
    cudaMemcpyAsync(&ovf, d_overflow, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { drs_set_error("drs_sssp_batch_dev: kernel failed: %s", cudaGetErrorString(cudaGetLastError())); rc = DRS_ERR_CUDA; break; }
#ifdef DRS_SSSP_PROFILE
    if (use_smem) {
      unsigned long long pr[16], zero[16] = {0};
      cudaMemcpyFromSymbol(pr, g_sssp_prof, sizeof(pr));
      cudaMemcpyToSymbol(g_sssp_prof, zero, sizeof(zero));
      const double st_ = (double)std::max<unsigned long long>(pr[0], 1);
      fprintf(stderr, "[sssp prof] grid %d steps/source %.1f entries/step %.1f cycles/step %.0f (thread 0, all its passes: load %.0f target %.0f atomic %.0f append %.0f) | far/source %.0f writeout/source %.0f | kernel cycles/CTA %.0f\n",
              grid, st_ / nsrc, pr[2] / st_, pr[1] / st_, pr[3] / st_, pr[4] / st_, pr[5] / st_, pr[9] / st_, (double)pr[6] / nsrc, (double)pr[7] / nsrc, (double)pr[8] / grid);
    }
#endif
    if (!ovf) return DRS_OK;
    cap *= 4;   // queue overflow: retry with larger queues
  }
  if (rc) return rc;
  DRS_REQUIRE(false, DRS_ERR_NOMEM, "drs_sssp_batch_dev: work queues overflowed three times");
}

// Host-output variant (BASELINE config C5): dist_out_host is nsrc x n doubles (inf = unreachable).
extern "C" int drs_sssp_batch(const drs_graph* g, int32_t mode, int32_t nsrc, const int32_t* sources_host, double* dist_out_host) {
  DRS_REQUIRE(g && (nsrc == 0 || (sources_host && dist_out_host)), DRS_ERR_ARG, "drs_sssp_batch: null argument");
  if (nsrc == 0) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  const int64_t ld = g->npad;
  void* d = nullptr;
  drs_count_alloc();
  DRS_CUDA(cudaMalloc(&d, sizeof(float) * ld * nsrc));
  int rc = drs_sssp_batch_dev(g, mode, nsrc, sources_host, d, ld, 0.0, g->stream);
  if (rc == DRS_OK) {
    std::vector<float> tmp((size_t)ld * nsrc);
    cudaError_t e = cudaMemcpy(tmp.data(), d, sizeof(float) * ld * nsrc, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { drs_set_error("drs_sssp_batch: %s", cudaGetErrorString(e)); rc = DRS_ERR_CUDA; }
    else {
      const double inf = std::numeric_limits<double>::infinity();
      for (int k = 0; k < nsrc; ++k)
        for (int j = 0; j < g->n; ++j) {
          double v;
          if (mode == DRS_MODE_F32) v = (double)tmp[(size_t)k * ld + j];
          else { int iv = reinterpret_cast<const int*>(tmp.data())[(size_t)k * ld + j]; v = iv >= DRS_I32_INF ? inf : (double)iv; }
          dist_out_host[(size_t)k * g->n + j] = v;
        }
    }
  }
  cudaFree(d);
  return rc;
}

// All-pairs by N searches: rows [row0, row0+nrows) of the padded matrix (same layout and
// padding as drs_fw_init + the FW rounds produce).
extern "C" int drs_apsp_rows_sssp(const drs_graph* g, int32_t mode, void* local_dev, int64_t ld, int32_t row0, int32_t nrows,
                                  double delta, void* stream) {
  DRS_REQUIRE(g && local_dev, DRS_ERR_ARG, "drs_apsp_rows_sssp: null argument");
  DRS_REQUIRE(ld == g->npad && row0 >= 0 && nrows >= 0 && row0 + nrows <= g->npad, DRS_ERR_ARG, "drs_apsp_rows_sssp: bad shard");
  if (nrows == 0) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int nreal = std::max(0, std::min(nrows, g->n - row0));
  if (nreal < nrows) {
    if (mode == DRS_MODE_F32) pad_rows_kernel<float><<<148, 256, 0, st>>>((float*)local_dev, ld, row0, nreal, nrows);
    else pad_rows_kernel<int><<<148, 256, 0, st>>>((int*)local_dev, ld, row0, nreal, nrows);
    drs_count_launch();
    DRS_CUDA(cudaGetLastError());
  }
  if (nreal == 0) return DRS_OK;
  std::vector<int32_t> src(nreal);
  for (int i = 0; i < nreal; ++i) src[i] = row0 + i;
  return drs_sssp_batch_dev(g, mode, nreal, src.data(), local_dev, ld, delta, stream);
}
