// Per-tick device state shared by the dispatch and insertion translation units.
#pragma once
#include "drs_common.cuh"

// Grow-only device allocation with a host staging image: a tick's arrays are slices of a few of these, filled by one
// copy each, so that a steady-state tick allocates nothing (drs_alloc_count() counts every cudaMalloc of the library).
struct DrsArenaLayout {
  size_t total = 0;
  size_t add(size_t bytes) {
    const size_t o = total;
    total = (total + (bytes > 0 ? bytes : 1) + 255) & ~(size_t)255;
    return o;
  }
};
struct DrsArena {
  void* ptr = nullptr;
  size_t cap = 0;
  std::vector<unsigned char> host;
  int reserve(size_t bytes) {
    if (ptr && bytes <= cap) return DRS_OK;
    release();
    const size_t want = bytes + bytes / 4 + 4096;
    drs_count_alloc();
    if (cudaMalloc(&ptr, want) != cudaSuccess) {
      cudaGetLastError();
      ptr = nullptr;
      drs_set_error("device allocation of %zu bytes failed", want);
      return DRS_ERR_NOMEM;
    }
    cap = want;
    return DRS_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

struct TickDev {
  // vehicles
  int n_veh = 0, n_stops = 0;
  int32_t *v_node = nullptr, *v_cap = nullptr, *v_pax = nullptr, *v_off = nullptr;
  double* v_delay = nullptr;
  int32_t *s_node = nullptr, *s_rid = nullptr;
  int8_t *s_pod = nullptr, *s_prev = nullptr;
  double *s_lb = nullptr, *s_ub = nullptr;
  int32_t* v_perm = nullptr;   // vehicles sorted by number of planned stops (uniform work per warp)
  int directed = 0;
  // Window comparisons on a float-weighted map: the reference builds a request's windows from the same fp64 path sums it
  // later compares arrivals with (Ced = Cep + Ts, lib/Agents.py:768), so a direct trip arrives EXACTLY at Ced; an fp32
  // matrix entry differs from that sum by rounding, which would flip such a boundary case.  `snap` (relative, 0 on
  // exact-integer maps) is the slack within which an arrival counts as being on the bound.
  double snap = 0.0;
  // requests
  int n_req = 0;
  int32_t *r_rid = nullptr, *r_o = nullptr, *r_d = nullptr;
  double *r_lbp = nullptr, *r_ubp = nullptr, *r_lbd = nullptr, *r_ubd = nullptr;
};

struct DistView {
  const void* dist;
  int64_t ld;
  int mode;
  __device__ __forceinline__ double at(int a, int b) const {
    if (mode == DRS_MODE_F32) return (double)((const float*)dist)[(int64_t)a * ld + b];
    const int d = ((const int*)dist)[(int64_t)a * ld + b];
    return d >= DRS_I32_INF ? __longlong_as_double(0x7ff0000000000000LL) : (double)d;
  }
};

struct LegacyDev;   // insertion_kernels.cu

struct drs_tick {
  drs_graph* g = nullptr;
  TickDev dev;
  // RV results (device)
  int64_t rv_pairs = 0;
  int rv_nsurv = 0;
  int8_t* d_status = nullptr;
  int32_t *d_surv_veh = nullptr, *d_surv_req = nullptr;
  int* d_surv_count = nullptr;
  double* d_cost = nullptr;
  uint8_t* d_order = nullptr;
  int32_t* d_nstops = nullptr;
  unsigned long long* d_counters = nullptr;
  int64_t surv_capacity = 0;
  LegacyDev* legacy = nullptr;   // plans + request table of the insertion heuristic
  DrsArena a_veh, a_req, a_rv, a_rvres, a_trip;   // vehicles, requests, RV survivors, RV results, trip-evaluation buffers
  double last_ms[4] = {0, 0, 0, 0};
};


void drs_tick_free_plans(drs_tick* t);

// CUDA-event stopwatch on one stream
struct DrsTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  explicit DrsTimer(cudaStream_t s) : st(s) { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~DrsTimer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  void start() { cudaEventRecord(a, st); }
  void stop() { cudaEventRecord(b, st); }
  double ms() { float m = 0; cudaEventSynchronize(b); cudaEventElapsedTime(&m, a, b); return (double)m; }
};
