// Internal definitions shared by the translation units of libdrs_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "drs_b200.h"

void drs_set_error(const char* fmt, ...);
void drs_count_launch(int n = 1);
void drs_count_alloc(int n = 1);   // every cudaMalloc of the library (drs_alloc_count)

#define DRS_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      drs_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,       \
                    __LINE__);                                                              \
      return DRS_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define DRS_REQUIRE(cond, code, ...)    \
  do {                                  \
    if (!(cond)) {                      \
      drs_set_error(__VA_ARGS__);       \
      return (code);                    \
    }                                   \
  } while (0)

struct drs_sep_plan;   // sep_apsp.cu: the part / separator tables and work arrays of the partitioned build

// Host-side graph handle.  Device arrays live on `device`.
struct drs_graph {
  int device = 0;
  int n = 0, m = 0, directed = 0;
  int npad = 0;  // n rounded up to DRS_FW_TILE
  int narcs = 0; // directed arcs in the CSR (m or 2m)

  // host copies of the edge list (edge id = position)
  std::vector<int32_t> h_src, h_dst;
  std::vector<double> h_tt, h_len;
  bool weights_integral = false;   // every travel time is an integer below 1e6: what DRS_MODE_I32 requires

  // host copies of the arc -> edge id maps of both CSRs (weight refresh without a device round trip)
  std::vector<int32_t> h_out_eid, h_in_eid;

  // every array below is a slice of ONE device allocation (d_arena), filled by ONE host-to-device copy
  void* d_arena = nullptr;
  size_t arena_bytes = 0;

  // per-edge attributes on device, indexed by edge id
  int32_t* d_esrc = nullptr;
  int32_t* d_edst = nullptr;
  double* d_tt64 = nullptr;
  double* d_len64 = nullptr;

  // out-adjacency CSR (arc u -> col[a]) and in-adjacency CSR (arc in_src[a] -> v)
  int32_t* d_out_ptr = nullptr;
  int32_t* d_out_col = nullptr;
  int32_t* d_out_eid = nullptr;
  float* d_out_w = nullptr;
  int32_t* d_in_ptr = nullptr;
  int32_t* d_in_src = nullptr;
  int32_t* d_in_eid = nullptr;
  float* d_in_w = nullptr;

  // resident matrices
  int mode = DRS_MODE_F32;
  int algorithm = DRS_APSP_FW;
  int pred_mode = DRS_PRED_LAZY;
  bool apsp_valid = false;
  bool owns_matrices = false;
  void* d_dist = nullptr;
  int32_t* d_pred = nullptr;
  int64_t ld = 0;

  cudaStream_t stream = nullptr;

  // persistent SSSP scratch (sssp_kernels.cu): work queues and packed (target, weight) arcs
  void* sssp_queues = nullptr;
  size_t sssp_queue_bytes = 0;
  void* sssp_arcs = nullptr;
  bool sssp_arcs_valid = false;
  int32_t* sssp_sources = nullptr;   // nsrc source ids followed by the overflow flag
  int sssp_sources_cap = 0;

  // query pool (drs_api.cu): grow-only device scratch and pinned, device-mapped host staging, so that a steady-state
  // query allocates nothing and a scalar query is ONE kernel that writes its answer straight into host memory
  void* q_dev = nullptr;
  size_t q_dev_cap = 0;
  void* q_pin = nullptr;       // cudaHostAlloc(mapped)
  void* q_pin_dev = nullptr;   // its device alias
  size_t q_pin_cap = 0;

  drs_sep_plan* sep = nullptr;   // set by drs_graph_set_partition

  double prof_ms[4] = {0, 0, 0, 0};   // last build, by phase (drs_apsp_last_profile)
  int prof_rounds = 0;
};

int drs_graph_upload_weights(drs_graph* g);
// large device blocks through the process-wide cache (drs_api.cu); *out = nullptr and an error code on failure
int drs_big_alloc(void** out, size_t bytes);
void drs_big_free(void* p, size_t bytes);
void drs_sep_free(drs_graph* g);
bool drs_sep_allowed(const drs_graph* g);  // partition present and the travel times are ones the build accepts
bool drs_sep_usable(const drs_graph* g);   // partition present, exact arithmetic, separators a small part of the map
// DRS_MODE_F32 / DRS_MODE_I32 known, and I32 only on integer travel times below 1e6 (a float weight would be truncated)
int drs_check_mode(const drs_graph* g, int mode, const char* who);

// DRS_TRACE_HOST=1: host-side phase times of the heavier entry points on stderr (development aid)
#include <chrono>
#include <cstdlib>
struct DrsHostTimer {
  const char* who;
  bool on;
  std::chrono::steady_clock::time_point t0;
  explicit DrsHostTimer(const char* w) : who(w) {
    const char* e = getenv("DRS_TRACE_HOST");
    on = e && e[0] == '1';
    t0 = std::chrono::steady_clock::now();
  }
  void lap(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[drs host] %s: %s %.3f ms\n", who, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// RAII device guard
struct DrsDeviceGuard {
  int prev = -1;
  explicit DrsDeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DrsDeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != prev && prev >= 0) cudaSetDevice(prev);
  }
};
