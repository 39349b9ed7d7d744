// C-ABI entry points: graph handle, single-GPU APSP build, queries and paths.
// Declarations and the reference interfaces they replace: include/drs_b200.h.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>

#include "drs_common.cuh"
#include "drs_paths.cuh"

static thread_local char g_err[512] = "";

void drs_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#include <atomic>
static std::atomic<long long> g_launches{0};
void drs_count_launch(int n) { g_launches += n; }
extern "C" int64_t drs_launch_count(void) { return (int64_t)g_launches.load(); }
static std::atomic<long long> g_allocs{0};
void drs_count_alloc(int n) { g_allocs += n; }
extern "C" int64_t drs_alloc_count(void) { return (int64_t)g_allocs.load(); }

extern "C" const char* drs_last_error(void) { return g_err; }
extern "C" int drs_version(void) { return 100; }

namespace {

void free_matrices(drs_graph* g) {
  if (g->owns_matrices) {
    if (g->d_dist) cudaFree(g->d_dist);
    if (g->d_pred) cudaFree(g->d_pred);
  }
  g->d_dist = nullptr;
  g->d_pred = nullptr;
  g->owns_matrices = false;
  g->apsp_valid = false;
}

}  // namespace

// (Re)builds the float weight arrays of both CSRs from h_tt (drs_graph_set_weights; creation fills them in the arena copy).
int drs_graph_upload_weights(drs_graph* g) {
  std::vector<float> out_w(g->narcs), in_w(g->narcs);
  for (int a = 0; a < g->narcs; ++a) {
    out_w[a] = (float)g->h_tt[g->h_out_eid[a]];
    in_w[a] = (float)g->h_tt[g->h_in_eid[a]];
  }
  DRS_CUDA(cudaMemcpy(g->d_out_w, out_w.data(), sizeof(float) * g->narcs, cudaMemcpyHostToDevice));
  DRS_CUDA(cudaMemcpy(g->d_in_w, in_w.data(), sizeof(float) * g->narcs, cudaMemcpyHostToDevice));
  DRS_CUDA(cudaMemcpy(g->d_tt64, g->h_tt.data(), sizeof(double) * g->m, cudaMemcpyHostToDevice));
  return DRS_OK;
}

extern "C" int drs_graph_create(int32_t n, int32_t m, int32_t directed, const int32_t* src, const int32_t* dst,
                                const double* tt, const double* length, drs_graph** out) {
  DRS_REQUIRE(out, DRS_ERR_ARG, "drs_graph_create: out is null");
  *out = nullptr;
  DRS_REQUIRE(n > 0 && m >= 0, DRS_ERR_ARG, "drs_graph_create: need n > 0 and m >= 0 (n=%d m=%d)", n, m);
  DRS_REQUIRE(m == 0 || (src && dst && tt && length), DRS_ERR_ARG, "drs_graph_create: null edge array");
  for (int e = 0; e < m; ++e) {
    DRS_REQUIRE(src[e] >= 0 && src[e] < n && dst[e] >= 0 && dst[e] < n, DRS_ERR_RANGE,
                "drs_graph_create: edge %d endpoint out of range (%d -> %d, n=%d)", e, src[e], dst[e], n);
    DRS_REQUIRE(std::isfinite(tt[e]) && tt[e] > 0.0, DRS_ERR_ARG,
                "drs_graph_create: edge %d travel time must be finite and > 0 (got %g)", e, tt[e]);
    DRS_REQUIRE(std::isfinite(length[e]), DRS_ERR_ARG, "drs_graph_create: edge %d length not finite", e);
  }
  drs_graph* g = new (std::nothrow) drs_graph();
  DRS_REQUIRE(g, DRS_ERR_NOMEM, "drs_graph_create: out of host memory");
  DRS_CUDA(cudaGetDevice(&g->device));
  g->n = n; g->m = m; g->directed = directed ? 1 : 0;
  g->npad = (n + DRS_FW_TILE - 1) / DRS_FW_TILE * DRS_FW_TILE;
  g->h_src.assign(src, src + m); g->h_dst.assign(dst, dst + m);
  g->h_tt.assign(tt, tt + m); g->h_len.assign(length, length + m);

  // arcs: (u, v, eid); undirected edges contribute both directions.  Adjacency
  // order = edge id order (igraph incidence-list order).
  struct Arc { int u, v, e; };
  std::vector<Arc> arcs;
  arcs.reserve((size_t)m * (directed ? 1 : 2));
  for (int e = 0; e < m; ++e) {
    arcs.push_back({src[e], dst[e], e});
    if (!directed && src[e] != dst[e]) arcs.push_back({dst[e], src[e], e});
  }
  g->narcs = (int)arcs.size();
  // pad to n+1 .. npad+1 so that padded vertices have empty adjacency
  std::vector<int32_t> out_ptr(g->npad + 1, 0), in_ptr(g->npad + 1, 0);
  for (const Arc& a : arcs) { out_ptr[a.u + 1]++; in_ptr[a.v + 1]++; }
  std::partial_sum(out_ptr.begin(), out_ptr.end(), out_ptr.begin());
  std::partial_sum(in_ptr.begin(), in_ptr.end(), in_ptr.begin());
  // one host image of every device array (256-byte aligned slices), one allocation, one copy
  const size_t na = (size_t)g->narcs, me = (size_t)m, np1 = (size_t)g->npad + 1;
  size_t off = 0;
  auto slice = [&](size_t bytes) { const size_t o = off; off = (off + std::max<size_t>(bytes, 1) + 255) & ~(size_t)255; return o; };
  const size_t o_tt = slice(me * 8), o_len = slice(me * 8), o_esrc = slice(me * 4), o_edst = slice(me * 4);
  const size_t o_optr = slice(np1 * 4), o_ocol = slice(na * 4), o_oeid = slice(na * 4), o_ow = slice(na * 4);
  const size_t o_iptr = slice(np1 * 4), o_isrc = slice(na * 4), o_ieid = slice(na * 4), o_iw = slice(na * 4);
  std::vector<unsigned char> img(off, 0);
  auto at = [&](size_t o) { return img.data() + o; };
  if (m) {
    std::memcpy(at(o_tt), tt, me * 8); std::memcpy(at(o_len), length, me * 8);
    std::memcpy(at(o_esrc), src, me * 4); std::memcpy(at(o_edst), dst, me * 4);
  }
  std::memcpy(at(o_optr), out_ptr.data(), np1 * 4); std::memcpy(at(o_iptr), in_ptr.data(), np1 * 4);
  int32_t* out_col = (int32_t*)at(o_ocol); int32_t* out_eid = (int32_t*)at(o_oeid); float* out_w = (float*)at(o_ow);
  int32_t* in_src = (int32_t*)at(o_isrc); int32_t* in_eid = (int32_t*)at(o_ieid); float* in_w = (float*)at(o_iw);
  {
    // arcs were generated in edge-id order, so a single pass keeps every adjacency list in edge-id order
    std::vector<int32_t> op(out_ptr.begin(), out_ptr.end() - 1), ip(in_ptr.begin(), in_ptr.end() - 1);
    for (const Arc& a : arcs) {
      const float w = (float)tt[a.e];
      out_col[op[a.u]] = a.v; out_eid[op[a.u]] = a.e; out_w[op[a.u]] = w; op[a.u]++;
      in_src[ip[a.v]] = a.u; in_eid[ip[a.v]] = a.e; in_w[ip[a.v]] = w; ip[a.v]++;
    }
  }
  g->h_out_eid.assign(out_eid, out_eid + na);
  g->h_in_eid.assign(in_eid, in_eid + na);
  auto fail = [&](int code) { drs_graph_destroy(g); return code; };
  if (cudaMalloc(&g->d_arena, img.size()) != cudaSuccess) {
    cudaGetLastError();
    drs_set_error("drs_graph_create: cudaMalloc(%zu bytes) failed", img.size());
    return fail(DRS_ERR_NOMEM);
  }
  if (cudaMemcpy(g->d_arena, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    drs_set_error("drs_graph_create: host-to-device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(DRS_ERR_CUDA);
  }
  unsigned char* base = (unsigned char*)g->d_arena;
  g->d_tt64 = (double*)(base + o_tt); g->d_len64 = (double*)(base + o_len);
  g->d_esrc = (int32_t*)(base + o_esrc); g->d_edst = (int32_t*)(base + o_edst);
  g->d_out_ptr = (int32_t*)(base + o_optr); g->d_out_col = (int32_t*)(base + o_ocol);
  g->d_out_eid = (int32_t*)(base + o_oeid); g->d_out_w = (float*)(base + o_ow);
  g->d_in_ptr = (int32_t*)(base + o_iptr); g->d_in_src = (int32_t*)(base + o_isrc);
  g->d_in_eid = (int32_t*)(base + o_ieid); g->d_in_w = (float*)(base + o_iw);
  if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) {
    drs_set_error("drs_graph_create: cudaStreamCreate failed");
    return fail(DRS_ERR_CUDA);
  }
  *out = g;
  return DRS_OK;
}

extern "C" int drs_graph_set_weights(drs_graph* g, const double* tt) {
  DRS_REQUIRE(g && tt, DRS_ERR_ARG, "drs_graph_set_weights: null argument");
  for (int e = 0; e < g->m; ++e)
    DRS_REQUIRE(std::isfinite(tt[e]) && tt[e] > 0.0, DRS_ERR_ARG,
                "drs_graph_set_weights: edge %d travel time must be finite and > 0 (got %g)", e, tt[e]);
  DrsDeviceGuard guard(g->device);
  g->h_tt.assign(tt, tt + g->m);
  g->apsp_valid = false;
  g->sssp_arcs_valid = false;
  return drs_graph_upload_weights(g);
}

extern "C" int drs_graph_destroy(drs_graph* g) {
  if (!g) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  free_matrices(g);
  if (g->d_arena) cudaFree(g->d_arena);
  if (g->sssp_queues) cudaFree(g->sssp_queues);
  if (g->sssp_arcs) cudaFree(g->sssp_arcs);
  if (g->sssp_sources) cudaFree(g->sssp_sources);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
  return DRS_OK;
}

extern "C" int drs_graph_info(const drs_graph* g, int32_t* n, int32_t* m, int32_t* npad, int32_t* mode,
                              int32_t* apsp_valid, int32_t* device) {
  DRS_REQUIRE(g, DRS_ERR_ARG, "drs_graph_info: null graph");
  if (n) *n = g->n;
  if (m) *m = g->m;
  if (npad) *npad = g->npad;
  if (mode) *mode = g->mode;
  if (apsp_valid) *apsp_valid = g->apsp_valid ? 1 : 0;
  if (device) *device = g->device;
  return DRS_OK;
}

extern "C" int drs_apsp_build(drs_graph* g, int32_t mode) {
  DRS_REQUIRE(g, DRS_ERR_ARG, "drs_apsp_build: null graph");
  DRS_REQUIRE(mode == DRS_MODE_F32 || mode == DRS_MODE_I32, DRS_ERR_ARG, "drs_apsp_build: unknown mode %d", mode);
  if (mode == DRS_MODE_I32) {
    for (int e = 0; e < g->m; ++e)
      DRS_REQUIRE(g->h_tt[e] == std::floor(g->h_tt[e]) && g->h_tt[e] < 1e6, DRS_ERR_ARG,
                  "drs_apsp_build: DRS_MODE_I32 needs integer travel times below 1e6 (edge %d = %g)", e, g->h_tt[e]);
  }
  DrsDeviceGuard guard(g->device);
  const int64_t ld = g->npad;
  if (!g->owns_matrices || g->ld != ld) {
    free_matrices(g);
    DRS_CUDA(cudaMalloc(&g->d_dist, sizeof(float) * ld * ld));
    g->owns_matrices = true;
    g->ld = ld;
  }
  if (g->pred_mode == DRS_PRED_MATRIX && !g->d_pred) DRS_CUDA(cudaMalloc((void**)&g->d_pred, sizeof(int32_t) * ld * ld));
  if (g->pred_mode == DRS_PRED_LAZY && g->d_pred) { cudaFree(g->d_pred); g->d_pred = nullptr; }
  g->apsp_valid = false;
  g->mode = mode;
  int rc;
  const int nb = g->npad / DRS_FW_TILE;
  cudaStream_t st = g->stream;
  int algo = g->algorithm;
  if (algo == DRS_APSP_AUTO) algo = (g->narcs <= 16 * (int64_t)g->n) ? DRS_APSP_SSSP : DRS_APSP_FW;
  // events: [0] start, [1] after init, then per round (after pivot row, after update), last after pred;
  // the SSSP build is one "round" booked under "update"
  const int ner = (algo == DRS_APSP_SSSP) ? 1 : nb;
  std::vector<cudaEvent_t> ev(3 + 2 * ner);
  for (auto& e : ev) DRS_CUDA(cudaEventCreate(&e));
  auto release = [&]() { for (auto& e : ev) cudaEventDestroy(e); };
  DRS_CUDA(cudaEventRecord(ev[0], st));
  if (algo == DRS_APSP_SSSP) {
    DRS_CUDA(cudaEventRecord(ev[1], st));
    DRS_CUDA(cudaEventRecord(ev[2], st));
    if ((rc = drs_apsp_rows_sssp(g, mode, g->d_dist, ld, 0, g->npad, 0.0, st))) { release(); return rc; }
    DRS_CUDA(cudaEventRecord(ev[3], st));
  } else {
  if ((rc = drs_fw_init(g, mode, g->d_dist, ld, 0, g->npad, st))) { release(); return rc; }
  DRS_CUDA(cudaEventRecord(ev[1], st));
  for (int kb = 0; kb < nb; ++kb) {
    char* panel = (char*)g->d_dist + sizeof(float) * (int64_t)kb * DRS_FW_TILE * ld;
    if ((rc = drs_fw_pivot_row(mode, panel, ld, kb, st))) { release(); return rc; }
    DRS_CUDA(cudaEventRecord(ev[2 + 2 * kb], st));
    if ((rc = drs_fw_update(mode, g->d_dist, ld, g->npad, panel, kb, kb, 0, nb, st))) { release(); return rc; }
    DRS_CUDA(cudaEventRecord(ev[3 + 2 * kb], st));
  }
  }
  if (g->d_pred && (rc = drs_pred_build(g, mode, g->d_dist, ld, 0, g->npad, g->d_pred, st))) { release(); return rc; }
  DRS_CUDA(cudaEventRecord(ev[2 + 2 * ner], st));
  DRS_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  g->prof_ms[0] = g->prof_ms[1] = g->prof_ms[2] = g->prof_ms[3] = 0;
  cudaEventElapsedTime(&ms, ev[0], ev[1]); g->prof_ms[0] = ms;
  for (int kb = 0; kb < ner; ++kb) {
    cudaEventElapsedTime(&ms, ev[1 + 2 * kb], ev[2 + 2 * kb]); g->prof_ms[1] += ms;
    cudaEventElapsedTime(&ms, ev[2 + 2 * kb], ev[3 + 2 * kb]); g->prof_ms[2] += ms;
  }
  cudaEventElapsedTime(&ms, ev[1 + 2 * ner], ev[2 + 2 * ner]); g->prof_ms[3] = ms;
  g->prof_rounds = nb;
  release();
  g->apsp_valid = true;
  return DRS_OK;
}

extern "C" int drs_apsp_set_algorithm(drs_graph* g, int32_t algorithm) {
  DRS_REQUIRE(g, DRS_ERR_ARG, "drs_apsp_set_algorithm: null graph");
  DRS_REQUIRE(algorithm == DRS_APSP_FW || algorithm == DRS_APSP_SSSP || algorithm == DRS_APSP_AUTO, DRS_ERR_ARG,
              "drs_apsp_set_algorithm: unknown algorithm %d", algorithm);
  g->algorithm = algorithm;
  return DRS_OK;
}

extern "C" int drs_apsp_set_pred_mode(drs_graph* g, int32_t pred_mode) {
  DRS_REQUIRE(g, DRS_ERR_ARG, "drs_apsp_set_pred_mode: null graph");
  DRS_REQUIRE(pred_mode == DRS_PRED_LAZY || pred_mode == DRS_PRED_MATRIX, DRS_ERR_ARG, "drs_apsp_set_pred_mode: unknown mode %d", pred_mode);
  g->pred_mode = pred_mode;
  return DRS_OK;
}

extern "C" int drs_graph_stream(const drs_graph* g, void** stream) {
  DRS_REQUIRE(g && stream, DRS_ERR_ARG, "drs_graph_stream: null argument");
  *stream = (void*)g->stream;
  return DRS_OK;
}

extern "C" int drs_apsp_last_profile(const drs_graph* g, double* ms, int32_t* rounds) {
  DRS_REQUIRE(g && ms, DRS_ERR_ARG, "drs_apsp_last_profile: null argument");
  for (int i = 0; i < 4; ++i) ms[i] = g->prof_ms[i];
  if (rounds) *rounds = g->prof_rounds;
  return DRS_OK;
}

extern "C" int drs_apsp_matrices(const drs_graph* g, void** dist_dev, int32_t** pred_dev, int64_t* ld) {
  DRS_REQUIRE(g, DRS_ERR_ARG, "drs_apsp_matrices: null graph");
  DRS_REQUIRE(g->apsp_valid, DRS_ERR_STATE, "drs_apsp_matrices: travel-time matrix not built (or invalidated)");
  if (dist_dev) *dist_dev = g->d_dist;
  if (pred_dev) *pred_dev = g->d_pred;
  if (ld) *ld = g->ld;
  return DRS_OK;
}

extern "C" int drs_apsp_adopt(drs_graph* g, int32_t mode, void* dist_dev, int32_t* pred_dev, int64_t ld) {
  DRS_REQUIRE(g && dist_dev, DRS_ERR_ARG, "drs_apsp_adopt: null argument");   // pred_dev may be null (lazy predecessors)
  DRS_REQUIRE(ld == g->npad, DRS_ERR_ARG, "drs_apsp_adopt: ld must equal npad (%d)", g->npad);
  DRS_REQUIRE(mode == DRS_MODE_F32 || mode == DRS_MODE_I32, DRS_ERR_ARG, "drs_apsp_adopt: unknown mode %d", mode);
  DrsDeviceGuard guard(g->device);
  free_matrices(g);
  g->d_dist = dist_dev; g->d_pred = pred_dev; g->ld = ld; g->mode = mode;
  g->owns_matrices = false;
  g->apsp_valid = true;
  return DRS_OK;
}

// ---------------------------------------------------------------------------- queries
namespace {

// hops[q] = number of edges on the canonical path (-1 unreachable / broken chain)
__global__ void path_count_kernel(int npairs, const int32_t* s, const int32_t* t, PredSource ps, int n,
                                  const int32_t* esrc, const int32_t* edst, int32_t* hops) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npairs) return;
  const int src = s[q];
  int cur = t[q], h = 0;
  while (cur != src) {
    const int e = ps.last_edge(src, cur);
    if (e < 0 || h >= n) { h = -1; break; }
    cur = prev_node(e, cur, esrc, edst);
    ++h;
  }
  hops[q] = h;
}

// writes the path of query q in travel order at path[offset[q] ..] and the fp64 left-to-right sums
__global__ void path_fill_kernel(int npairs, const int32_t* s, const int32_t* t, PredSource ps,
                                 const int32_t* esrc, const int32_t* edst, const int32_t* hops,
                                 const int64_t* offset, int32_t* path, const double* tt64, const double* len64,
                                 double* tt_out, double* len_out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npairs) return;
  const int h = hops[q];
  if (h < 0) {
    if (tt_out) tt_out[q] = __longlong_as_double(0x7ff0000000000000LL);
    if (len_out) len_out[q] = __longlong_as_double(0x7ff0000000000000LL);
    return;
  }
  const int src = s[q];
  int cur = t[q];
  int32_t* out = path + offset[q];
  for (int i = h - 1; i >= 0; --i) {
    const int e = ps.last_edge(src, cur);
    out[i] = e;
    cur = prev_node(e, cur, esrc, edst);
  }
  double tt = 0.0, len = 0.0;
  for (int i = 0; i < h; ++i) {
    tt += tt64[out[i]];
    len += len64[out[i]];
  }
  if (tt_out) tt_out[q] = tt;
  if (len_out) len_out[q] = len;
}

__device__ __forceinline__ double dist_to_double(float d) { return (double)d; }
__device__ __forceinline__ double dist_to_double(int d) {
  return d >= DRS_I32_INF ? __longlong_as_double(0x7ff0000000000000LL) : (double)d;
}

template <typename T>
__global__ void matrix_gather_kernel(int npairs, const int32_t* s, const int32_t* t, const T* dist, int64_t ld,
                                     double* out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npairs) return;
  out[q] = dist_to_double(dist[(int64_t)s[q] * ld + t[q]]);
}

template <typename T>
__global__ void rows_to_double_kernel(const T* dist, int64_t ld, int row0, int nrows, int n, double* out) {
  const int64_t total = (int64_t)nrows * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n), j = (int)(idx % n);
    out[idx] = dist_to_double(dist[(int64_t)(row0 + i) * ld + j]);
  }
}

PredSource pred_source(const drs_graph* g) { return drs_pred_source(g); }

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) {
    DRS_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
    return DRS_OK;
  }
  template <typename T> T* as() { return (T*)p; }
};

int check_pairs(const drs_graph* g, int npairs, const int32_t* s, const int32_t* t, const char* who) {
  DRS_REQUIRE(g, DRS_ERR_ARG, "%s: null graph", who);
  DRS_REQUIRE(g->apsp_valid, DRS_ERR_STATE, "%s: travel-time matrix not built (or invalidated by a weight update)", who);
  DRS_REQUIRE(npairs >= 0 && (npairs == 0 || (s && t)), DRS_ERR_ARG, "%s: bad pair arrays", who);
  for (int q = 0; q < npairs; ++q)
    DRS_REQUIRE(s[q] >= 0 && s[q] < g->n && t[q] >= 0 && t[q] < g->n, DRS_ERR_RANGE,
                "%s: pair %d (%d -> %d) out of range (n=%d)", who, q, s[q], t[q], g->n);
  return DRS_OK;
}

}  // namespace

extern "C" int drs_query_pairs(const drs_graph* g, int32_t npairs, const int32_t* s, const int32_t* t,
                               double* tt_out, double* len_out, int32_t* hops_out) {
  int rc = check_pairs(g, npairs, s, t, "drs_query_pairs");
  if (rc) return rc;
  if (npairs == 0) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  cudaStream_t st = g->stream;
  DevBuf ds, dt, dh, doff, dpath, dtt, dlen;
  if ((rc = ds.alloc(sizeof(int32_t) * npairs)) || (rc = dt.alloc(sizeof(int32_t) * npairs)) ||
      (rc = dh.alloc(sizeof(int32_t) * npairs)) || (rc = doff.alloc(sizeof(int64_t) * npairs)) ||
      (rc = dtt.alloc(sizeof(double) * npairs)) || (rc = dlen.alloc(sizeof(double) * npairs)))
    return rc;
  DRS_CUDA(cudaMemcpyAsync(ds.p, s, sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, st));
  DRS_CUDA(cudaMemcpyAsync(dt.p, t, sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, st));
  const int blocks = (npairs + 127) / 128;
  path_count_kernel<<<blocks, 128, 0, st>>>(npairs, ds.as<int32_t>(), dt.as<int32_t>(), pred_source(g), g->n,
                                            g->d_esrc, g->d_edst, dh.as<int32_t>());
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  std::vector<int32_t> hops(npairs);
  DRS_CUDA(cudaMemcpyAsync(hops.data(), dh.p, sizeof(int32_t) * npairs, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  std::vector<int64_t> off(npairs);
  int64_t total = 0;
  for (int q = 0; q < npairs; ++q) { off[q] = total; total += std::max(hops[q], 0); }
  if ((rc = dpath.alloc(sizeof(int32_t) * (size_t)total))) return rc;
  DRS_CUDA(cudaMemcpyAsync(doff.p, off.data(), sizeof(int64_t) * npairs, cudaMemcpyHostToDevice, st));
  path_fill_kernel<<<blocks, 128, 0, st>>>(npairs, ds.as<int32_t>(), dt.as<int32_t>(), pred_source(g), g->d_esrc,
                                           g->d_edst, dh.as<int32_t>(), doff.as<int64_t>(), dpath.as<int32_t>(),
                                           g->d_tt64, g->d_len64, dtt.as<double>(), dlen.as<double>());
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  if (tt_out) DRS_CUDA(cudaMemcpyAsync(tt_out, dtt.p, sizeof(double) * npairs, cudaMemcpyDeviceToHost, st));
  if (len_out) DRS_CUDA(cudaMemcpyAsync(len_out, dlen.p, sizeof(double) * npairs, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  if (hops_out) std::memcpy(hops_out, hops.data(), sizeof(int32_t) * npairs);
  return DRS_OK;
}

extern "C" int drs_query_matrix(const drs_graph* g, int32_t npairs, const int32_t* s, const int32_t* t,
                                double* tt_out) {
  int rc = check_pairs(g, npairs, s, t, "drs_query_matrix");
  if (rc) return rc;
  DRS_REQUIRE(npairs == 0 || tt_out, DRS_ERR_ARG, "drs_query_matrix: null output");
  if (npairs == 0) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  cudaStream_t st = g->stream;
  DevBuf ds, dt, dout;
  if ((rc = ds.alloc(sizeof(int32_t) * npairs)) || (rc = dt.alloc(sizeof(int32_t) * npairs)) ||
      (rc = dout.alloc(sizeof(double) * npairs)))
    return rc;
  DRS_CUDA(cudaMemcpyAsync(ds.p, s, sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, st));
  DRS_CUDA(cudaMemcpyAsync(dt.p, t, sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, st));
  const int blocks = (npairs + 127) / 128;
  if (g->mode == DRS_MODE_F32)
    matrix_gather_kernel<float><<<blocks, 128, 0, st>>>(npairs, ds.as<int32_t>(), dt.as<int32_t>(),
                                                        (const float*)g->d_dist, g->ld, dout.as<double>());
  else
    matrix_gather_kernel<int><<<blocks, 128, 0, st>>>(npairs, ds.as<int32_t>(), dt.as<int32_t>(),
                                                      (const int*)g->d_dist, g->ld, dout.as<double>());
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  DRS_CUDA(cudaMemcpyAsync(tt_out, dout.p, sizeof(double) * npairs, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  return DRS_OK;
}

extern "C" int drs_apsp_rows_to_host(const drs_graph* g, int32_t row0, int32_t nrows, double* out) {
  DRS_REQUIRE(g && out, DRS_ERR_ARG, "drs_apsp_rows_to_host: null argument");
  DRS_REQUIRE(g->apsp_valid, DRS_ERR_STATE, "drs_apsp_rows_to_host: travel-time matrix not built");
  DRS_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= g->n, DRS_ERR_RANGE,
              "drs_apsp_rows_to_host: rows [%d, %d) out of range (n=%d)", row0, row0 + nrows, g->n);
  if (nrows == 0) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  DevBuf tmp;
  int rc;
  const size_t bytes = sizeof(double) * (size_t)nrows * g->n;
  if ((rc = tmp.alloc(bytes))) return rc;
  if (g->mode == DRS_MODE_F32)
    rows_to_double_kernel<float><<<148 * 4, 256, 0, g->stream>>>((const float*)g->d_dist, g->ld, row0, nrows, g->n,
                                                                  tmp.as<double>());
  else
    rows_to_double_kernel<int><<<148 * 4, 256, 0, g->stream>>>((const int*)g->d_dist, g->ld, row0, nrows, g->n,
                                                                tmp.as<double>());
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  DRS_CUDA(cudaMemcpyAsync(out, tmp.p, bytes, cudaMemcpyDeviceToHost, g->stream));
  DRS_CUDA(cudaStreamSynchronize(g->stream));
  return DRS_OK;
}

extern "C" int drs_path(const drs_graph* g, int32_t s, int32_t t, int32_t* edge_ids, int32_t* nedges,
                        int32_t* reachable) {
  DRS_REQUIRE(nedges, DRS_ERR_ARG, "drs_path: nedges is null");
  int rc = check_pairs(g, 1, &s, &t, "drs_path");
  if (rc) return rc;
  const int cap = *nedges;
  DrsDeviceGuard guard(g->device);
  cudaStream_t st = g->stream;
  DevBuf dq, dh, doff, dpath;
  if ((rc = dq.alloc(sizeof(int32_t) * 2)) || (rc = dh.alloc(sizeof(int32_t))) || (rc = doff.alloc(sizeof(int64_t))))
    return rc;
  const int32_t q[2] = {s, t};
  DRS_CUDA(cudaMemcpyAsync(dq.p, q, sizeof(q), cudaMemcpyHostToDevice, st));
  path_count_kernel<<<1, 32, 0, st>>>(1, dq.as<int32_t>(), dq.as<int32_t>() + 1, pred_source(g), g->n, g->d_esrc,
                                      g->d_edst, dh.as<int32_t>());
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  int32_t h = 0;
  DRS_CUDA(cudaMemcpyAsync(&h, dh.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  if (reachable) *reachable = (h >= 0) ? 1 : 0;
  if (h <= 0) { *nedges = 0; return DRS_OK; }
  *nedges = h;
  DRS_REQUIRE(edge_ids && cap >= h, DRS_ERR_CAPACITY, "drs_path: path has %d edges, buffer holds %d", h, cap);
  if ((rc = dpath.alloc(sizeof(int32_t) * h))) return rc;
  DRS_CUDA(cudaMemsetAsync(doff.p, 0, sizeof(int64_t), st));
  path_fill_kernel<<<1, 32, 0, st>>>(1, dq.as<int32_t>(), dq.as<int32_t>() + 1, pred_source(g), g->d_esrc, g->d_edst,
                                     dh.as<int32_t>(), doff.as<int64_t>(), dpath.as<int32_t>(), g->d_tt64, g->d_len64,
                                     nullptr, nullptr);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  DRS_CUDA(cudaMemcpyAsync(edge_ids, dpath.p, sizeof(int32_t) * h, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  return DRS_OK;
}
