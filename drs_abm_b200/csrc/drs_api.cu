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

// Process-wide cache of the large device blocks (travel-time / predecessor matrices, the work arrays of the separator
// build): a simulation that reloads its map, or a benchmark that builds fresh graphs, gets the blocks of the graph it just
// closed back instead of paying cudaFree + cudaMalloc of 10 GB each time (milliseconds, and a device-wide synchronisation).
// The small per-graph blocks (CSR arena, partition tables, query scratch, pinned staging) go through it too: a fresh graph
// then costs no cudaMalloc / cudaHostAlloc at all.  At most 24 device blocks and 4 pinned ones are kept per process;
// drs_pool_trim() returns them to the driver.
#include <mutex>
namespace {
struct BigBlock { void* p; size_t bytes; int dev; };
struct PinnedBlock { void* host; void* dev_alias; size_t bytes; int dev; };
std::mutex g_pool_mu;
std::vector<BigBlock> g_pool;
std::vector<PinnedBlock> g_pinned;
constexpr size_t kPoolMin = 256;
constexpr size_t kPoolSlots = 24;
}  // namespace

int drs_big_alloc(void** out, size_t bytes) {
  *out = nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  if (bytes >= kPoolMin) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pool.size(); ++i)
      if (g_pool[i].dev == dev && g_pool[i].bytes >= bytes && g_pool[i].bytes <= bytes + bytes / 4 + (1 << 16) &&
          (best < 0 || g_pool[i].bytes < g_pool[best].bytes))
        best = i;
    if (best >= 0) {
      *out = g_pool[best].p;
      g_pool.erase(g_pool.begin() + best);
      return DRS_OK;
    }
  }
  drs_count_alloc();
  if (cudaMalloc(out, bytes) != cudaSuccess) {
    cudaGetLastError();
    drs_pool_trim();                       // the cache may be what is in the way
    if (cudaMalloc(out, bytes) != cudaSuccess) {
      cudaGetLastError();
      *out = nullptr;
      DRS_REQUIRE(false, DRS_ERR_NOMEM, "cudaMalloc(%zu bytes) failed", bytes);
    }
  }
  return DRS_OK;
}

void drs_big_free(void* p, size_t bytes) {
  if (!p) return;
  int dev = 0;
  cudaGetDevice(&dev);
  if (bytes >= kPoolMin) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    void* evicted = nullptr;
    if (g_pool.size() >= kPoolSlots) {      // full: the block that has waited longest goes back to the driver
      evicted = g_pool.front().p;
      if (g_pool.front().dev != dev) cudaSetDevice(g_pool.front().dev);
      cudaFree(evicted);
      if (g_pool.front().dev != dev) cudaSetDevice(dev);
      g_pool.erase(g_pool.begin());
    }
    g_pool.push_back({p, bytes, dev});
    return;
  }
  cudaFree(p);
}

extern "C" int drs_pool_trim(void) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (const BigBlock& b : g_pool) { cudaSetDevice(b.dev); cudaFree(b.p); }
  for (const PinnedBlock& b : g_pinned) { cudaSetDevice(b.dev); cudaFreeHost(b.host); }
  cudaSetDevice(cur);
  g_pool.clear();
  g_pinned.clear();
  return DRS_OK;
}

// pinned, device-mapped staging of at least `bytes` (from the cache when one fits)
static int pinned_alloc(size_t bytes, void** host, void** dev_alias, size_t* got) {
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pinned.size(); ++i)
      if (g_pinned[i].dev == dev && g_pinned[i].bytes >= bytes) {
        *host = g_pinned[i].host; *dev_alias = g_pinned[i].dev_alias; *got = g_pinned[i].bytes;
        g_pinned.erase(g_pinned.begin() + i);
        return DRS_OK;
      }
  }
  drs_count_alloc();
  if (cudaHostAlloc(host, bytes, cudaHostAllocMapped) != cudaSuccess || cudaHostGetDevicePointer(dev_alias, *host, 0) != cudaSuccess) {
    cudaGetLastError();
    DRS_REQUIRE(false, DRS_ERR_NOMEM, "query staging: cudaHostAlloc(%zu bytes) failed", bytes);
  }
  *got = bytes;
  return DRS_OK;
}

static void pinned_free(void* host, void* dev_alias, size_t bytes) {
  if (!host) return;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pinned.size() < 4) { g_pinned.push_back({host, dev_alias, bytes, dev}); return; }
  }
  cudaFreeHost(host);
}

namespace {

void free_matrices(drs_graph* g) {
  if (g->owns_matrices) {
    const size_t bytes = sizeof(float) * (size_t)g->ld * (size_t)g->ld;
    if (g->d_dist) drs_big_free(g->d_dist, bytes);
    if (g->d_pred) drs_big_free(g->d_pred, bytes);
  }
  g->d_dist = nullptr;
  g->d_pred = nullptr;
  g->owns_matrices = false;
  g->apsp_valid = false;
}

}  // namespace

static void classify_weights(drs_graph* g) {
  g->weights_integral = true;
  for (double w : g->h_tt)
    if (!(w == std::floor(w) && w < 1e6)) { g->weights_integral = false; break; }
}

int drs_check_mode(const drs_graph* g, int mode, const char* who) {
  DRS_REQUIRE(mode == DRS_MODE_F32 || mode == DRS_MODE_I32, DRS_ERR_ARG, "%s: unknown mode %d", who, mode);
  DRS_REQUIRE(mode != DRS_MODE_I32 || g->weights_integral, DRS_ERR_ARG,
              "%s: DRS_MODE_I32 needs integer travel times below 1e6 (this map has others; use DRS_MODE_F32)", who);
  return DRS_OK;
}

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
  DrsHostTimer tm("drs_graph_create");
  for (int e = 0; e < m; ++e) {
    DRS_REQUIRE(src[e] >= 0 && src[e] < n && dst[e] >= 0 && dst[e] < n, DRS_ERR_RANGE,
                "drs_graph_create: edge %d endpoint out of range (%d -> %d, n=%d)", e, src[e], dst[e], n);
    DRS_REQUIRE(std::isfinite(tt[e]) && tt[e] > 0.0, DRS_ERR_ARG,
                "drs_graph_create: edge %d travel time must be finite and > 0 (got %g)", e, tt[e]);
    DRS_REQUIRE(std::isfinite(length[e]), DRS_ERR_ARG, "drs_graph_create: edge %d length not finite", e);
  }
  tm.lap("validate");
  drs_graph* g = new (std::nothrow) drs_graph();
  DRS_REQUIRE(g, DRS_ERR_NOMEM, "drs_graph_create: out of host memory");
  DRS_CUDA(cudaGetDevice(&g->device));
  g->n = n; g->m = m; g->directed = directed ? 1 : 0;
  g->npad = (n + DRS_FW_TILE - 1) / DRS_FW_TILE * DRS_FW_TILE;
  g->h_src.assign(src, src + m); g->h_dst.assign(dst, dst + m);
  g->h_tt.assign(tt, tt + m); g->h_len.assign(length, length + m);
  classify_weights(g);

  // arcs u -> v of edge e; undirected edges contribute both directions.  Adjacency order = edge id order (igraph
  // incidence-list order): two passes over the edge list, no intermediate arc array.
  // pad to n+1 .. npad+1 so that padded vertices have empty adjacency
  std::vector<int32_t> out_ptr(g->npad + 1, 0), in_ptr(g->npad + 1, 0);
  int narcs = 0;
  for (int e = 0; e < m; ++e) {
    const int u = src[e], v = dst[e];
    out_ptr[u + 1]++; in_ptr[v + 1]++; ++narcs;
    if (!directed && u != v) { out_ptr[v + 1]++; in_ptr[u + 1]++; ++narcs; }
  }
  g->narcs = narcs;
  std::partial_sum(out_ptr.begin(), out_ptr.end(), out_ptr.begin());
  std::partial_sum(in_ptr.begin(), in_ptr.end(), in_ptr.begin());
  // one host image of every device array (256-byte aligned slices), one allocation, one copy
  const size_t na = (size_t)g->narcs, me = (size_t)m, np1 = (size_t)g->npad + 1;
  size_t off = 0;
  auto slice = [&](size_t bytes) { const size_t o = off; off = (off + std::max<size_t>(bytes, 1) + 255) & ~(size_t)255; return o; };
  const size_t o_tt = slice(me * 8), o_len = slice(me * 8), o_esrc = slice(me * 4), o_edst = slice(me * 4);
  const size_t o_optr = slice(np1 * 4), o_ocol = slice(na * 4), o_oeid = slice(na * 4), o_ow = slice(na * 4);
  const size_t o_iptr = slice(np1 * 4), o_isrc = slice(na * 4), o_ieid = slice(na * 4), o_iw = slice(na * 4);
  std::vector<unsigned char> img(off);
  auto at = [&](size_t o) { return img.data() + o; };
  if (m) {
    std::memcpy(at(o_tt), tt, me * 8); std::memcpy(at(o_len), length, me * 8);
    std::memcpy(at(o_esrc), src, me * 4); std::memcpy(at(o_edst), dst, me * 4);
  }
  std::memcpy(at(o_optr), out_ptr.data(), np1 * 4); std::memcpy(at(o_iptr), in_ptr.data(), np1 * 4);
  int32_t* out_col = (int32_t*)at(o_ocol); int32_t* out_eid = (int32_t*)at(o_oeid); float* out_w = (float*)at(o_ow);
  int32_t* in_src = (int32_t*)at(o_isrc); int32_t* in_eid = (int32_t*)at(o_ieid); float* in_w = (float*)at(o_iw);
  {
    // edges are visited in id order, so every adjacency list comes out in edge-id order
    std::vector<int32_t> op(out_ptr.begin(), out_ptr.end() - 1), ip(in_ptr.begin(), in_ptr.end() - 1);
    auto put = [&](int u, int v, int e, float w) {
      out_col[op[u]] = v; out_eid[op[u]] = e; out_w[op[u]] = w; op[u]++;
      in_src[ip[v]] = u; in_eid[ip[v]] = e; in_w[ip[v]] = w; ip[v]++;
    };
    for (int e = 0; e < m; ++e) {
      const float w = (float)tt[e];
      put(src[e], dst[e], e, w);
      if (!directed && src[e] != dst[e]) put(dst[e], src[e], e, w);
    }
  }
  g->h_out_eid.assign(out_eid, out_eid + na);
  g->h_in_eid.assign(in_eid, in_eid + na);
  auto fail = [&](int code) { drs_graph_destroy(g); return code; };
  tm.lap("host image");
  g->arena_bytes = img.size();
  if (drs_big_alloc(&g->d_arena, img.size())) return fail(DRS_ERR_NOMEM);
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
  tm.lap("malloc + copy + stream");
  *out = g;
  return DRS_OK;
}

// Vertex ids in the order of the DEVICE numbering (order_out[k] = caller's id of device vertex k): along a Morton
// (Z-order) curve over the ranks of the distinct lng / lat values, at most 16 bits each -- a regular grid gets its exact
// Z-curve, an irregular map a density-equalised one.  The kernels index distance rows by vertex id, so neighbouring
// vertices should have neighbouring ids (the same 50k-node map builds in 84 ms numbered this way, 93 ms row-major,
// 141 ms at random).  Host-side helper of the bindings (routing.RoadGraph); returns DRS_ERR_ARG when the coordinates are
// unusable (non-finite, or all equal), in which case the caller falls back to a bandwidth-reducing order of its own.
extern "C" int drs_locality_order(int32_t n, const double* lng, const double* lat, int32_t* order_out) {
  DRS_REQUIRE(n > 0 && lng && lat && order_out, DRS_ERR_ARG, "drs_locality_order: bad argument");
  auto ranks = [&](const double* x, std::vector<uint32_t>& r) -> int {
    std::vector<std::pair<double, int32_t>> v(n);
    for (int i = 0; i < n; ++i) {
      if (!std::isfinite(x[i])) return -1;
      v[i] = {x[i], i};
    }
    std::sort(v.begin(), v.end());
    uint32_t distinct = 0;
    r.assign(n, 0);
    for (int i = 0; i < n; ++i) {
      if (i > 0 && v[i].first != v[i - 1].first) ++distinct;
      r[v[i].second] = distinct;
    }
    const uint32_t count = distinct + 1;
    if (count > 65536)
      for (int i = 0; i < n; ++i) r[i] = (uint32_t)(((uint64_t)r[i] * 65536) / count);
    return (int)count;
  };
  std::vector<uint32_t> rx, ry;
  const int cx = ranks(lng, rx), cy = ranks(lat, ry);
  DRS_REQUIRE(cx > 0 && cy > 0 && !(cx == 1 && cy == 1), DRS_ERR_ARG, "drs_locality_order: coordinates are not usable");
  auto spread = [](uint32_t v) {
    uint64_t x = v & 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0full;
    x = (x | (x << 2)) & 0x33333333ull;
    x = (x | (x << 1)) & 0x55555555ull;
    return x;
  };
  std::vector<std::pair<uint64_t, int32_t>> keys(n);
  for (int i = 0; i < n; ++i) keys[i] = {spread(rx[i]) | (spread(ry[i]) << 1), i};
  std::sort(keys.begin(), keys.end());       // ties by caller's id: a stable order
  for (int i = 0; i < n; ++i) order_out[i] = keys[i].second;
  return DRS_OK;
}

extern "C" int drs_graph_set_weights(drs_graph* g, const double* tt) {
  DRS_REQUIRE(g && tt, DRS_ERR_ARG, "drs_graph_set_weights: null argument");
  for (int e = 0; e < g->m; ++e)
    DRS_REQUIRE(std::isfinite(tt[e]) && tt[e] > 0.0, DRS_ERR_ARG,
                "drs_graph_set_weights: edge %d travel time must be finite and > 0 (got %g)", e, tt[e]);
  DrsDeviceGuard guard(g->device);
  g->h_tt.assign(tt, tt + g->m);
  classify_weights(g);
  g->apsp_valid = false;
  g->sssp_arcs_valid = false;
  return drs_graph_upload_weights(g);
}

extern "C" int drs_graph_destroy(drs_graph* g) {
  if (!g) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  free_matrices(g);
  drs_sep_free(g);
  if (g->d_arena) drs_big_free(g->d_arena, g->arena_bytes);
  if (g->sssp_queues) cudaFree(g->sssp_queues);
  if (g->sssp_arcs) cudaFree(g->sssp_arcs);
  if (g->sssp_sources) cudaFree(g->sssp_sources);
  if (g->q_dev) drs_big_free(g->q_dev, g->q_dev_cap);
  pinned_free(g->q_pin, g->q_pin_dev, g->q_pin_cap);
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
  { int rcm = drs_check_mode(g, mode, "drs_apsp_build"); if (rcm) return rcm; }
  DrsDeviceGuard guard(g->device);
  const int64_t ld = g->npad;
  DrsHostTimer tm("drs_apsp_build");
  if (!g->owns_matrices || g->ld != ld) {
    free_matrices(g);
    { int rca = drs_big_alloc(&g->d_dist, sizeof(float) * (size_t)ld * (size_t)ld); if (rca) return rca; }
    g->owns_matrices = true;
    g->ld = ld;
  }
  tm.lap("matrix malloc");
  if (g->pred_mode == DRS_PRED_MATRIX && !g->d_pred) {
    int rca = drs_big_alloc((void**)&g->d_pred, sizeof(int32_t) * (size_t)ld * (size_t)ld);
    if (rca) return rca;
  }
  if (g->pred_mode == DRS_PRED_LAZY && g->d_pred) { drs_big_free(g->d_pred, sizeof(int32_t) * (size_t)ld * (size_t)ld); g->d_pred = nullptr; }
  g->apsp_valid = false;
  g->mode = mode;
  int rc;
  const int nb = g->npad / DRS_FW_TILE;
  cudaStream_t st = g->stream;
  int algo = g->algorithm;
  if (algo == DRS_APSP_AUTO)
    algo = drs_sep_usable(g) ? DRS_APSP_SEP : (g->narcs <= 16 * (int64_t)g->n) ? DRS_APSP_SSSP : DRS_APSP_FW;
  if (algo == DRS_APSP_SEP)
    DRS_REQUIRE(g->sep && drs_sep_allowed(g), DRS_ERR_STATE,
                "drs_apsp_build: DRS_APSP_SEP needs a partition (drs_graph_set_partition); with DRS_SEP_FLOAT=0 also integer travel times");
  // events: [0] start, [1] after init, then per round (after pivot row, after update), last after pred;
  // the SSSP build is one "round" booked under "update"
  const int ner = (algo == DRS_APSP_SSSP || algo == DRS_APSP_SEP) ? 1 : nb;
  std::vector<cudaEvent_t> ev(3 + 2 * ner);
  for (auto& e : ev) DRS_CUDA(cudaEventCreate(&e));
  auto release = [&]() { for (auto& e : ev) cudaEventDestroy(e); };
  DRS_CUDA(cudaEventRecord(ev[0], st));
  if (algo == DRS_APSP_SSSP) {
    DRS_CUDA(cudaEventRecord(ev[1], st));
    DRS_CUDA(cudaEventRecord(ev[2], st));
    if ((rc = drs_apsp_rows_sssp(g, mode, g->d_dist, ld, 0, g->npad, 0.0, st))) { release(); return rc; }
    DRS_CUDA(cudaEventRecord(ev[3], st));
  } else if (algo == DRS_APSP_SEP) {
    DRS_CUDA(cudaEventRecord(ev[1], st));
    DRS_CUDA(cudaEventRecord(ev[2], st));
    if ((rc = drs_apsp_rows_sep(g, mode, g->d_dist, ld, 0, g->npad, st))) { release(); return rc; }
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
  tm.lap("build");
  g->apsp_valid = true;
  return DRS_OK;
}

extern "C" int drs_apsp_set_algorithm(drs_graph* g, int32_t algorithm) {
  DRS_REQUIRE(g, DRS_ERR_ARG, "drs_apsp_set_algorithm: null graph");
  DRS_REQUIRE(algorithm == DRS_APSP_FW || algorithm == DRS_APSP_SSSP || algorithm == DRS_APSP_AUTO || algorithm == DRS_APSP_SEP, DRS_ERR_ARG,
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

// One query = one thread: walks the canonical path s -> t BACKWARDS (the predecessor rule gives the last edge), keeping
// the edge ids in the query's scratch segment, then forms the fp64 sums LEFT TO RIGHT in travel order, which is what
// the reference's loops over the epath do (lib/RoutingEngine.py:199-203, 212-216) -- so get_duration is bit-identical to
// the sum of the step durations get_routing reports (invariant lib/Agents.py:358, 999-1000).  A path longer than the
// segment sets hops to its true length and status 1: the host repeats the query with a larger segment.
struct PathAnswer { double tt, len; int32_t hops, status; };

__global__ void path_query_kernel(int npairs, const int32_t* s, const int32_t* t, PredSource ps, int n, const int32_t* esrc,
                                  const int32_t* edst, const double* tt64, const double* len64, int32_t* scratch, int seg,
                                  PathAnswer* out, int32_t* path_out /* may be null: forward edge ids of query 0 */) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npairs) return;
  const int src = s[q];
  int cur = t[q], h = 0;
  int32_t* mine = scratch + (int64_t)q * seg;
  bool broken = false;
  while (cur != src) {
    const int e = ps.last_edge(src, cur);
    if (e < 0 || h >= n) { broken = true; break; }
    if (h < seg) mine[h] = e;
    cur = prev_node(e, cur, esrc, edst);
    ++h;
  }
  PathAnswer a;
  a.status = 0;
  if (broken) {
    a.hops = -1; a.tt = a.len = __longlong_as_double(0x7ff0000000000000LL);
  } else if (h > seg) {
    a.hops = h; a.status = 1; a.tt = a.len = 0.0;
  } else {
    double tt = 0.0, len = 0.0;
    for (int i = h - 1; i >= 0; --i) {
      const int e = mine[i];
      tt += tt64[e];
      len += len64[e];
      if (path_out) path_out[h - 1 - i] = e;
    }
    a.hops = h; a.tt = tt; a.len = len;
  }
  out[q] = a;
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

int check_pairs(const drs_graph* g, int npairs, const int32_t* s, const int32_t* t, const char* who) {
  DRS_REQUIRE(g, DRS_ERR_ARG, "%s: null graph", who);
  DRS_REQUIRE(g->apsp_valid, DRS_ERR_STATE, "%s: travel-time matrix not built (or invalidated by a weight update)", who);
  DRS_REQUIRE(npairs >= 0 && (npairs == 0 || (s && t)), DRS_ERR_ARG, "%s: bad pair arrays", who);
  for (int q = 0; q < npairs; ++q)
    DRS_REQUIRE(s[q] >= 0 && s[q] < g->n && t[q] >= 0 && t[q] < g->n, DRS_ERR_RANGE,
                "%s: pair %d (%d -> %d) out of range (n=%d)", who, q, s[q], t[q], g->n);
  return DRS_OK;
}

// grow-only pools on the graph handle
int pool_device(const drs_graph* gc, size_t bytes, void** out) {
  drs_graph* g = const_cast<drs_graph*>(gc);
  if (g->q_dev_cap < bytes) {
    if (g->q_dev) drs_big_free(g->q_dev, g->q_dev_cap);
    g->q_dev = nullptr; g->q_dev_cap = 0;
    const size_t want = std::max<size_t>(bytes + bytes / 2, 1 << 20);
    { int rca = drs_big_alloc(&g->q_dev, want); if (rca) return rca; }
    g->q_dev_cap = want;
  }
  *out = g->q_dev;
  return DRS_OK;
}

int pool_pinned(const drs_graph* gc, size_t bytes, void** host, void** dev) {
  drs_graph* g = const_cast<drs_graph*>(gc);
  if (g->q_pin_cap < bytes) {
    pinned_free(g->q_pin, g->q_pin_dev, g->q_pin_cap);
    g->q_pin = g->q_pin_dev = nullptr; g->q_pin_cap = 0;
    const size_t want = std::max<size_t>(bytes + bytes / 2, 1 << 16);
    size_t got = 0;
    { int rca = pinned_alloc(want, &g->q_pin, &g->q_pin_dev, &got); if (rca) return rca; }
    g->q_pin_cap = got;
  }
  *host = g->q_pin; *dev = g->q_pin_dev;
  return DRS_OK;
}

constexpr int SMALL_QUERY = 2048;      // up to this many pairs a query is ONE kernel on mapped host memory (no copies)
constexpr int PATH_SEG_SMALL = 1024;   // path edges kept per query on the first try

// Path sums of npairs queries.  Small batches: the kernel reads the pairs from, and writes the answers to, pinned host
// memory mapped into the device (one launch, one stream synchronisation, no copy).  Large batches: staged copies.
int run_path_queries(const drs_graph* g, int npairs, const int32_t* s, const int32_t* t, double* tt_out, double* len_out,
                     int32_t* hops_out, int32_t* path_out_host, int path_cap, int* path_len) {
  cudaStream_t st = g->stream;
  int seg = PATH_SEG_SMALL;   // edges kept per query on the first try (the caller chunks batches to bound the scratch)
  for (int attempt = 0; attempt < 2; ++attempt) {
    const size_t in_bytes = sizeof(int32_t) * 2 * (size_t)npairs, ans_bytes = sizeof(PathAnswer) * (size_t)npairs;
    const size_t path_bytes = path_out_host ? sizeof(int32_t) * (size_t)seg : 0;
    const size_t ans_off = (in_bytes + 15) & ~(size_t)15, path_off = (ans_off + ans_bytes + 15) & ~(size_t)15;
    void *hp, *hd, *dscratch;
    int rc;
    if ((rc = pool_pinned(g, path_off + path_bytes, &hp, &hd))) return rc;
    const bool small = npairs <= SMALL_QUERY;
    const size_t scratch_bytes = sizeof(int32_t) * (size_t)seg * npairs;
    const size_t dev_in_off = (scratch_bytes + 255) & ~(size_t)255, dev_ans_off = (dev_in_off + in_bytes + 255) & ~(size_t)255;
    if ((rc = pool_device(g, small ? scratch_bytes : dev_ans_off + ans_bytes, &dscratch))) return rc;
    int32_t* h_in = (int32_t*)hp;
    std::memcpy(h_in, s, sizeof(int32_t) * npairs);
    std::memcpy(h_in + npairs, t, sizeof(int32_t) * npairs);
    PathAnswer* h_ans = (PathAnswer*)((char*)hp + ans_off);
    const int32_t *k_s, *k_t;
    PathAnswer* k_ans;
    int32_t* k_path = path_out_host ? (int32_t*)((char*)hd + path_off) : nullptr;
    if (small) {
      k_s = (const int32_t*)hd; k_t = k_s + npairs; k_ans = (PathAnswer*)((char*)hd + ans_off);
    } else {
      char* db = (char*)dscratch;
      DRS_CUDA(cudaMemcpyAsync(db + dev_in_off, h_in, in_bytes, cudaMemcpyHostToDevice, st));
      k_s = (const int32_t*)(db + dev_in_off); k_t = k_s + npairs; k_ans = (PathAnswer*)(db + dev_ans_off);
    }
    const int threads = npairs < 128 ? 32 : 128;
    path_query_kernel<<<(npairs + threads - 1) / threads, threads, 0, st>>>(npairs, k_s, k_t, pred_source(g), g->n, g->d_esrc, g->d_edst,
                                                                          g->d_tt64, g->d_len64, (int32_t*)dscratch, seg, k_ans, k_path);
    drs_count_launch();
    DRS_CUDA(cudaGetLastError());
    if (!small) DRS_CUDA(cudaMemcpyAsync(h_ans, k_ans, ans_bytes, cudaMemcpyDeviceToHost, st));
    DRS_CUDA(cudaStreamSynchronize(st));
    int longest = 0;
    for (int q = 0; q < npairs; ++q)
      if (h_ans[q].status == 1) longest = std::max(longest, h_ans[q].hops);
    if (longest == 0 || attempt == 1) {
      DRS_REQUIRE(longest == 0, DRS_ERR_STATE, "path query: a path of %d edges did not fit the second attempt", longest);
      for (int q = 0; q < npairs; ++q) {
        if (tt_out) tt_out[q] = h_ans[q].tt;
        if (len_out) len_out[q] = h_ans[q].len;
        if (hops_out) hops_out[q] = h_ans[q].hops;
      }
      if (path_out_host) {
        const int h = std::max(h_ans[0].hops, 0);
        *path_len = h_ans[0].hops;
        if (h <= path_cap) std::memcpy(path_out_host, (char*)hp + path_off, sizeof(int32_t) * h);
      }
      return DRS_OK;
    }
    seg = longest;   // at least one path outgrew its segment: once more with room for the longest
  }
  return DRS_OK;
}

}  // namespace

extern "C" int drs_query_pairs(const drs_graph* g, int32_t npairs, const int32_t* s, const int32_t* t,
                               double* tt_out, double* len_out, int32_t* hops_out) {
  int rc = check_pairs(g, npairs, s, t, "drs_query_pairs");
  if (rc) return rc;
  if (npairs == 0) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  const int chunk = 1 << 16;   // 256 MB of path scratch at most
  for (int q0 = 0; q0 < npairs; q0 += chunk) {
    const int cnt = std::min(chunk, npairs - q0);
    rc = run_path_queries(g, cnt, s + q0, t + q0, tt_out ? tt_out + q0 : nullptr, len_out ? len_out + q0 : nullptr,
                          hops_out ? hops_out + q0 : nullptr, nullptr, 0, nullptr);
    if (rc) return rc;
  }
  return DRS_OK;
}

// Scalar form of drs_query_pairs (RoutingEngine.get_duration / get_distance are called one pair at a time,
// lib/Agents.py:898-923, 756, 1073): one kernel, answer written straight into mapped host memory.
extern "C" int drs_query_pair(const drs_graph* g, int32_t s, int32_t t, double* tt_out, double* len_out, int32_t* hops_out) {
  int rc = check_pairs(g, 1, &s, &t, "drs_query_pair");
  if (rc) return rc;
  DrsDeviceGuard guard(g->device);
  return run_path_queries(g, 1, &s, &t, tt_out, len_out, hops_out, nullptr, 0, nullptr);
}

extern "C" int drs_query_matrix(const drs_graph* g, int32_t npairs, const int32_t* s, const int32_t* t,
                                double* tt_out) {
  int rc = check_pairs(g, npairs, s, t, "drs_query_matrix");
  if (rc) return rc;
  DRS_REQUIRE(npairs == 0 || tt_out, DRS_ERR_ARG, "drs_query_matrix: null output");
  if (npairs == 0) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  cudaStream_t st = g->stream;
  const size_t in_bytes = sizeof(int32_t) * 2 * (size_t)npairs, out_bytes = sizeof(double) * (size_t)npairs;
  const size_t out_off = (in_bytes + 15) & ~(size_t)15;
  void *hp, *hd;
  if ((rc = pool_pinned(g, out_off + out_bytes, &hp, &hd))) return rc;
  int32_t* h_in = (int32_t*)hp;
  std::memcpy(h_in, s, sizeof(int32_t) * npairs);
  std::memcpy(h_in + npairs, t, sizeof(int32_t) * npairs);
  const bool small = npairs <= SMALL_QUERY;
  const int32_t *k_s, *k_t;
  double* k_out;
  if (small) {
    k_s = (const int32_t*)hd; k_t = k_s + npairs; k_out = (double*)((char*)hd + out_off);
  } else {
    void* db;
    if ((rc = pool_device(g, out_off + out_bytes, &db))) return rc;
    DRS_CUDA(cudaMemcpyAsync(db, h_in, in_bytes, cudaMemcpyHostToDevice, st));
    k_s = (const int32_t*)db; k_t = k_s + npairs; k_out = (double*)((char*)db + out_off);
  }
  const int threads = npairs < 128 ? 32 : 128, blocks = (npairs + threads - 1) / threads;
  if (g->mode == DRS_MODE_F32)
    matrix_gather_kernel<float><<<blocks, threads, 0, st>>>(npairs, k_s, k_t, (const float*)g->d_dist, g->ld, k_out);
  else
    matrix_gather_kernel<int><<<blocks, threads, 0, st>>>(npairs, k_s, k_t, (const int*)g->d_dist, g->ld, k_out);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  if (!small) DRS_CUDA(cudaMemcpyAsync((char*)hp + out_off, k_out, out_bytes, cudaMemcpyDeviceToHost, st));
  DRS_CUDA(cudaStreamSynchronize(st));
  std::memcpy(tt_out, (char*)hp + out_off, out_bytes);
  return DRS_OK;
}

extern "C" int drs_apsp_rows_to_host(const drs_graph* g, int32_t row0, int32_t nrows, double* out) {
  DRS_REQUIRE(g && out, DRS_ERR_ARG, "drs_apsp_rows_to_host: null argument");
  DRS_REQUIRE(g->apsp_valid, DRS_ERR_STATE, "drs_apsp_rows_to_host: travel-time matrix not built");
  DRS_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= g->n, DRS_ERR_RANGE,
              "drs_apsp_rows_to_host: rows [%d, %d) out of range (n=%d)", row0, row0 + nrows, g->n);
  if (nrows == 0) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  int rc;
  // at most 256 MB of converted rows at a time through the pooled scratch
  const int chunk = (int)std::max<int64_t>(1, std::min<int64_t>(nrows, ((int64_t)256 << 20) / (8 * (int64_t)g->n)));
  void* tmp;
  if ((rc = pool_device(g, sizeof(double) * (size_t)chunk * g->n, &tmp))) return rc;
  for (int r = 0; r < nrows; r += chunk) {
    const int cnt = std::min(chunk, nrows - r);
    const size_t bytes = sizeof(double) * (size_t)cnt * g->n;
    if (g->mode == DRS_MODE_F32)
      rows_to_double_kernel<float><<<148 * 4, 256, 0, g->stream>>>((const float*)g->d_dist, g->ld, row0 + r, cnt, g->n, (double*)tmp);
    else
      rows_to_double_kernel<int><<<148 * 4, 256, 0, g->stream>>>((const int*)g->d_dist, g->ld, row0 + r, cnt, g->n, (double*)tmp);
    drs_count_launch();
    DRS_CUDA(cudaGetLastError());
    DRS_CUDA(cudaMemcpyAsync(out + (size_t)r * g->n, tmp, bytes, cudaMemcpyDeviceToHost, g->stream));
    DRS_CUDA(cudaStreamSynchronize(g->stream));
  }
  return DRS_OK;
}

extern "C" int drs_path(const drs_graph* g, int32_t s, int32_t t, int32_t* edge_ids, int32_t* nedges,
                        int32_t* reachable) {
  DRS_REQUIRE(nedges, DRS_ERR_ARG, "drs_path: nedges is null");
  int rc = check_pairs(g, 1, &s, &t, "drs_path");
  if (rc) return rc;
  const int cap = *nedges;
  DrsDeviceGuard guard(g->device);
  // one kernel: the path's edge ids land in mapped host memory in travel order
  std::vector<int32_t> dummy;
  int32_t one = 0;
  int32_t* dst = (edge_ids && cap > 0) ? edge_ids : &one;
  int h = 0;
  if ((rc = run_path_queries(g, 1, &s, &t, nullptr, nullptr, nullptr, dst, (edge_ids && cap > 0) ? cap : 0, &h))) return rc;
  if (reachable) *reachable = (h >= 0) ? 1 : 0;
  if (h <= 0) { *nedges = 0; return DRS_OK; }
  *nedges = h;
  DRS_REQUIRE(edge_ids && cap >= h, DRS_ERR_CAPACITY, "drs_path: path has %d edges, buffer holds %d", h, cap);
  return DRS_OK;
}
