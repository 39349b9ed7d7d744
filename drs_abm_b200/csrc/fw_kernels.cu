// Blocked min-plus Floyd-Warshall for sm_100a (K1 in SURVEY.md section 2).
//
// Replaces the per-query Dijkstra of the reference (lib/RoutingEngine.py:87,199,212;
// lib/optimUtils.py:528,531,659) by one all-pairs travel-time matrix.
//
// Layout: row-major matrix, leading dimension ld (multiple of 128), element type
// float (DRS_MODE_F32) or int32 (DRS_MODE_I32).  Round kb of the blocked
// algorithm touches
//   phase 1  the diagonal tile (kb,kb): closed in shared memory by one CTA;
//   phase 2  the pivot row panel  C[kb,j] = D* (x) C[kb,j]  and the pivot column
//            panel C[i,kb] = C[i,kb] (x) D*  (D* = closed diagonal tile);
//   phase 3  every other tile    C[i,j] = min(C[i,j], C[i,kb] (x) C[kb,j]).
// Phases 2 and 3 are the same kernel: a 128x128x128 min-plus tile product whose
// accumulators start from the C tile.  Per CTA: 256 threads, an 8x8 register
// tile per thread, A/B streamed through a 3-stage cp.async ring in 32-wide k
// chunks.  The work is ALU-bound (one add + one min per relaxation, two issue
// slots); HBM traffic per round is one read + one write of the matrix, a factor
// ~5 below the ALU time at tile width 128 (kernel K1).
#include <cstring>

#include "drs_common.cuh"
#include "minplus.cuh"

namespace {

constexpr int TILE = DRS_FW_TILE;  // 128
constexpr int KC = 32;             // k chunk streamed per pipeline stage
constexpr int STAGES = 3;
constexpr int APAD = KC + 4;       // padded row of the A chunk (keeps 16 B alignment)
constexpr int NTHREADS = 256;

using mp::Semiring; using mp::Vec4; using mp::Vec2; using mp::cp_async16; using mp::cp_async4; using mp::cp_async_commit; using mp::cp_async_wait;

// Tile operands are addressed as base + i*si + j*sj (element strides), i = blockIdx.y
// (+ i_off), j = blockIdx.x.
template <typename T> struct TileArgs {
  const T* A; int64_t a_si, a_sj, lda;
  const T* B; int64_t b_si, b_sj, ldb;
  T* C;       int64_t c_si, c_sj, ldc;
  int i_off;   // first block-row index handled by blockIdx.y == 0
  int skip_i;  // block-row to skip (-1: none), compared against i_off + blockIdx.y
  int skip_j;  // block-column to skip (-1: none)
  int n_peers; // fused broadcast: the finished C tile is also stored at the same offset of these buffers
  T* peer[DRS_MAX_PEERS];
};

template <typename T>
__global__ void __launch_bounds__(NTHREADS, 2) fw_minplus_tile_kernel(TileArgs<T> p) {
  using V4 = typename Vec4<T>::type;
  using V2 = typename Vec2<T>::type;
  const int bi = p.i_off + blockIdx.y, bj = blockIdx.x;
  if (bi == p.skip_i || bj == p.skip_j) return;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T(*As)[TILE][APAD] = reinterpret_cast<T(*)[TILE][APAD]>(smem_raw);
  T(*Bs)[KC][TILE] = reinterpret_cast<T(*)[KC][TILE]>(smem_raw + sizeof(T) * STAGES * TILE * APAD);

  const T* __restrict__ A = p.A + bi * p.a_si + bj * p.a_sj;
  const T* __restrict__ B = p.B + bi * p.b_si + bj * p.b_sj;
  T* __restrict__ C = p.C + bi * p.c_si + bj * p.c_sj;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  constexpr bool kPacked = Semiring<T>::packed;

  auto load_chunk = [&](int kc, int stage) {
    const int k0 = kc * KC;
#pragma unroll
    for (int q = 0; q < 4; ++q) {   // A chunk: 128 rows x 32 k
      const int v = tid + NTHREADS * q;
      const int row = v >> 3, cv = v & 7;
      cp_async16(&As[stage][row][cv * 4], A + (int64_t)row * p.lda + k0 + cv * 4);
    }
    if (kPacked) {
      // fp32: B is stored k-pair interleaved, Bs2[k/2][column] = (b[k][column], b[k+1][column]), so that one
      // 8-byte shared load yields the operand pair of a packed add (FADD2) -- 4-byte cp.async, coalesced per row
      float* Bf = reinterpret_cast<float*>(&Bs[stage][0][0]);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int e = tid + NTHREADS * q;
        const int row = e >> 7, col = e & 127;
        cp_async4(Bf + ((row >> 1) * TILE + col) * 2 + (row & 1), B + (int64_t)(k0 + row) * p.ldb + col);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {   // B chunk: 32 k x 128 columns
        const int v = tid + NTHREADS * q;
        const int row = v >> 5, cv = v & 31;
        cp_async16(&Bs[stage][row][cv * 4], B + (int64_t)(k0 + row) * p.ldb + cv * 4);
      }
    }
  };

  constexpr int NCHUNK = TILE / KC;
  load_chunk(0, 0);
  cp_async_commit();
  load_chunk(1, 1);
  cp_async_commit();

  // accumulators start from the C tile: rows ty+16r; columns tx*4.. and 64+tx*4.. (scalar / int path) or the
  // four column pairs 32j + 2tx (packed path: keeps the k-pair-interleaved B loads conflict-free)
  T acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const T* crow = C + (int64_t)(ty + 16 * r) * p.ldc;
    if (kPacked) {
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

#pragma unroll 1
  for (int kc = 0; kc < NCHUNK; ++kc) {
    const int stage = kc % STAGES;
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kc + STAGES - 1 < NCHUNK) load_chunk(kc + STAGES - 1, (kc + STAGES - 1) % STAGES);
    cp_async_commit();

    if (kPacked) {
      // acc = min3(acc, a0+b0, a1+b1) with BOTH sums from one packed add: 1 FADD2 + 1 FMNMX3 per two
      // relaxations = one issue slot per relaxation (the scalar form needs 1.5)
      const float2* Bp = reinterpret_cast<const float2*>(&Bs[stage][0][0]);
#pragma unroll
      for (int kk = 0; kk < KC; kk += 2) {
        float2 a[8], b[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const float2*>(&As[stage][ty + 16 * r][kk]);
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
      V2 a[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const V2*>(&As[stage][ty + 16 * r][kk]);
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
  }
  cp_async_wait<0>();

#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int64_t roff = (int64_t)(ty + 16 * r) * p.ldc;
    if (kPacked) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        V2 v; v.x = acc[r][2 * j]; v.y = acc[r][2 * j + 1];
        *reinterpret_cast<V2*>(C + roff + 32 * j + tx * 2) = v;
        // fused broadcast: P2P stores of the finished tile into every peer's panel buffer (NVLink)
        for (int pr = 0; pr < p.n_peers; ++pr)
          *reinterpret_cast<V2*>(p.peer[pr] + bi * p.c_si + bj * p.c_sj + roff + 32 * j + tx * 2) = v;
      }
    } else {
      V4 lo, hi;
      lo.x = acc[r][0]; lo.y = acc[r][1]; lo.z = acc[r][2]; lo.w = acc[r][3];
      hi.x = acc[r][4]; hi.y = acc[r][5]; hi.z = acc[r][6]; hi.w = acc[r][7];
      *reinterpret_cast<V4*>(C + roff + tx * 4) = lo;
      *reinterpret_cast<V4*>(C + roff + 64 + tx * 4) = hi;
      for (int pr = 0; pr < p.n_peers; ++pr) {
        T* prow = p.peer[pr] + bi * p.c_si + bj * p.c_sj + roff;
        *reinterpret_cast<V4*>(prow + tx * 4) = lo;
        *reinterpret_cast<V4*>(prow + 64 + tx * 4) = hi;
      }
    }
  }
}

constexpr size_t kTileSmemBytes = sizeof(float) * STAGES * (TILE * APAD + KC * TILE);

// Phase 1: close the 128x128 diagonal tile (one CTA; the tile in registers, pivot row / column through shared memory:
// minplus.cuh, fw_close_diag_tile).
template <typename T> struct PeerList { int n; T* ptr[DRS_MAX_PEERS]; };

template <typename T>
__global__ void __launch_bounds__(mp::DIAG_THREADS) fw_diag_kernel(T* D, int64_t ld, PeerList<T> peers) {
  using V4 = typename Vec4<T>::type;
  mp::fw_close_diag_tile<T>(D, ld, threadIdx.x, [&](int64_t off, V4 lo, V4 hi) {
    for (int pr = 0; pr < peers.n; ++pr) {
      *reinterpret_cast<V4*>(peers.ptr[pr] + off) = lo;
      *reinterpret_cast<V4*>(peers.ptr[pr] + off + 4) = hi;
    }
  });
}

// dist[i][j] = (row0+i == j) ? 0 : inf
template <typename T>
__global__ void fw_fill_kernel(T* local, int64_t ld, int row0, int nrows) {
  const int64_t total = (int64_t)nrows * ld;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / ld), j = (int)(idx % ld);
    local[idx] = (row0 + i == j) ? (T)0 : Semiring<T>::inf();
  }
}

// scatter arcs u->v (out CSR) into the rows owned by this shard; multi-edges keep the minimum
template <typename T>
__global__ void fw_scatter_kernel(T* local, int64_t ld, int row0, int nrows, const int32_t* out_ptr,
                                  const int32_t* out_col, const float* out_w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const int u = row0 + i;
  for (int a = out_ptr[u]; a < out_ptr[u + 1]; ++a) {
    const int v = out_col[a];
    if (v == u) continue;
    T w = (T)out_w[a];
    T* cell = local + (int64_t)i * ld + v;
    if (w < *cell) *cell = w;   // one thread per row: no race
  }
}

// Canonical predecessor edge: among in-arcs (u->t, w, eid) minimise
// (dist[s,u] + w, dist[s,u], eid) lexicographically.  One thread per target t keeps t's in-arcs in
// registers and sweeps PRED_ROWS source rows, so the CSR is read once per 32 rows instead of once per
// matrix element; row reads are coalesced over t and the neighbour reads hit L1/L2 (neighbours of
// consecutive vertices are close in a road network).
constexpr int PRED_ROWS = 32;
#ifndef DRS_PRED_THREADS
#define DRS_PRED_THREADS 1024   // a CTA spans 1024 consecutive targets: the +-n_side neighbours of a grid city stay inside its L1
#endif
constexpr int PRED_THREADS = DRS_PRED_THREADS;
constexpr int PRED_MAXDEG = 6;    // in-arcs held in registers; vertices with more fall back to the CSR loop
template <typename T>
__global__ void __launch_bounds__(PRED_THREADS) pred_kernel(const T* __restrict__ local, int64_t ld, int row0, int nrows, int n,
                                                   const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_src,
                                                   const int32_t* __restrict__ in_eid, const float* __restrict__ in_w,
                                                   int32_t* __restrict__ pred) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ld) return;
  const int r_begin = blockIdx.y * PRED_ROWS;
  const int r_end = min(r_begin + PRED_ROWS, nrows);
  int a0 = 0, deg = 0;
  if (t < n) { a0 = in_ptr[t]; deg = in_ptr[t + 1] - a0; }
  int us[PRED_MAXDEG], es[PRED_MAXDEG];
  T ws[PRED_MAXDEG];
  const int dreg = min(deg, PRED_MAXDEG);
#pragma unroll
  for (int k = 0; k < PRED_MAXDEG; ++k) {
    us[k] = 0; es[k] = 0; ws[k] = (T)0;
    if (k < dreg) { us[k] = in_src[a0 + k]; es[k] = in_eid[a0 + k]; ws[k] = (T)in_w[a0 + k]; }
  }
  if (deg <= PRED_MAXDEG) {
    // branch-free body, four rows in flight per thread (the loads of different rows are independent)
#pragma unroll 4
    for (int ls = r_begin; ls < r_end; ++ls) {
      const int s = row0 + ls;
      const T* row = local + (int64_t)ls * ld;
      const T dt = row[t];
      T du[PRED_MAXDEG];
#pragma unroll
      for (int k = 0; k < PRED_MAXDEG; ++k) du[k] = (k < dreg) ? row[us[k]] : Semiring<T>::inf();
      T best_c = Semiring<T>::inf(), best_du = Semiring<T>::inf();
      int best_e = -1;
#pragma unroll
      for (int k = 0; k < PRED_MAXDEG; ++k) {
        const bool fin = (k < dreg) && (du[k] < Semiring<T>::inf());
        const T cand = du[k] + ws[k];
        const bool better = fin && (cand < best_c || (cand == best_c && (du[k] < best_du || (du[k] == best_du && es[k] < best_e))));
        if (better) { best_c = cand; best_du = du[k]; best_e = es[k]; }
      }
      const bool valid = t < n && s < n && t != s && dt < Semiring<T>::inf();
      pred[(int64_t)ls * ld + t] = valid ? best_e : -1;
    }
    return;
  }
  for (int ls = r_begin; ls < r_end; ++ls) {   // high-degree vertices: CSR loop
    const int s = row0 + ls;
    const T* row = local + (int64_t)ls * ld;
    int best_e = -1;
    if (s < n && t != s && row[t] < Semiring<T>::inf()) {
      T best_c = Semiring<T>::inf(), best_du = Semiring<T>::inf();
      for (int a = a0; a < a0 + deg; ++a) {
        const T du = row[in_src[a]];
        if (!(du < Semiring<T>::inf())) continue;
        const T cand = du + (T)in_w[a];
        const int e = in_eid[a];
        if (cand < best_c || (cand == best_c && (du < best_du || (du == best_du && e < best_e)))) {
          best_c = cand; best_du = du; best_e = e;
        }
      }
    }
    pred[(int64_t)ls * ld + t] = best_e;
  }
}

template <typename T>
int launch_tile(const TileArgs<T>& p, dim3 grid, cudaStream_t st) {
  static bool configured[64] = {};   // per device: opt in to >48 KB dynamic shared memory once
  int dev = 0;
  DRS_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    DRS_CUDA(cudaFuncSetAttribute(fw_minplus_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kTileSmemBytes));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  if (grid.x == 0 || grid.y == 0) return DRS_OK;
  fw_minplus_tile_kernel<T><<<grid, NTHREADS, kTileSmemBytes, st>>>(p);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}

template <typename T>
int pivot_row_impl(T* panel, int64_t ld, int kb, cudaStream_t st, int n_peers = 0, void* const* peer_panels = nullptr) {
  const int nb = (int)(ld / TILE);
  T* diag = panel + (int64_t)kb * TILE;
  TileArgs<T> p{};
  PeerList<T> pl{};
  pl.n = n_peers; p.n_peers = n_peers;
  for (int i = 0; i < n_peers; ++i) {
    p.peer[i] = (T*)peer_panels[i];
    pl.ptr[i] = (T*)peer_panels[i] + (int64_t)kb * TILE;
  }
  // make sure attributes are configured before the first diag launch
  p.A = diag; p.a_si = 0; p.a_sj = 0; p.lda = ld;
  p.B = panel; p.b_si = 0; p.b_sj = TILE; p.ldb = ld;
  p.C = panel; p.c_si = 0; p.c_sj = TILE; p.ldc = ld;
  p.i_off = 0; p.skip_i = -1; p.skip_j = kb;
  int rc = launch_tile<T>(p, dim3(0, 0), st);   // configure only
  if (rc) return rc;
  fw_diag_kernel<T><<<1, mp::DIAG_THREADS, 0, st>>>(diag, ld, pl);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  return launch_tile<T>(p, dim3(nb, 1), st);
}

template <typename T>
int update_impl(T* local, int64_t ld, int nrows, const T* panel, int kb, int skip_local, int blo,
                int bhi, cudaStream_t st) {
  const int nb = (int)(ld / TILE);
  if (bhi <= blo) return DRS_OK;
  const T* diag = panel + (int64_t)kb * TILE;
  // phase 2, column part: C[i,kb] = C[i,kb] (x) D*
  TileArgs<T> q{};
  q.A = local + (int64_t)kb * TILE; q.a_si = (int64_t)TILE * ld; q.a_sj = 0; q.lda = ld;
  q.B = diag; q.b_si = 0; q.b_sj = 0; q.ldb = ld;
  q.C = local + (int64_t)kb * TILE; q.c_si = (int64_t)TILE * ld; q.c_sj = 0; q.ldc = ld;
  q.i_off = blo; q.skip_i = skip_local; q.skip_j = -1;
  int rc = launch_tile<T>(q, dim3(1, bhi - blo), st);
  if (rc) return rc;
  // phase 3: C[i,j] = min(C[i,j], C[i,kb] (x) panel[kb,j])
  TileArgs<T> p{};
  p.A = local + (int64_t)kb * TILE; p.a_si = (int64_t)TILE * ld; p.a_sj = 0; p.lda = ld;
  p.B = panel; p.b_si = 0; p.b_sj = TILE; p.ldb = ld;
  p.C = local; p.c_si = (int64_t)TILE * ld; p.c_sj = TILE; p.ldc = ld;
  p.i_off = blo; p.skip_i = skip_local; p.skip_j = kb;
  return launch_tile<T>(p, dim3(nb, bhi - blo), st);
}

}  // namespace

extern "C" int drs_fw_init(const drs_graph* g, int32_t mode, void* local_dev, int64_t ld, int32_t row0,
                           int32_t nrows, void* stream) {
  DRS_REQUIRE(g && local_dev, DRS_ERR_ARG, "drs_fw_init: null argument");
  { int rcm = drs_check_mode(g, mode, "drs_fw_init"); if (rcm) return rcm; }
  DRS_REQUIRE(ld == g->npad && row0 >= 0 && nrows >= 0 && row0 + nrows <= g->npad, DRS_ERR_ARG,
              "drs_fw_init: bad shard (ld=%lld row0=%d nrows=%d npad=%d)", (long long)ld, row0, nrows, g->npad);
  cudaStream_t st = (cudaStream_t)stream;
  if (nrows == 0) return DRS_OK;
  const int nreal = max(0, min(nrows, g->n - row0));   // rows that are real vertices
  if (mode == DRS_MODE_F32) {
    fw_fill_kernel<float><<<148 * 8, 256, 0, st>>>((float*)local_dev, ld, row0, nrows);
    if (nreal > 0)
      fw_scatter_kernel<float><<<(nreal + 127) / 128, 128, 0, st>>>((float*)local_dev, ld, row0, nreal,
                                                                     g->d_out_ptr, g->d_out_col, g->d_out_w);
  } else if (mode == DRS_MODE_I32) {
    fw_fill_kernel<int><<<148 * 8, 256, 0, st>>>((int*)local_dev, ld, row0, nrows);
    if (nreal > 0)
      fw_scatter_kernel<int><<<(nreal + 127) / 128, 128, 0, st>>>((int*)local_dev, ld, row0, nreal,
                                                                   g->d_out_ptr, g->d_out_col, g->d_out_w);
  } else {
    DRS_REQUIRE(false, DRS_ERR_ARG, "drs_fw_init: unknown mode %d", mode);
  }
  drs_count_launch(nreal > 0 ? 2 : 1);
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}

extern "C" int drs_fw_pivot_row(int32_t mode, void* panel_dev, int64_t ld, int32_t kb, void* stream) {
  DRS_REQUIRE(panel_dev && ld > 0 && ld % TILE == 0 && kb >= 0 && kb < ld / TILE, DRS_ERR_ARG,
              "drs_fw_pivot_row: bad argument (ld=%lld kb=%d)", (long long)ld, kb);
  if (mode == DRS_MODE_F32) return pivot_row_impl<float>((float*)panel_dev, ld, kb, (cudaStream_t)stream);
  if (mode == DRS_MODE_I32) return pivot_row_impl<int>((int*)panel_dev, ld, kb, (cudaStream_t)stream);
  DRS_REQUIRE(false, DRS_ERR_ARG, "drs_fw_pivot_row: unknown mode %d", mode);
}

// ---------------------------------------------------------------- fused pivot + P2P broadcast
namespace {
struct WordList { int n; int32_t* w[DRS_MAX_PEERS + 1]; };

__global__ void post_words_kernel(WordList wl, int value) {
  __threadfence_system();   // everything this stream wrote before (the panel tiles) is visible first
  if (threadIdx.x < wl.n) atomicMax_system(wl.w[threadIdx.x], value);   // monotone: posts from different owners may race
  __threadfence_system();
}

__global__ void wait_words_kernel(WordList wl, int value, long long timeout_cycles, int32_t* timed_out) {
  if (threadIdx.x >= wl.n) return;
  const volatile int32_t* w = wl.w[threadIdx.x];
  const long long t0 = clock64();
  while (*w < value) {
    if (clock64() - t0 > timeout_cycles) { *timed_out = 1; break; }
    __nanosleep(200);
  }
  __threadfence_system();
}
}  // namespace

extern "C" int drs_p2p_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle) {
  DRS_REQUIRE(bytes > 0 && dev_ptr && handle, DRS_ERR_ARG, "drs_p2p_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  drs_count_alloc();
  DRS_CUDA(cudaMalloc(dev_ptr, (size_t)bytes));
  DRS_CUDA(cudaMemset(*dev_ptr, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  DRS_CUDA(cudaIpcGetMemHandle(&h, *dev_ptr));
  memcpy(handle, &h, 64);
  return DRS_OK;
}

extern "C" int drs_p2p_open(const unsigned char* handle, void** peer_ptr) {
  DRS_REQUIRE(handle && peer_ptr, DRS_ERR_ARG, "drs_p2p_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  DRS_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DRS_OK;
}

extern "C" int drs_p2p_close(void* peer_ptr) {
  if (peer_ptr) DRS_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return DRS_OK;
}

extern "C" int drs_p2p_free(void* dev_ptr) {
  if (dev_ptr) DRS_CUDA(cudaFree(dev_ptr));
  return DRS_OK;
}

extern "C" int drs_fw_post_words(int32_t n, int32_t* const* words, int32_t value, void* stream) {
  DRS_REQUIRE(n >= 0 && n <= DRS_MAX_PEERS + 1 && (n == 0 || words), DRS_ERR_ARG, "drs_fw_post_words: bad argument");
  if (n == 0) return DRS_OK;
  WordList wl{};
  wl.n = n;
  for (int i = 0; i < n; ++i) wl.w[i] = words[i];
  post_words_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(wl, value);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}

extern "C" int drs_fw_wait_words(int32_t n, const int32_t* const* words, int32_t value, int64_t timeout_ms,
                                 int32_t* timed_out_dev, void* stream) {
  DRS_REQUIRE(n >= 0 && n <= DRS_MAX_PEERS + 1 && (n == 0 || words) && timed_out_dev, DRS_ERR_ARG, "drs_fw_wait_words: bad argument");
  if (n == 0) return DRS_OK;
  WordList wl{};
  wl.n = n;
  for (int i = 0; i < n; ++i) wl.w[i] = const_cast<int32_t*>(words[i]);
  static int khz = 0;          // cudaDevAttrClockRate is a slow query (milliseconds): ask once
  if (!khz) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) != cudaSuccess || khz <= 0) khz = 1965000;
  }
  wait_words_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(wl, value, (long long)timeout_ms * khz, timed_out_dev);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}

extern "C" int drs_fw_pivot_row_p2p(int32_t mode, void* panel_dev, int64_t ld, int32_t kb, int32_t n_peers,
                                    void* const* peer_panels, int32_t* const* peer_signal_words, int32_t signal_value,
                                    void* stream) {
  DRS_REQUIRE(panel_dev && ld > 0 && ld % TILE == 0 && kb >= 0 && kb < ld / TILE, DRS_ERR_ARG,
              "drs_fw_pivot_row_p2p: bad argument (ld=%lld kb=%d)", (long long)ld, kb);
  DRS_REQUIRE(n_peers >= 0 && n_peers <= DRS_MAX_PEERS && (n_peers == 0 || (peer_panels && peer_signal_words)), DRS_ERR_ARG,
              "drs_fw_pivot_row_p2p: at most %d peers", DRS_MAX_PEERS);
  int rc;
  if (mode == DRS_MODE_F32) rc = pivot_row_impl<float>((float*)panel_dev, ld, kb, (cudaStream_t)stream, n_peers, peer_panels);
  else if (mode == DRS_MODE_I32) rc = pivot_row_impl<int>((int*)panel_dev, ld, kb, (cudaStream_t)stream, n_peers, peer_panels);
  else DRS_REQUIRE(false, DRS_ERR_ARG, "drs_fw_pivot_row_p2p: unknown mode %d", mode);
  if (rc) return rc;
  return drs_fw_post_words(n_peers, peer_signal_words, signal_value, stream);
}

extern "C" int drs_fw_update(int32_t mode, void* local_dev, int64_t ld, int32_t nrows, const void* panel_dev,
                             int32_t kb, int32_t skip_local_block, int32_t block_lo, int32_t block_hi,
                             void* stream) {
  DRS_REQUIRE(local_dev && panel_dev && ld > 0 && ld % TILE == 0 && nrows % TILE == 0 && kb >= 0 &&
                  kb < ld / TILE && block_lo >= 0 && block_hi <= nrows / TILE,
              DRS_ERR_ARG, "drs_fw_update: bad argument (ld=%lld nrows=%d kb=%d blocks=[%d,%d))",
              (long long)ld, nrows, kb, block_lo, block_hi);
  if (mode == DRS_MODE_F32)
    return update_impl<float>((float*)local_dev, ld, nrows, (const float*)panel_dev, kb, skip_local_block,
                              block_lo, block_hi, (cudaStream_t)stream);
  if (mode == DRS_MODE_I32)
    return update_impl<int>((int*)local_dev, ld, nrows, (const int*)panel_dev, kb, skip_local_block,
                            block_lo, block_hi, (cudaStream_t)stream);
  DRS_REQUIRE(false, DRS_ERR_ARG, "drs_fw_update: unknown mode %d", mode);
}

extern "C" int drs_pred_build(const drs_graph* g, int32_t mode, const void* local_dev, int64_t ld,
                              int32_t row0, int32_t nrows, int32_t* pred_local_dev, void* stream) {
  DRS_REQUIRE(g && local_dev && pred_local_dev, DRS_ERR_ARG, "drs_pred_build: null argument");
  DRS_REQUIRE(ld == g->npad && row0 >= 0 && nrows >= 0 && row0 + nrows <= g->npad, DRS_ERR_ARG,
              "drs_pred_build: bad shard");
  if (nrows == 0) return DRS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((ld + PRED_THREADS - 1) / PRED_THREADS), (unsigned)((nrows + PRED_ROWS - 1) / PRED_ROWS));
  if (mode == DRS_MODE_F32)
    pred_kernel<float><<<grid, PRED_THREADS, 0, st>>>((const float*)local_dev, ld, row0, nrows, g->n, g->d_in_ptr,
                                             g->d_in_src, g->d_in_eid, g->d_in_w, pred_local_dev);
  else if (mode == DRS_MODE_I32)
    pred_kernel<int><<<grid, PRED_THREADS, 0, st>>>((const int*)local_dev, ld, row0, nrows, g->n, g->d_in_ptr,
                                           g->d_in_src, g->d_in_eid, g->d_in_w, pred_local_dev);
  else
    DRS_REQUIRE(false, DRS_ERR_ARG, "drs_pred_build: unknown mode %d", mode);
  drs_count_launch();
  DRS_CUDA(cudaGetLastError());
  return DRS_OK;
}
