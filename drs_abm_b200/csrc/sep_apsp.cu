// Separator-partitioned min-plus all-pairs build (DRS_APSP_SEP): the blocked Floyd-Warshall of fw_kernels.cu run in
// nested-dissection order, so that a road network's small vertex separators bound the pivots every output tile sees.
//
// Replaces, like the other builders, the per-query Dijkstra of the reference (lib/RoutingEngine.py:87,199,212;
// lib/optimUtils.py:528,531,659) by one resident travel-time matrix; the matrix it writes is the one the SSSP and the plain
// Floyd-Warshall builders write: bit-identical on integer travel times, last-bit differences on float ones (the sums are
// re-associated; DESIGN.md, kernel K0).
//
// The vertices are numbered part-major (drs_partition_order, at the end of this file): the interiors of the parts first,
// one contiguous id range each, and the separator vertices S last.  No arc joins the interiors of two different parts.
// With S_j = the separators adjacent to part j and Q_j = P_j + S_j (the "closed part"):
//   S1  M_j = closure of the subgraph induced by Q_j             batched blocked FW, one small dense matrix per part
//   S2  dS  = closure of the separator graph (arcs inside S and, per part, the S_j x S_j block of M_j): the true
//             distances between separators                        this same build on the separator graph (second level:
//                                                                 run_level recursion on a dense input), batched FW at the top
//   S3  D_uni[u, s] = min_k M_i[u, S_i[k]] + dS[S_i[k], s]        u in P_i: every vertex to every separator; then spread
//             into the per-part k ranges (D_all[u, (j,k)] = D_uni[u, S_j[k]], the A operand of S4) and into the
//             separator columns of the output
//   S4  C[u, v] = min( M_j[u, v] if u in P_j,  min_k D_all[u, (j,k)] + M_j[S_j[k], v] )     v in P_j: the output tiles,
//             each a 128 x 128 x |S_j| min-plus product written exactly once; on undirected maps only the tiles on or
//             above the diagonal, each also stored transposed
// Every tile stage is the 128x128 register-tile product of minplus.cuh: FADD2 + FMNMX3 / VIADDMNMX on 32-bit elements,
// VIADDMNMX.U16x2 (two relaxations per lane) for S3 and S4 when the map's distances stay below 2^15.  Work: n^2 * mean|S_j|
// relaxations for S4 against n^3 for the plain FW and n^2 * (arcs/n) * (re-relaxation factor) dependent L2 round trips for
// the SSSP builder; the matrix leaves the SMs once.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <queue>
#include <thread>
#include <type_traits>

#include "drs_common.cuh"
#include "minplus.cuh"

using namespace mp;

struct drs_sep_plan {
  int nparts = 0, sep0 = 0, nsep = 0, spad = 0, ktot = 0, kmax = 0, maxnb = 0, max_rowtiles = 0, ncoltiles = 0, maxq = 0;
  std::vector<int32_t> pstart, pend, sj, kpad, koff, qpad, cbase, wpad, sall, sepcol, ct_part, ct_col0;
  std::vector<int64_t> moff, apoff;
  int64_t m_elems = 0, ap_elems = 0;
  // device: one arena for the tables, one for the work arrays
  void* d_tables = nullptr;
  size_t tables_bytes = 0;
  void* d_work = nullptr;
  size_t work_bytes = 0;
  const int32_t *d_pstart = nullptr, *d_pend = nullptr, *d_sj = nullptr, *d_kpad = nullptr, *d_koff = nullptr, *d_qpad = nullptr,
                *d_cbase = nullptr, *d_wpad = nullptr, *d_sall = nullptr, *d_sepcol = nullptr, *d_ct_part = nullptr,
                *d_ct_col0 = nullptr;
  const int64_t *d_moff = nullptr, *d_apoff = nullptr, *d_top_moff = nullptr;
  const int32_t* d_top_qpad = nullptr;   // the separator graph's matrix as a one-part table for the batched FW kernels
  void *d_M = nullptr, *d_dS = nullptr, *d_Dall = nullptr, *d_dSr = nullptr, *d_Duni = nullptr, *d_Aprime = nullptr, *d_Apad = nullptr;
  void* d_Apad16 = nullptr;      // the A_j panels as 16-bit distances (the B operand of the 16-bit output tiles)
  void* d_Aprime16 = nullptr;    // A' with the 16-bit distance in both halves of a word (A operand of the 16-bit rows stage)
  void* d_dSr16 = nullptr;       // the gathered separator rows as 16-bit distances (its B operand)
  int32_t* d_flags = nullptr;    // [0] largest finite separator-to-separator distance, [1] largest finite distance inside a closed part (float bits)
  bool used16 = false;           // the last build ran S3 and S4 on the 16-bit path
  bool no16 = false;             // DRS_SEP_NO_U16=1 (measurement knob)
  int tile_rows = 32;            // rows of a 16-bit output tile: 32 (default), DRS_SEP_TILE_ROWS=64 or 128
  double prof_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool no_mirror = false;   // DRS_SEP_NO_MIRROR=1: compute every output tile (measurement knob)
  int n = 0, npad = 0;      // vertices of this level and their padded count (the leading dimension of its matrix)
  drs_sep_plan* child = nullptr;   // level 2: the same build on this level's separator graph
};

namespace {

struct SepTab {
  int nparts, sep0, nsep, n, npad, spad, ktot, kmax;
  const int32_t *pstart, *pend, *sj, *kpad, *koff, *qpad, *cbase, *wpad;
  const int64_t *moff, *apoff;
  const int32_t *sall, *sepcol, *ct_part, *ct_col0;
};

template <typename T> __device__ __forceinline__ void atomic_min_nonneg(T* addr, T v);
template <> __device__ __forceinline__ void atomic_min_nonneg<float>(float* addr, float v) {
  atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));   // non-negative floats (and +inf) order like their bits
}
template <> __device__ __forceinline__ void atomic_min_nonneg<int>(int* addr, int v) { atomicMin(addr, v); }

// ------------------------------------------------------------------------------------------------ S1: closed parts
template <typename T>
__global__ void sep_fill_parts_kernel(T* M, SepTab t) {
  const int j = blockIdx.y;
  const int ld = t.qpad[j];
  T* Mj = M + t.moff[j];
  for (int i = blockIdx.x; i < ld; i += gridDim.x)
    for (int c = threadIdx.x; c < ld; c += blockDim.x) Mj[(int64_t)i * ld + c] = (i == c) ? (T)0 : Semiring<T>::inf();
}

// local index of vertex y in Q_j (interior first, then S_j in list order), -1 when y is outside the closed part
__device__ __forceinline__ int local_of(const SepTab& t, int j, int y) {
  const int ps = t.pstart[j], pe = t.pend[j];
  if (y >= ps && y < pe) return y - ps;
  if (y < t.sep0) return -1;
  const int key = y - t.sep0;
  const int32_t* lst = t.sall + t.koff[j];
  int lo = 0, hi = t.sj[j];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (lst[mid] < key) lo = mid + 1; else hi = mid;
  }
  return (lo < t.sj[j] && lst[lo] == key) ? (pe - ps) + lo : -1;
}

template <typename T>
__global__ void sep_scatter_parts_kernel(T* M, SepTab t, const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_col,
                                         const float* __restrict__ out_w) {
  const int j = blockIdx.y;
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  const int nj = t.pend[j] - t.pstart[j];
  if (li >= nj + t.sj[j]) return;
  const int x = li < nj ? t.pstart[j] + li : t.sep0 + t.sall[t.koff[j] + (li - nj)];
  const int64_t ld = t.qpad[j];
  T* row = M + t.moff[j] + (int64_t)li * ld;
  for (int a = out_ptr[x]; a < out_ptr[x + 1]; ++a) {
    const int y = out_col[a];
    if (y == x) continue;
    const int ly = local_of(t, j, y);
    if (ly < 0) continue;
    const T w = (T)out_w[a];
    if (w < row[ly]) row[ly] = w;   // one thread per row: no race; multi-edges keep the minimum
  }
}

// round r of every part's blocked FW, phase 1: close the diagonal tile (r, r) (minplus.cuh: fw_close_diag_tile)
template <typename T>
__global__ void __launch_bounds__(DIAG_THREADS) sep_fw_diag_kernel(T* M, SepTab t, int r) {
  using V4 = typename Vec4<T>::type;
  const int j = blockIdx.x;
  const int64_t ld = t.qpad[j];
  if ((int64_t)r * TILE >= ld) return;
  fw_close_diag_tile<T>(M + t.moff[j] + (int64_t)r * TILE * (ld + 1), ld, threadIdx.x, [](int64_t, V4, V4) {});
}

// phases 2 (pivot row / column panels: grid (maxnb, 2, nparts)) and 3 (every other tile: grid (maxnb, maxnb, nparts))
template <typename T>
__global__ void __launch_bounds__(NTHREADS, 2) sep_fw_tile_kernel(T* M, SepTab t, int r, int phase) {
  const int j = blockIdx.z;
  const int64_t ld = t.qpad[j];
  const int nb = (int)(ld / TILE);
  if (r >= nb) return;
  T* Mj = M + t.moff[j];
  const int bx = blockIdx.x, by = blockIdx.y;
  const T *A, *B;
  T* C;
  if (phase == 2) {
    if (bx >= nb || bx == r) return;
    const T* diag = Mj + (int64_t)r * TILE * (ld + 1);
    if (by == 0) { C = Mj + (int64_t)r * TILE * ld + (int64_t)bx * TILE; A = diag; B = C; }
    else { C = Mj + (int64_t)bx * TILE * ld + (int64_t)r * TILE; A = C; B = diag; }
  } else {
    if (bx >= nb || by >= nb || bx == r || by == r) return;
    C = Mj + (int64_t)by * TILE * ld + (int64_t)bx * TILE;
    A = Mj + (int64_t)by * TILE * ld + (int64_t)r * TILE;
    B = Mj + (int64_t)r * TILE * ld + (int64_t)bx * TILE;
  }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  T acc[8][8];
  acc_load<T>(acc, C, ld, tid);
  tile_accumulate<T>(acc, A, ld, B, ld, TILE, smem_raw, tid);
  acc_store<T>(acc, C, ld, TILE, TILE, tid);
}

// S1x: what the later stages read from the closed parts.  blockIdx.z: 0 = A' rows (interior -> own separators),
// 1 = padded A_j panels (own separators -> interior, columns on the global 128-column grid), 2 = S_j x S_j into dS
template <typename T> __device__ __forceinline__ int finite_bits(T v);   // ordering key of a finite distance, 0 for +inf
template <> __device__ __forceinline__ int finite_bits<float>(float v) { return v < Semiring<float>::inf() ? __float_as_int(v) : 0; }
template <> __device__ __forceinline__ int finite_bits<int>(int v) { return v < Semiring<int>::inf() ? __float_as_int((float)v) : 0; }
__device__ __forceinline__ void flag_max(int32_t* flag, int key) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, o));
  if ((threadIdx.x & 31) == 0 && key > 0) atomicMax(flag, key);
}

template <typename T>
__global__ void sep_extract_kernel(const T* __restrict__ M, SepTab t, T* __restrict__ Aprime, T* __restrict__ Apad, T* dS,
                                   uint16_t* __restrict__ Apad16, uint32_t* __restrict__ Aprime16, int32_t* flags) {
  const int j = blockIdx.y;
  const int ps = t.pstart[j], nj = t.pend[j] - ps, sj = t.sj[j], kp = t.kpad[j];
  const int64_t ld = t.qpad[j];
  const T* Mj = M + t.moff[j];
  const unsigned g0 = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;   // every extent here is far below 2^31
  if (blockIdx.z == 0) {
    const unsigned total = (unsigned)nj * t.kmax, kmax = t.kmax;
    int key = 0;
    for (unsigned idx = g0; idx < total; idx += gs) {
      const int li = (int)(idx / kmax), k = (int)(idx - (unsigned)li * kmax);
      const T v = k < sj ? Mj[(int64_t)li * ld + nj + k] : Semiring<T>::inf();
      Aprime[(int64_t)(ps + li) * t.kmax + k] = v;
      Aprime16[(int64_t)(ps + li) * t.kmax + k] = dist_to_u16((float)v) * 0x10001u;
      key = max(key, finite_bits<T>(v));
    }
    flag_max(flags + 1, key);
  } else if (blockIdx.z == 1) {
    const int wp = t.wpad[j], cb = t.cbase[j];
    T* Aj = Apad + t.apoff[j];
    uint16_t* Aj16 = Apad16 + t.apoff[j];
    const unsigned total = (unsigned)((kp + KC - 1) / KC * KC) * wp;
    int key = 0;
    for (unsigned idx = g0; idx < total; idx += gs) {
      const int k = (int)(idx / (unsigned)wp), c = (int)(idx - (unsigned)k * wp);
      const int lv = cb + c - ps;
      const T v = (k < sj && lv >= 0 && lv < nj) ? Mj[(int64_t)(nj + k) * ld + lv] : Semiring<T>::inf();
      Aj[idx] = v;
      Aj16[idx] = (uint16_t)dist_to_u16((float)v);
      key = max(key, finite_bits<T>(v));
    }
    // distances inside the part (the other candidates of its diagonal tiles) count towards the same bound
    const unsigned inner = (unsigned)nj * nj;
    for (unsigned idx = g0; idx < inner; idx += gs) {
      const int li = (int)(idx / (unsigned)nj), lj = (int)(idx - (unsigned)li * nj);
      key = max(key, finite_bits<T>(Mj[(int64_t)li * ld + lj]));
    }
    flag_max(flags + 1, key);
  } else {
    const int32_t* lst = t.sall + t.koff[j];
    const unsigned total = (unsigned)sj * sj;
    for (unsigned idx = g0; idx < total; idx += gs) {
      const int ka = (int)(idx / (unsigned)sj), kb = (int)(idx - (unsigned)ka * sj);
      const T v = Mj[(int64_t)(nj + ka) * ld + nj + kb];
      if (v < Semiring<T>::inf()) atomic_min_nonneg<T>(dS + (int64_t)lst[ka] * t.spad + lst[kb], v);
    }
  }
}

// ------------------------------------------------------------------------------------------------ S2: separator graph
template <typename T>
__global__ void sep_fill_identity_kernel(T* M, int64_t ld, int64_t rows) {
  for (int i = blockIdx.x; i < rows; i += gridDim.x)
    for (int c = threadIdx.x; c < ld; c += blockDim.x) M[(int64_t)i * ld + c] = (i == c) ? (T)0 : Semiring<T>::inf();
}

template <typename T>
__global__ void sep_arcs_kernel(T* dS, SepTab t, const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_col,
                                const float* __restrict__ out_w) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= t.nsep) return;
  const int x = t.sep0 + r;
  for (int a = out_ptr[x]; a < out_ptr[x + 1]; ++a) {
    const int y = out_col[a];
    if (y < t.sep0 || y == x) continue;
    atomic_min_nonneg<T>(dS + (int64_t)r * t.spad + (y - t.sep0), (T)out_w[a]);
  }
}

// S2x: the rows of dS in the order of the concatenated separator lists (the B operand of S3), and the separator rows of
// D_uni (a separator's distances to the separators are its dS row); padding rows = +inf
template <typename T>
__global__ void sep_gather_dsr_kernel(const T* __restrict__ dS, SepTab t, T* __restrict__ dSr, T* __restrict__ Duni,
                                      uint16_t* __restrict__ dSr16, int32_t* flags) {
  const int rows1 = t.ktot, rows2 = t.npad - t.sep0;
  int key = 0;
  for (int row = blockIdx.x; row < rows1 + rows2; row += gridDim.x) {
    if (row < rows1) {
      const int s = t.sall[row];
      const T* src = dS + (int64_t)max(s, 0) * t.spad;
      T* dst = dSr + (int64_t)row * t.spad;
      uint16_t* dst16 = dSr16 + (int64_t)row * t.spad;
      for (int c = threadIdx.x; c < t.spad; c += blockDim.x) {
        const T v = s >= 0 ? src[c] : Semiring<T>::inf();
        dst[c] = v;
        dst16[c] = (uint16_t)dist_to_u16((float)v);
      }
    } else {
      const int r = row - rows1;
      const T* src = dS + (int64_t)min(r, t.spad - 1) * t.spad;
      T* dst = Duni + (int64_t)(t.sep0 + r) * t.spad;
      for (int c = threadIdx.x; c < t.spad; c += blockDim.x) {
        const T v = r < t.nsep ? src[c] : Semiring<T>::inf();
        dst[c] = v;
        key = max(key, finite_bits<T>(v));
      }
    }
  }
  flag_max(flags, key);
}

// ------------------------------------------------------------------------------------------------ S3: vertex -> separators
// D_uni[u, s] = min_k A'[u, k] + dS[S_i[k], s] for the interior vertices u of part i: grid (spad / 128, max row tiles of a
// part, nparts).  Row tiles start at the part's first vertex; rows past its last one are computed from the next part's A'
// rows and not stored.
constexpr int XROWS = 64, XSTG = 2;   // 32-bit expansion tiles: 64 rows (128 threads), two-stage ring: four CTAs per SM
using XCfg = TileCfg<XROWS, XSTG>;
template <typename T>
__global__ void __launch_bounds__(XCfg::NT, 4) sep_expand_rows_kernel(const T* __restrict__ Aprime, const T* __restrict__ dSr,
                                                                        T* __restrict__ Duni, SepTab t, int row_lo, int row_hi) {
  const int i = blockIdx.z;
  const int row0 = t.pstart[i] + blockIdx.y * XROWS, pe = t.pend[i];
  if (row0 >= pe || row0 >= row_hi || min(row0 + XROWS, pe) <= row_lo) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  T acc[8][8];
  acc_fill<T>(acc, Semiring<T>::inf());
  tile_accumulate<T, XROWS, XSTG>(acc, Aprime + (int64_t)row0 * t.kmax, t.kmax,
                                  dSr + (int64_t)t.koff[i] * t.spad + (int64_t)blockIdx.x * TILE, t.spad, t.kpad[i], smem_raw, tid);
  acc_store<T, XROWS>(acc, Duni + (int64_t)row0 * t.spad + (int64_t)blockIdx.x * TILE, t.spad, pe - row0, TILE, tid);
}

// the same tiles on the 16-bit path (fp32 D_uni out), ROWS rows high (grid.y = row tiles of that height)
template <int ROWS>
__global__ void __launch_bounds__(2 * ROWS, ROWS == TILE ? 3 : ROWS == 64 ? 6 : 8)
sep_expand_rows16_kernel(const uint32_t* __restrict__ Aprime16, const uint16_t* __restrict__ dSr16, float* __restrict__ Duni, SepTab t,
                         int row_lo, int row_hi) {
  const int i = blockIdx.z;
  const int row0 = t.pstart[i] + blockIdx.y * ROWS, pe = t.pend[i];
  if (row0 >= pe || row0 >= row_hi || min(row0 + ROWS, pe) <= row_lo) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  uint32_t acc2[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc2[r][p] = U16_INF * 0x10001u;
  tile_accumulate_u16<ROWS>(acc2, Aprime16 + (int64_t)row0 * t.kmax, t.kmax, dSr16 + (int64_t)t.koff[i] * t.spad + (int64_t)blockIdx.x * TILE,
                            t.spad, t.kpad[i], smem_raw, tid);
  acc16_store<ROWS>(acc2, Duni + (int64_t)row0 * t.spad + (int64_t)blockIdx.x * TILE, t.spad, pe - row0, TILE, tid, 0);
}

// S3x: SPREAD_ROWS rows of the shard per CTA: the rows of D_uni spread over the per-part k ranges (D_all, the A operand
// of S4) and copied into the separator columns of the output; padding as the other builders leave it (0 on the diagonal
// of the padding rows, +inf elsewhere)
constexpr int SPREAD_ROWS = 4;   // 4 rows of D_uni (88 KB on the 50k map) stay in L1 while they are gathered ~3 times
template <typename T, bool U16>   // U16: D_all holds the 16-bit distance in both halves of a 32-bit word (A operand of the 16-bit tiles)
__global__ void sep_spread_kernel(const T* __restrict__ Duni, T* __restrict__ Dall, T* __restrict__ C, int64_t ldc, SepTab t, int row_lo) {
  const int lr0 = blockIdx.x * SPREAD_ROWS, u0 = row_lo + lr0;
  for (int c = threadIdx.x; c < t.ktot; c += blockDim.x) {
    const int s = t.sall[c];
#pragma unroll
    for (int r = 0; r < SPREAD_ROWS; ++r) {
      const int u = u0 + r;
      const T v = (s >= 0 && u < t.n) ? Duni[(int64_t)u * t.spad + s] : Semiring<T>::inf();
      if (U16) reinterpret_cast<uint32_t*>(Dall)[(int64_t)u * t.ktot + c] = dist_to_u16((float)v) * 0x10001u;
      else Dall[(int64_t)u * t.ktot + c] = v;
    }
  }
  for (int c = t.sep0 + threadIdx.x; c < t.npad; c += blockDim.x) {
#pragma unroll
    for (int r = 0; r < SPREAD_ROWS; ++r) {
      const int u = u0 + r;
      T v;
      if (u >= t.n || c >= t.n) v = (u == c) ? (T)0 : Semiring<T>::inf();
      else v = Duni[(int64_t)u * t.spad + (c - t.sep0)];
      C[(int64_t)(lr0 + r) * ldc + c] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ S4: the output tiles
// grid (column tiles, row tiles of the shard).  Column tiles are laid out per part: tile t of part j covers the columns
// cbase_j + 128 t .. of the part (cbase_j = the part's first vertex rounded down to 4, so vector stores stay aligned); a
// part's padded A_j panel is +inf outside its own columns and the store is clipped to them, so every element of the
// matrix is written by exactly one tile.  MIRROR (undirected maps, whole-matrix builds): only the tiles that reach the
// diagonal or lie above it are computed, and each also writes its transpose -- the matrix is symmetric.
template <typename T, bool MIRROR>
__global__ void __launch_bounds__(XCfg::NT, 4) sep_expand_kernel(const T* __restrict__ Dall, const T* __restrict__ Apad,
                                                                   const T* __restrict__ M, T* __restrict__ C, int64_t ldc, SepTab t,
                                                                   int row_lo) {
  const int ct = blockIdx.x;
  const int j = t.ct_part[ct], C0 = t.ct_col0[ct];
  const int R0 = row_lo + blockIdx.y * XROWS;
  const int ps = t.pstart[j], pe = t.pend[j];
  const int col_lo = max(ps - C0, 0), col_hi = min(pe - C0, TILE);
  if (MIRROR && !(C0 + col_hi > R0 || R0 + XROWS > t.sep0)) return;   // strictly below the diagonal, interior rows: mirrored
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  T acc[8][8];
  acc_fill<T>(acc, Semiring<T>::inf());
  if (R0 < pe && R0 + XROWS > ps) {   // rows of this tile lie in part j: paths that stay inside the closed part
    const int64_t ld = t.qpad[j];
    const T* Mj = M + t.moff[j];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int u = R0 + acc_row<T, XROWS>(r, ty);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int v = C0 + acc_col<T>(c, tx);
        if (u >= ps && u < pe && v >= ps && v < pe) acc[r][c] = Mj[(int64_t)(u - ps) * ld + (v - ps)];
      }
    }
  }
  tile_accumulate<T, XROWS, XSTG>(acc, Dall + (int64_t)R0 * t.ktot + t.koff[j], t.ktot, Apad + t.apoff[j] + (C0 - t.cbase[j]), t.wpad[j],
                                  t.kpad[j], smem_raw, tid);
  acc_store<T, XROWS>(acc, C + (int64_t)(R0 - row_lo) * ldc + C0, ldc, XROWS, col_hi, tid, col_lo);
  if (MIRROR && R0 < t.sep0)   // C[v][u] = C[u][v]; columns u are interior vertices only (separator columns: S4x)
    acc_store_transposed<T, XROWS>(acc, smem_raw, C + (int64_t)C0 * ldc + R0, ldc, col_lo, col_hi, min(XROWS, t.sep0 - R0), tid);
}

// S4 on the 16-bit path (fp32 matrices of maps whose distances stay below 2^15): same tiles, VIADDMNMX.U16x2 inside.
// ROWS = 128 (three CTAs of 256 threads per SM) or 64 (six CTAs of 128 threads: minplus.cuh, Tile16)
template <bool MIRROR, int ROWS>
__global__ void __launch_bounds__(2 * ROWS, ROWS == TILE ? 3 : ROWS == 64 ? 6 : 8)
sep_expand16_kernel(const uint32_t* __restrict__ Dall2, const uint16_t* __restrict__ Apad16, const float* __restrict__ M,
                    float* __restrict__ C, int64_t ldc, SepTab t, int row_lo) {
  constexpr int RG = Tile16<ROWS>::RG;
  const int ct = blockIdx.x;
  const int j = t.ct_part[ct], C0 = t.ct_col0[ct];
  const int R0 = row_lo + blockIdx.y * ROWS;
  const int ps = t.pstart[j], pe = t.pend[j];
  const int col_lo = max(ps - C0, 0), col_hi = min(pe - C0, TILE);
  if (MIRROR && !(C0 + col_hi > R0 || R0 + ROWS > t.sep0)) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  uint32_t acc2[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc2[r][p] = U16_INF * 0x10001u;
  if (R0 < pe && R0 + ROWS > ps) {
    const int64_t ld = t.qpad[j];
    const float* Mj = M + t.moff[j];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int u = R0 + ty + RG * r;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int v = C0 + tx * 8 + q;
        if (u >= ps && u < pe && v >= ps && v < pe) {
          const unsigned d = dist_to_u16(Mj[(int64_t)(u - ps) * ld + (v - ps)]);
          acc2[r][q >> 1] = (q & 1) ? ((acc2[r][q >> 1] & 0xffffu) | (d << 16)) : ((acc2[r][q >> 1] & 0xffff0000u) | d);
        }
      }
    }
  }
  tile_accumulate_u16<ROWS>(acc2, Dall2 + (int64_t)R0 * t.ktot + t.koff[j], t.ktot, Apad16 + t.apoff[j] + (C0 - t.cbase[j]), t.wpad[j],
                            t.kpad[j], smem_raw, tid);
  acc16_store<ROWS>(acc2, C + (int64_t)(R0 - row_lo) * ldc + C0, ldc, ROWS, col_hi, tid, col_lo);
  if (MIRROR && R0 < t.sep0)
    acc16_store_transposed<ROWS>(acc2, smem_raw, C + (int64_t)C0 * ldc + R0, ldc, col_lo, col_hi, min(ROWS, t.sep0 - R0), tid);
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Level 2 (the separator graph of level 1 handled by the same build): the closed parts and the top separator graph
// start from a dense matrix W (the level-1 separator graph: arcs inside S and the S_j x S_j blocks) instead of a CSR
template <typename T>
__global__ void sep_gather_parts_kernel(T* M, SepTab t, const T* __restrict__ W, int64_t ldw) {
  const int j = blockIdx.y;
  const int ps = t.pstart[j], nj = t.pend[j] - ps, q = nj + t.sj[j];
  const int ld = t.qpad[j];
  const int32_t* lst = t.sall + t.koff[j];
  T* Mj = M + t.moff[j];
  for (int li = blockIdx.x; li < ld; li += gridDim.x) {
    const int x = li < nj ? ps + li : (li < q ? t.sep0 + lst[li - nj] : 0);
    for (int lj = threadIdx.x; lj < ld; lj += blockDim.x) {
      T v = (li == lj) ? (T)0 : Semiring<T>::inf();
      if (li < q && lj < q && li != lj) {
        const int y = lj < nj ? ps + lj : t.sep0 + lst[lj - nj];
        v = W[(int64_t)x * ldw + y];
      }
      Mj[(int64_t)li * ld + lj] = v;
    }
  }
}

template <typename T>
__global__ void sep_gather_top_kernel(T* dS, SepTab t, const T* __restrict__ W, int64_t ldw) {
  for (int a = blockIdx.x; a < t.spad; a += gridDim.x)
    for (int b = threadIdx.x; b < t.spad; b += blockDim.x) {
      T v = (a == b) ? (T)0 : Semiring<T>::inf();
      if (a < t.nsep && b < t.nsep && a != b) v = W[(int64_t)(t.sep0 + a) * ldw + (t.sep0 + b)];
      dS[(int64_t)a * t.spad + b] = v;
    }
}

SepTab device_tab(const drs_sep_plan* p) {
  SepTab t{};
  t.nparts = p->nparts; t.sep0 = p->sep0; t.nsep = p->nsep; t.n = p->n; t.npad = p->npad; t.spad = p->spad; t.ktot = p->ktot;
  t.kmax = p->kmax;
  t.pstart = p->d_pstart; t.pend = p->d_pend; t.sj = p->d_sj; t.kpad = p->d_kpad; t.koff = p->d_koff; t.qpad = p->d_qpad;
  t.cbase = p->d_cbase; t.wpad = p->d_wpad; t.moff = p->d_moff; t.apoff = p->d_apoff; t.sall = p->d_sall; t.sepcol = p->d_sepcol;
  t.ct_part = p->d_ct_part; t.ct_col0 = p->d_ct_col0;
  return t;
}

template <typename T>
int configure_kernels() {
  static bool configured[64] = {};
  int dev = 0;
  DRS_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && configured[dev]) return DRS_OK;
  DRS_CUDA(cudaFuncSetAttribute(sep_fw_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes));
  DRS_CUDA(cudaFuncSetAttribute(sep_expand_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XCfg::smem_bytes));
  DRS_CUDA(cudaFuncSetAttribute(sep_expand_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XCfg::smem_bytes));
  DRS_CUDA(cudaFuncSetAttribute(sep_expand_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XCfg::smem_bytes));
  DRS_CUDA(cudaFuncSetAttribute(sep_expand_rows16_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile16<64>::smem_bytes));
  DRS_CUDA(cudaFuncSetAttribute(sep_expand16_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile16<128>::smem_bytes));
  DRS_CUDA(cudaFuncSetAttribute(sep_expand16_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile16<128>::smem_bytes));
  DRS_CUDA(cudaFuncSetAttribute(sep_expand16_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile16<64>::smem_bytes));
  DRS_CUDA(cudaFuncSetAttribute(sep_expand16_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile16<64>::smem_bytes));
  DRS_CUDA(cudaFuncSetAttribute(sep_expand16_kernel<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile16<32>::smem_bytes));
  DRS_CUDA(cudaFuncSetAttribute(sep_expand16_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile16<32>::smem_bytes));
  if (dev >= 0 && dev < 64) configured[dev] = true;
  return DRS_OK;
}

int ensure_work(drs_sep_plan* p, cudaStream_t st) {
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t b_M = al(4 * (size_t)p->m_elems), b_dS = al(4 * (size_t)p->spad * p->spad);
  const size_t b_Dall = al(4 * (size_t)p->npad * p->ktot), b_dSr = al(4 * (size_t)p->ktot * p->spad);
  const size_t b_Duni = al(4 * (size_t)p->npad * p->spad);
  const size_t b_Apr = al(4 * (size_t)(p->npad + TILE) * p->kmax), b_Apad = al(4 * (size_t)std::max<int64_t>(p->ap_elems, 1));
  const size_t b_Apad16 = al(2 * (size_t)std::max<int64_t>(p->ap_elems, 1)), b_flags = 256;
  const size_t b_dSr16 = al(2 * (size_t)p->ktot * p->spad);
  const size_t total = b_M + b_dS + b_Dall + b_dSr + b_Duni + 2 * b_Apr + b_Apad + b_Apad16 + b_dSr16 + b_flags;
  const bool fresh = !p->d_work || p->work_bytes < total;
  if (fresh) {
    if (p->d_work) drs_big_free(p->d_work, p->work_bytes);
    p->d_work = nullptr; p->work_bytes = 0;
    { int rca = drs_big_alloc(&p->d_work, total); if (rca) return rca; }
    p->work_bytes = total;
  }
  char* b = (char*)p->d_work;
  p->d_M = b; b += b_M; p->d_dS = b; b += b_dS; p->d_Dall = b; b += b_Dall; p->d_dSr = b; b += b_dSr; p->d_Duni = b; b += b_Duni;
  p->d_Aprime = b; b += b_Apr; p->d_Aprime16 = b; b += b_Apr; p->d_Apad = b; b += b_Apad; p->d_Apad16 = b; b += b_Apad16;
  p->d_dSr16 = b; b += b_dSr16; p->d_flags = (int32_t*)b;
  // rows of A' past the last interior vertex are read by the last row tile of a part; what is computed from them is never
  // stored, but they should hold numbers that stay numbers
  if (fresh) DRS_CUDA(cudaMemsetAsync(p->d_Aprime, 0x7f, 2 * b_Apr, st));   // A' and its 16-bit twin are adjacent
  return DRS_OK;
}

// where a level's arcs come from: the device CSR of the map (level 1) or a dense matrix (level 2)
struct LevelInput {
  const int32_t* out_ptr = nullptr; const int32_t* out_col = nullptr; const float* out_w = nullptr;
  const void* W = nullptr; int64_t ldw = 0;
};

// Rows [row0, row0 + nrows) of the level's all-pairs matrix into C (nrows x ldc).  Launches only: the caller
// synchronises.  ev (9 events, may be null) brackets the stages of this level.
template <typename T>
int run_level(drs_sep_plan* p, const LevelInput& in, int mode, bool directed, T* C, int64_t ldc, int row0, int nrows, cudaStream_t st,
              cudaEvent_t* ev, bool allow16) {
  int rc;
  DrsHostTimer tm("sep run_level");
  if ((rc = ensure_work(p, st))) return rc;
  tm.lap("work arrays");
  const SepTab t = device_tab(p);
  T *M = (T*)p->d_M, *dS = (T*)p->d_dS, *Dall = (T*)p->d_Dall, *dSr = (T*)p->d_dSr, *Duni = (T*)p->d_Duni, *Apr = (T*)p->d_Aprime,
    *Apad = (T*)p->d_Apad;
#define SEP_CHECK() DRS_REQUIRE(cudaGetLastError() == cudaSuccess, DRS_ERR_CUDA, "separator build: kernel launch failed (%s:%d)", __FILE__, __LINE__)
#define SEP_EVENT(i) do { if (ev) cudaEventRecord(ev[i], st); } while (0)
  DRS_CUDA(cudaMemsetAsync(p->d_flags, 0, 8, st));
  SEP_EVENT(0);
  // S1: closed parts
  if (in.W) {
    sep_gather_parts_kernel<T><<<dim3(64, p->nparts), 256, 0, st>>>(M, t, (const T*)in.W, in.ldw);
    drs_count_launch();
  } else {
    sep_fill_parts_kernel<T><<<dim3(64, p->nparts), 256, 0, st>>>(M, t);
    sep_scatter_parts_kernel<T><<<dim3((p->maxq + 127) / 128, p->nparts), 128, 0, st>>>(M, t, in.out_ptr, in.out_col, in.out_w);
    drs_count_launch(2);
  }
  SEP_CHECK();
  SEP_EVENT(1);
  for (int r = 0; r < p->maxnb; ++r) {
    sep_fw_diag_kernel<T><<<p->nparts, DIAG_THREADS, 0, st>>>(M, t, r);
    drs_count_launch();
    if (p->maxnb > 1) {
      sep_fw_tile_kernel<T><<<dim3(p->maxnb, 2, p->nparts), NTHREADS, kTileSmemBytes, st>>>(M, t, r, 2);
      sep_fw_tile_kernel<T><<<dim3(p->maxnb, p->maxnb, p->nparts), NTHREADS, kTileSmemBytes, st>>>(M, t, r, 3);
      drs_count_launch(2);
    }
    SEP_CHECK();
  }
  SEP_EVENT(2);
  // S1x: the separator graph's matrix, A' and the padded A_j panels
  if (in.W) sep_gather_top_kernel<T><<<148 * 8, 256, 0, st>>>(dS, t, (const T*)in.W, in.ldw);
  else sep_fill_identity_kernel<T><<<148 * 8, 256, 0, st>>>(dS, p->spad, p->spad);
  sep_extract_kernel<T><<<dim3(32, p->nparts, 3), 256, 0, st>>>(M, t, Apr, Apad, dS, (uint16_t*)p->d_Apad16, (uint32_t*)p->d_Aprime16, p->d_flags);
  drs_count_launch(2);
  if (p->nsep > 0 && !in.W) {
    sep_arcs_kernel<T><<<(p->nsep + 127) / 128, 128, 0, st>>>(dS, t, in.out_ptr, in.out_col, in.out_w);
    drs_count_launch();
  }
  SEP_CHECK();
  SEP_EVENT(3);
  // S2: closure of the separator graph -- by this same build one level up when the partition has a second level
  if (p->nsep > 0) {
    if (p->child) {
      LevelInput up;
      up.W = dS; up.ldw = p->spad;
      if ((rc = run_level<T>(p->child, up, mode, directed, dS, p->spad, 0, p->spad, st, nullptr, false))) return rc;   // in place
    } else {
      const int nbs = p->spad / TILE;
      SepTab top = t;
      top.moff = p->d_top_moff; top.qpad = p->d_top_qpad;
      for (int r = 0; r < nbs; ++r) {
        sep_fw_diag_kernel<T><<<1, DIAG_THREADS, 0, st>>>(dS, top, r);
        drs_count_launch();
        if (nbs > 1) {
          sep_fw_tile_kernel<T><<<dim3(nbs, 2, 1), NTHREADS, kTileSmemBytes, st>>>(dS, top, r, 2);
          sep_fw_tile_kernel<T><<<dim3(nbs, nbs, 1), NTHREADS, kTileSmemBytes, st>>>(dS, top, r, 3);
          drs_count_launch(2);
        }
        SEP_CHECK();
      }
    }
  }
  SEP_EVENT(4);
  // S2x
  sep_gather_dsr_kernel<T><<<148 * 8, 256, 0, st>>>(dS, t, dSr, Duni, (uint16_t*)p->d_dSr16, p->d_flags);
  drs_count_launch();
  SEP_CHECK();
  SEP_EVENT(5);
  // S3 + S3x.  16-bit tiles for both remaining stages when every candidate sum stays below 2^15: a vertex->separator
  // distance is at most (largest distance inside a closed part) + (largest separator-to-separator distance), and an output
  // candidate adds one more in-part distance.  One 8-byte read-back decides.
  bool use16 = false;
  if (allow16 && std::is_same<T, float>::value && !p->no16) {
    float h[2] = {0, 0};
    DRS_CUDA(cudaMemcpyAsync(h, p->d_flags, 8, cudaMemcpyDeviceToHost, st));
    DRS_CUDA(cudaStreamSynchronize(st));
    use16 = h[0] + 2.0f * h[1] < (float)U16_INF;
  }
  {
    const dim3 grid(p->spad / TILE, p->max_rowtiles, p->nparts);
    if (use16) {
      const dim3 g64(p->spad / TILE, p->max_rowtiles * 2, p->nparts);
      sep_expand_rows16_kernel<64><<<g64, 128, Tile16<64>::smem_bytes, st>>>((const uint32_t*)p->d_Aprime16, (const uint16_t*)p->d_dSr16,
                                                                             (float*)Duni, t, row0, row0 + nrows);
    }
    else {
      const dim3 gx(p->spad / TILE, p->max_rowtiles * (TILE / XROWS), p->nparts);
      sep_expand_rows_kernel<T><<<gx, XCfg::NT, XCfg::smem_bytes, st>>>(Apr, dSr, Duni, t, row0, row0 + nrows);
    }
  }
  drs_count_launch();
  SEP_CHECK();
  p->used16 = use16;
  if (use16) sep_spread_kernel<T, true><<<nrows / SPREAD_ROWS, 256, 0, st>>>(Duni, Dall, C, ldc, t, row0);
  else sep_spread_kernel<T, false><<<nrows / SPREAD_ROWS, 256, 0, st>>>(Duni, Dall, C, ldc, t, row0);
  drs_count_launch();
  SEP_CHECK();
  SEP_EVENT(6);
  // S4
  if (p->ncoltiles > 0) {
    const dim3 grid(p->ncoltiles, nrows / TILE);
    const bool mirror = !directed && row0 == 0 && nrows == p->npad && !p->no_mirror;
    if (use16) {
      const uint32_t* D2 = (const uint32_t*)p->d_Dall;
      const uint16_t* A16 = (const uint16_t*)p->d_Apad16;
      const float* Mf = (const float*)p->d_M;
      float* Cf = (float*)C;
      if (p->tile_rows == 64) {
        const dim3 g64(p->ncoltiles, nrows / 64);
        if (mirror) sep_expand16_kernel<true, 64><<<g64, 128, Tile16<64>::smem_bytes, st>>>(D2, A16, Mf, Cf, ldc, t, row0);
        else sep_expand16_kernel<false, 64><<<g64, 128, Tile16<64>::smem_bytes, st>>>(D2, A16, Mf, Cf, ldc, t, row0);
      } else if (p->tile_rows == 32) {
        const dim3 g32(p->ncoltiles, nrows / 32);
        if (mirror) sep_expand16_kernel<true, 32><<<g32, 64, Tile16<32>::smem_bytes, st>>>(D2, A16, Mf, Cf, ldc, t, row0);
        else sep_expand16_kernel<false, 32><<<g32, 64, Tile16<32>::smem_bytes, st>>>(D2, A16, Mf, Cf, ldc, t, row0);
      } else if (mirror) sep_expand16_kernel<true, 128><<<grid, NTHREADS, Tile16<128>::smem_bytes, st>>>(D2, A16, Mf, Cf, ldc, t, row0);
      else sep_expand16_kernel<false, 128><<<grid, NTHREADS, Tile16<128>::smem_bytes, st>>>(D2, A16, Mf, Cf, ldc, t, row0);
    } else {
      const dim3 gx(p->ncoltiles, nrows / XROWS);
      if (mirror) sep_expand_kernel<T, true><<<gx, XCfg::NT, XCfg::smem_bytes, st>>>(Dall, Apad, M, C, ldc, t, row0);
      else sep_expand_kernel<T, false><<<gx, XCfg::NT, XCfg::smem_bytes, st>>>(Dall, Apad, M, C, ldc, t, row0);
    }
    drs_count_launch();
  }
  SEP_CHECK();
  SEP_EVENT(7);
  SEP_EVENT(8);
#undef SEP_CHECK
#undef SEP_EVENT
  return DRS_OK;
}

template <typename T>
int build_rows(drs_graph* g, int mode, T* local, int64_t ld, int row0, int nrows, cudaStream_t st) {
  drs_sep_plan* p = g->sep;
  int rc;
  if ((rc = configure_kernels<T>())) return rc;
  cudaEvent_t ev[9];
  for (auto& e : ev) DRS_CUDA(cudaEventCreate(&e));
  LevelInput in;
  in.out_ptr = g->d_out_ptr; in.out_col = g->d_out_col; in.out_w = g->d_out_w;
  rc = run_level<T>(p, in, mode, g->directed != 0, local, ld, row0, nrows, st, ev, /*allow16=*/g->weights_integral);
  if (!rc && cudaStreamSynchronize(st) != cudaSuccess) {
    drs_set_error("separator build: %s", cudaGetErrorString(cudaGetLastError()));
    rc = DRS_ERR_CUDA;
  }
  if (!rc)
    for (int i = 0; i < 8; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      p->prof_ms[i] = ms;
    }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

void free_plan(drs_sep_plan* p) {
  if (!p) return;
  free_plan(p->child);
  if (p->d_tables) drs_big_free(p->d_tables, p->tables_bytes);
  if (p->d_work) drs_big_free(p->d_work, p->work_bytes);
  delete p;
}

// Tables of one level.  Vertices [part_ptr[j], part_ptr[j+1]) are the interior of part j, [part_ptr[nparts], n) the
// separators; edges (src[e], dst[e]) in that numbering.  S_lists_out (optional): the sorted separator lists (separator-local
// ids) of the parts, from which the next level's edges are formed.
int make_plan(int n, int npad, int m, const int32_t* src, const int32_t* dst, int nparts, const int32_t* part_ptr, drs_sep_plan** out,
              std::vector<std::vector<int32_t>>* S_lists_out) {
  *out = nullptr;
  DRS_REQUIRE(part_ptr && part_ptr[0] == 0, DRS_ERR_ARG, "drs_graph_set_partition: part_ptr must start at 0");
  for (int j = 0; j < nparts; ++j)
    DRS_REQUIRE(part_ptr[j + 1] > part_ptr[j], DRS_ERR_ARG, "drs_graph_set_partition: part %d is empty or out of order", j);
  const int sep0 = part_ptr[nparts];
  DRS_REQUIRE(sep0 <= n, DRS_ERR_RANGE, "drs_graph_set_partition: parts cover %d vertices, the graph has %d", sep0, n);
  std::vector<int32_t> part_of(n, -1);
  for (int j = 0; j < nparts; ++j)
    for (int v = part_ptr[j]; v < part_ptr[j + 1]; ++v) part_of[v] = j;
  std::vector<std::vector<int32_t>> S(nparts);
  for (int e = 0; e < m; ++e) {
    const int a = src[e], b = dst[e];
    const int pa = part_of[a], pb = part_of[b];
    if (pa >= 0 && pb >= 0) {
      DRS_REQUIRE(pa == pb, DRS_ERR_ARG, "drs_graph_set_partition: edge %d (%d - %d) joins the interiors of parts %d and %d", e, a, b,
                  pa, pb);
    } else if (pa >= 0) S[pa].push_back(b - sep0);
    else if (pb >= 0) S[pb].push_back(a - sep0);
  }
  const int nsep = n - sep0;
  std::vector<char> listed(std::max(nsep, 1), 0);
  for (auto& s : S) {
    std::sort(s.begin(), s.end());
    s.erase(std::unique(s.begin(), s.end()), s.end());
    for (int r : s) listed[r] = 1;
  }
  for (int r = 0; r < nsep; ++r)
    if (!listed[r]) {   // a separator with separator neighbours only: any closed part may carry it
      int best = 0;
      for (int j = 1; j < nparts; ++j)
        if (S[j].size() < S[best].size()) best = j;
      S[best].insert(std::lower_bound(S[best].begin(), S[best].end(), r), r);
    }
  drs_sep_plan* p = new (std::nothrow) drs_sep_plan();
  DRS_REQUIRE(p, DRS_ERR_NOMEM, "drs_graph_set_partition: out of host memory");
  p->n = n; p->npad = npad;
  p->nparts = nparts; p->sep0 = sep0; p->nsep = nsep;
  p->spad = std::max(TILE, round_up(nsep, TILE));
  p->pstart.assign(part_ptr, part_ptr + nparts); p->pend.assign(part_ptr + 1, part_ptr + nparts + 1);
  p->sj.resize(nparts); p->kpad.resize(nparts); p->koff.resize(nparts); p->qpad.resize(nparts); p->cbase.resize(nparts);
  p->wpad.resize(nparts); p->moff.resize(nparts); p->apoff.resize(nparts);
  int ksum = 0;
  int64_t melems = 0, apelems = 0;
  p->kmax = KC;
  for (int j = 0; j < nparts; ++j) {
    const int nj = p->pend[j] - p->pstart[j], sj = (int)S[j].size();
    p->sj[j] = sj; p->kpad[j] = round_up(sj, KC / 2); p->koff[j] = ksum; ksum += p->kpad[j];
    p->kmax = std::max(p->kmax, round_up(p->kpad[j], KC));
    p->qpad[j] = round_up(nj + sj, TILE); p->moff[j] = melems; melems += (int64_t)p->qpad[j] * p->qpad[j];
    p->maxq = std::max(p->maxq, nj + sj);
    p->maxnb = std::max(p->maxnb, p->qpad[j] / TILE);
    p->max_rowtiles = std::max(p->max_rowtiles, (nj + TILE - 1) / TILE);
    p->cbase[j] = p->pstart[j] / 4 * 4; p->wpad[j] = round_up(p->pend[j] - p->cbase[j], TILE);
    p->apoff[j] = apelems; apelems += (int64_t)round_up(p->kpad[j], KC) * p->wpad[j];
    for (int c0 = p->cbase[j]; c0 < p->pend[j]; c0 += TILE) { p->ct_part.push_back(j); p->ct_col0.push_back(c0); }
  }
  p->ktot = std::max(TILE, round_up(ksum + KC / 2, TILE));   // a half chunk is loaded whole: 16 readable columns past the last range
  p->m_elems = melems; p->ap_elems = apelems;
  p->sall.assign(p->ktot, -1);
  p->sepcol.assign(std::max(npad - sep0, 1), 0);
  std::vector<char> have(std::max(nsep, 1), 0);
  for (int j = 0; j < nparts; ++j)
    for (int k = 0; k < (int)S[j].size(); ++k) {
      const int r = S[j][k];
      p->sall[p->koff[j] + k] = r;
      if (!have[r]) { have[r] = 1; p->sepcol[r] = p->koff[j] + k; }
    }
  p->ncoltiles = (int)p->ct_part.size();
  if (p->ct_part.empty()) { p->ct_part.push_back(0); p->ct_col0.push_back(0); }
  // one device arena for the tables
  struct Slice { const void* src; size_t bytes; size_t off; };
  std::vector<Slice> sl;
  size_t off = 0;
  auto add = [&](const void* src_, size_t bytes) { sl.push_back({src_, bytes, off}); const size_t o = off; off = (off + std::max<size_t>(bytes, 4) + 255) & ~(size_t)255; return o; };
  const size_t o_ps = add(p->pstart.data(), 4 * nparts), o_pe = add(p->pend.data(), 4 * nparts), o_sj = add(p->sj.data(), 4 * nparts);
  const size_t o_kp = add(p->kpad.data(), 4 * nparts), o_ko = add(p->koff.data(), 4 * nparts), o_qp = add(p->qpad.data(), 4 * nparts);
  const size_t o_cb = add(p->cbase.data(), 4 * nparts), o_wp = add(p->wpad.data(), 4 * nparts);
  const size_t o_mo = add(p->moff.data(), 8 * nparts), o_ao = add(p->apoff.data(), 8 * nparts);
  const size_t o_sa = add(p->sall.data(), 4 * p->sall.size()), o_sc = add(p->sepcol.data(), 4 * p->sepcol.size());
  const size_t o_cp = add(p->ct_part.data(), 4 * p->ct_part.size()), o_cq = add(p->ct_col0.data(), 4 * p->ct_col0.size());
  const int64_t top_moff = 0;
  const int32_t top_qpad = p->spad;
  const size_t o_tm = add(&top_moff, 8), o_tq = add(&top_qpad, 4);
  std::vector<unsigned char> img(off, 0);
  for (const Slice& x : sl) std::memcpy(img.data() + x.off, x.src, x.bytes);
  p->tables_bytes = img.size();
  if (drs_big_alloc(&p->d_tables, img.size()) ||
      cudaMemcpy(p->d_tables, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    if (p->d_tables) drs_big_free(p->d_tables, p->tables_bytes);
    delete p;
    DRS_REQUIRE(false, DRS_ERR_CUDA, "drs_graph_set_partition: device tables (%zu bytes) could not be created", img.size());
  }
  const char* b = (const char*)p->d_tables;
  p->d_pstart = (const int32_t*)(b + o_ps); p->d_pend = (const int32_t*)(b + o_pe); p->d_sj = (const int32_t*)(b + o_sj);
  p->d_kpad = (const int32_t*)(b + o_kp); p->d_koff = (const int32_t*)(b + o_ko); p->d_qpad = (const int32_t*)(b + o_qp);
  p->d_cbase = (const int32_t*)(b + o_cb); p->d_wpad = (const int32_t*)(b + o_wp);
  p->d_moff = (const int64_t*)(b + o_mo); p->d_apoff = (const int64_t*)(b + o_ao);
  p->d_sall = (const int32_t*)(b + o_sa); p->d_sepcol = (const int32_t*)(b + o_sc);
  p->d_ct_part = (const int32_t*)(b + o_cp); p->d_ct_col0 = (const int32_t*)(b + o_cq);
  p->d_top_moff = (const int64_t*)(b + o_tm); p->d_top_qpad = (const int32_t*)(b + o_tq);
  { const char* e = getenv("DRS_SEP_NO_MIRROR"); p->no_mirror = e && e[0] == '1'; }
  { const char* e = getenv("DRS_SEP_NO_U16"); p->no16 = e && e[0] == '1'; }
  { const char* e = getenv("DRS_SEP_TILE_ROWS"); const int v = e ? atoi(e) : 32; p->tile_rows = (v == 128 || v == 64) ? v : 32; }
  if (S_lists_out) *S_lists_out = std::move(S);
  *out = p;
  return DRS_OK;
}

}  // namespace

void drs_sep_free(drs_graph* g) {
  free_plan(g->sep);
  g->sep = nullptr;
}

// Non-integer travel times: the build re-associates the fp32 sums, so its matrix differs from the SSSP / FW one in the last
// bits (relative 1e-7; the drop-in's tolerance on float maps is 1e-5 and its answers are fp64 sums along the path the matrix
// selects).  On by default; DRS_SEP_FLOAT=0 keeps float maps on the SSSP builder.
static bool sep_float_ok() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DRS_SEP_FLOAT"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

bool drs_sep_allowed(const drs_graph* g) { return g->sep && (g->weights_integral || sep_float_ok()); }

bool drs_sep_usable(const drs_graph* g) {
  // exact arithmetic (any association of the sums gives the same bits) and separators that are a small part of the map
  return g->sep && (g->weights_integral || sep_float_ok()) && (int64_t)g->sep->nsep * 3 <= g->n &&
         (int64_t)g->sep->ktot * 2 <= g->npad + 256;
}

extern "C" int drs_graph_set_partition(drs_graph* g, int32_t nparts, const int32_t* part_ptr, int32_t n_sep_parts,
                                       const int32_t* sep_part_ptr) {
  DRS_REQUIRE(g, DRS_ERR_ARG, "drs_graph_set_partition: null graph");
  DrsDeviceGuard guard(g->device);
  drs_sep_free(g);
  if (nparts <= 0) return DRS_OK;   // partition removed
  DrsHostTimer tm("drs_graph_set_partition");
  drs_sep_plan* p = nullptr;
  std::vector<std::vector<int32_t>> S;
  int rc = make_plan(g->n, g->npad, g->m, g->h_src.data(), g->h_dst.data(), nparts, part_ptr, &p, &S);
  if (rc) return rc;
  tm.lap("level 1 tables");
  if (n_sep_parts > 0 && sep_part_ptr && p->nsep > 0) {
    // level 2: the separator graph (separator-local ids): the map's edges inside S and a clique per S_j
    std::vector<int32_t> s2, d2;
    for (int e = 0; e < g->m; ++e)
      if (g->h_src[e] >= p->sep0 && g->h_dst[e] >= p->sep0 && g->h_src[e] != g->h_dst[e]) {
        s2.push_back(g->h_src[e] - p->sep0); d2.push_back(g->h_dst[e] - p->sep0);
      }
    // a part's separator list is a clique of that graph; for the tables a star is enough (same components, and the
    // separator lists of the groups only need the group-to-top adjacencies): centre = the first member that is not a top
    // separator (a list made of top separators only joins no group to anything)
    const int top0 = sep_part_ptr[n_sep_parts];
    for (const auto& lst : S) {
      int centre = -1;
      for (int r : lst) if (r < top0) { centre = r; break; }
      if (centre < 0) continue;
      for (int r : lst) if (r != centre) { s2.push_back(centre); d2.push_back(r); }
    }
    tm.lap("level 2 edges");
    if (make_plan(p->nsep, p->spad, (int)s2.size(), s2.data(), d2.data(), n_sep_parts, sep_part_ptr, &p->child, nullptr)) p->child = nullptr;
    tm.lap("level 2 tables");
  }
  g->sep = p;
  return DRS_OK;
}

extern "C" int drs_sep_info(const drs_graph* g, int32_t* nparts, int32_t* nsep, int32_t* ktot, int32_t* kmax, int32_t* usable) {
  DRS_REQUIRE(g, DRS_ERR_ARG, "drs_sep_info: null graph");
  const drs_sep_plan* p = g->sep;
  if (nparts) *nparts = p ? p->nparts : 0;
  if (nsep) *nsep = p ? p->nsep : 0;
  if (ktot) *ktot = p ? p->ktot : 0;
  if (kmax) *kmax = p ? p->kmax : 0;
  if (usable) *usable = drs_sep_usable(g) ? 1 : 0;
  return DRS_OK;
}

extern "C" int drs_sep_info2(const drs_graph* g, int32_t* n_sep_parts, int32_t* n_top, int32_t* used16) {
  DRS_REQUIRE(g, DRS_ERR_ARG, "drs_sep_info2: null graph");
  const drs_sep_plan* c = g->sep ? g->sep->child : nullptr;
  if (n_sep_parts) *n_sep_parts = c ? c->nparts : 0;
  if (n_top) *n_top = c ? c->nsep : 0;
  if (used16) *used16 = (g->sep && g->sep->used16) ? 1 : 0;
  return DRS_OK;
}

namespace {
// relaxations (one add + one min) the stages of a level execute: [0] part closures, [1] separator closure, [2] vertex ->
// separator rows, [3] output tiles as launched, [4] output tiles without padding (n * sum_j n_j * s_j, halved when mirrored)
void level_work(const drs_sep_plan* p, bool directed, int row0, int nrows, double* w) {
  w[0] = w[1] = w[2] = w[3] = w[4] = 0;
  const bool whole = row0 == 0 && nrows == p->npad;
  const int row1 = row0 + nrows;
  for (int j = 0; j < p->nparts; ++j) {
    const double q = p->qpad[j];
    w[0] += q * q * q;
    const int rows = std::max(0, std::min(row1, p->pend[j]) - std::max(row0, p->pstart[j]));
    w[2] += (double)round_up(rows, TILE) * p->spad * p->kpad[j];
    w[4] += (double)std::max(0, std::min(row1, p->n) - row0) * (p->pend[j] - p->pstart[j]) * p->sj[j];
  }
  if (p->nsep > 0) {
    if (p->child) {
      double c[5];
      level_work(p->child, directed, 0, p->child->npad, c);
      w[1] = c[0] + c[1] + c[2] + c[3];
    } else {
      w[1] = (double)p->spad * p->spad * p->spad;
    }
  }
  const bool mirror = !directed && !p->no_mirror && whole;
  for (int ct = 0; ct < p->ncoltiles; ++ct) {
    const int j = p->ct_part[ct], c0 = p->ct_col0[ct];
    const int c1 = c0 + std::min(p->pend[j] - c0, TILE);
    const int th = p->used16 ? p->tile_rows : XROWS;   // height of the output tiles of the last build (16-bit or 32-bit path)
    for (int r0 = row0; r0 < row1; r0 += th)
      if (!mirror || c1 > r0 || r0 + th > p->sep0) w[3] += (double)th * TILE * p->kpad[j];
  }
  if (mirror) w[4] *= 0.5;
}
}  // namespace

extern "C" int drs_sep_work(const drs_graph* g, int32_t row0, int32_t nrows, double* relax5) {
  DRS_REQUIRE(g && relax5, DRS_ERR_ARG, "drs_sep_work: null argument");
  DRS_REQUIRE(g->sep, DRS_ERR_STATE, "drs_sep_work: the graph has no partition");
  if (nrows <= 0) { row0 = 0; nrows = g->npad; }
  DRS_REQUIRE(row0 >= 0 && row0 + nrows <= g->npad && row0 % TILE == 0 && nrows % TILE == 0, DRS_ERR_ARG, "drs_sep_work: bad shard");
  level_work(g->sep, g->directed != 0, row0, nrows, relax5);
  return DRS_OK;
}

extern "C" int drs_apsp_sep_profile(const drs_graph* g, double* ms8) {
  DRS_REQUIRE(g && ms8, DRS_ERR_ARG, "drs_apsp_sep_profile: null argument");
  DRS_REQUIRE(g->sep, DRS_ERR_STATE, "drs_apsp_sep_profile: the graph has no partition");
  for (int i = 0; i < 8; ++i) ms8[i] = g->sep->prof_ms[i];
  return DRS_OK;
}

extern "C" int drs_apsp_rows_sep(const drs_graph* gc, int32_t mode, void* local_dev, int64_t ld, int32_t row0, int32_t nrows,
                                 void* stream) {
  DRS_REQUIRE(gc && local_dev, DRS_ERR_ARG, "drs_apsp_rows_sep: null argument");
  drs_graph* g = const_cast<drs_graph*>(gc);
  { int rcm = drs_check_mode(g, mode, "drs_apsp_rows_sep"); if (rcm) return rcm; }
  DRS_REQUIRE(g->sep, DRS_ERR_STATE, "drs_apsp_rows_sep: the graph has no partition (drs_graph_set_partition)");
  DRS_REQUIRE(g->weights_integral || sep_float_ok(), DRS_ERR_ARG,
              "drs_apsp_rows_sep: non-integer travel times and DRS_SEP_FLOAT=0 (the build re-associates sums; use the SSSP builder)");
  DRS_REQUIRE(ld == g->npad && row0 >= 0 && nrows >= 0 && row0 + nrows <= g->npad && row0 % TILE == 0 && nrows % TILE == 0,
              DRS_ERR_ARG, "drs_apsp_rows_sep: bad shard (ld=%lld row0=%d nrows=%d npad=%d; rows in multiples of %d)", (long long)ld,
              row0, nrows, g->npad, TILE);
  if (nrows == 0) return DRS_OK;
  DrsDeviceGuard guard(g->device);
  if (mode == DRS_MODE_F32) return build_rows<float>(g, mode, (float*)local_dev, ld, row0, nrows, (cudaStream_t)stream);
  return build_rows<int>(g, mode, (int*)local_dev, ld, row0, nrows, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ host: the partition
namespace {

struct Bisector {
  int n, target;
  const double *x, *y;           // null: no usable coordinates
  std::vector<int32_t> adj_ptr, adj;   // undirected adjacency (for the coordinate-free split)
  std::vector<int32_t> idx;
  std::vector<std::pair<int, int>> leaves;
  std::vector<int32_t> leaf_super;     // group of each leaf at the second level: the subtree of <= leaves_per_super leaves it lies in
  int leaves_per_super = 1, nsuper = 0;
  std::vector<int32_t> stamp_in, stamp_seen;
  int stamp = 0, seen_ctr = 0;

  void bfs_order(int lo, int hi) {
    // level-structure order of the subgraph induced by idx[lo, hi), started from a pseudo-peripheral vertex
    ++stamp;
    for (int i = lo; i < hi; ++i) stamp_in[idx[i]] = stamp;
    std::vector<int32_t> order;
    order.reserve(hi - lo);
    auto sweep = [&](int start, int seen_stamp, std::vector<int32_t>& out) {
      size_t head = out.size();
      out.push_back(start);
      stamp_seen[start] = seen_stamp;
      while (head < out.size()) {
        const int u = out[head++];
        for (int a = adj_ptr[u]; a < adj_ptr[u + 1]; ++a) {
          const int v = adj[a];
          if (stamp_in[v] == stamp && stamp_seen[v] != seen_stamp) { stamp_seen[v] = seen_stamp; out.push_back(v); }
        }
      }
    };
    std::vector<int32_t> probe;
    sweep(idx[lo], ++seen_ctr, probe);
    const int far = probe.back();
    const int seen2 = ++seen_ctr;
    sweep(far, seen2, order);
    for (int i = lo; i < hi; ++i)
      if (stamp_seen[idx[i]] != seen2) sweep(idx[i], seen2, order);   // further components
    std::copy(order.begin(), order.end(), idx.begin() + lo);
  }

  // where a range of c vertices is cut: L = leaves it still has to yield, the left child takes floor(L/2) of them
  int cut(int lo, int hi) const {
    const int c = hi - lo, L = (c + target - 1) / target;
    return lo + (int)((int64_t)c * (L / 2) / L);
  }

  // pass 1 (coordinates): median cuts along the wider axis; subtrees are independent, the upper levels run on threads
  void arrange(int lo, int hi, int depth) {
    if (hi - lo <= target) return;
    const int mid = cut(lo, hi);
    double x0 = x[idx[lo]], x1 = x0, y0 = y[idx[lo]], y1 = y0;
    for (int i = lo; i < hi; ++i) {
      x0 = std::min(x0, x[idx[i]]); x1 = std::max(x1, x[idx[i]]);
      y0 = std::min(y0, y[idx[i]]); y1 = std::max(y1, y[idx[i]]);
    }
    const double* k = (x1 - x0 >= y1 - y0) ? x : y;
    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                     [k](int a, int b) { return k[a] < k[b] || (k[a] == k[b] && a < b); });
    if (depth < 3 && hi - lo > 8192) {
      std::thread left([this, lo, mid, depth]() { arrange(lo, mid, depth + 1); });
      arrange(mid, hi, depth + 1);
      left.join();
    } else {
      arrange(lo, mid, depth + 1);
      arrange(mid, hi, depth + 1);
    }
  }

  // pass 2: the leaves in tree order, and the second-level group of each
  void split(int lo, int hi, int sup) {
    const int c = hi - lo;
    const int L = (c + target - 1) / target;
    if (sup < 0 && L <= leaves_per_super) sup = nsuper++;
    if (c <= target) { if (c > 0) { leaves.push_back({lo, hi}); leaf_super.push_back(sup); } return; }
    const int mid = cut(lo, hi);
    if (!x) bfs_order(lo, hi);
    split(lo, mid, sup);
    split(mid, hi, sup);
  }
};

}  // namespace

// Part-major device numbering for the separator build (and a locality numbering for everything else: parts follow the
// bisection tree, so neighbouring vertices keep neighbouring ids).  order_out[k] = caller's id of device vertex k;
// part_ptr_out[0..nparts] = id ranges of the part interiors, part_ptr_out[nparts] = first separator id.
// Parts: recursive coordinate bisection (median cuts along the wider axis) when lng/lat are usable, otherwise recursive
// bisection of breadth-first level structures.  Separators: a greedy vertex cover of the cut edges.
extern "C" int drs_partition_order(int32_t n, int32_t m, const int32_t* src, const int32_t* dst, const double* lng, const double* lat,
                                   int32_t target, int32_t* order_out, int32_t* part_ptr_out, int32_t max_parts, int32_t* nparts_out,
                                   int32_t* sep_part_ptr_out, int32_t* n_sep_parts_out) {
  DRS_REQUIRE(n > 0 && m >= 0 && (m == 0 || (src && dst)) && order_out && part_ptr_out && nparts_out && target > 0 && max_parts > 0,
              DRS_ERR_ARG, "drs_partition_order: bad argument");
  for (int e = 0; e < m; ++e)
    DRS_REQUIRE(src[e] >= 0 && src[e] < n && dst[e] >= 0 && dst[e] < n, DRS_ERR_RANGE, "drs_partition_order: edge %d endpoint out of range", e);
  DrsHostTimer tm("drs_partition_order");
  Bisector b;
  b.n = n; b.target = target; b.x = b.y = nullptr;
  if (lng && lat) {
    bool ok = true, varied = false;
    for (int i = 0; i < n && ok; ++i) {
      ok = std::isfinite(lng[i]) && std::isfinite(lat[i]);
      varied = varied || lng[i] != lng[0] || lat[i] != lat[0];
    }
    if (ok && varied) { b.x = lng; b.y = lat; }
  }
  b.idx.resize(n);
  std::iota(b.idx.begin(), b.idx.end(), 0);
  if (!b.x) {
    b.adj_ptr.assign(n + 1, 0);
    for (int e = 0; e < m; ++e) if (src[e] != dst[e]) { b.adj_ptr[src[e] + 1]++; b.adj_ptr[dst[e] + 1]++; }
    std::partial_sum(b.adj_ptr.begin(), b.adj_ptr.end(), b.adj_ptr.begin());
    b.adj.resize(b.adj_ptr[n]);
    std::vector<int32_t> pos(b.adj_ptr.begin(), b.adj_ptr.end() - 1);
    for (int e = 0; e < m; ++e) if (src[e] != dst[e]) { b.adj[pos[src[e]]++] = dst[e]; b.adj[pos[dst[e]]++] = src[e]; }
    b.stamp_in.assign(n, 0); b.stamp_seen.assign(n, 0);
  }
  {
    const int Ltot = (n + target - 1) / target;
    const char* e = getenv("DRS_SEP_SUPER");
    b.leaves_per_super = e ? std::max(1, atoi(e)) : std::max(2, (int)std::ceil(1.25 * std::sqrt((double)Ltot)));
  }
  tm.lap("setup");
  if (b.x) b.arrange(0, n, 0);
  b.split(0, n, -1);
  tm.lap("bisection");
  std::vector<int32_t> part(n);
  for (int j = 0; j < (int)b.leaves.size(); ++j)
    for (int i = b.leaves[j].first; i < b.leaves[j].second; ++i) part[b.idx[i]] = j;
  std::vector<int32_t> cutdeg(n, 0);
  for (int e = 0; e < m; ++e)
    if (part[src[e]] != part[dst[e]]) { cutdeg[src[e]]++; cutdeg[dst[e]]++; }
  std::vector<char> sep(n, 0);
  for (int e = 0; e < m; ++e) {
    const int u = src[e], v = dst[e];
    if (part[u] == part[v] || sep[u] || sep[v]) continue;
    const int pick = cutdeg[u] > cutdeg[v] ? u : cutdeg[v] > cutdeg[u] ? v : (part[u] < part[v] ? u : v);
    sep[pick] = 1;
  }
  {
    // a cover vertex whose neighbours are all cover vertices covers nothing on its own: back to its part's interior
    std::vector<int32_t> nbr_ptr(n + 1, 0);
    for (int e = 0; e < m; ++e) if (src[e] != dst[e]) { nbr_ptr[src[e] + 1]++; nbr_ptr[dst[e] + 1]++; }
    std::partial_sum(nbr_ptr.begin(), nbr_ptr.end(), nbr_ptr.begin());
    std::vector<int32_t> nbr(nbr_ptr[n]), pos(nbr_ptr.begin(), nbr_ptr.end() - 1);
    for (int e = 0; e < m; ++e) if (src[e] != dst[e]) { nbr[pos[src[e]]++] = dst[e]; nbr[pos[dst[e]]++] = src[e]; }
    for (int v = 0; v < n; ++v) {
      if (!sep[v]) continue;
      bool all_sep = true;
      for (int a = nbr_ptr[v]; a < nbr_ptr[v + 1] && all_sep; ++a) all_sep = sep[nbr[a]] != 0;
      if (all_sep) sep[v] = 0;
    }
  }
  tm.lap("vertex cover");
  int k = 0, np = 0;
  part_ptr_out[0] = 0;
  for (const auto& lf : b.leaves) {
    const int k0 = k;
    for (int i = lf.first; i < lf.second; ++i)
      if (!sep[b.idx[i]]) order_out[k++] = b.idx[i];
    if (k > k0) {
      DRS_REQUIRE(np < max_parts, DRS_ERR_CAPACITY, "drs_partition_order: more than %d parts (raise max_parts or the target size)", max_parts);
      part_ptr_out[++np] = k;
    }
  }
  *nparts_out = np;
  // second level: a separator with a neighbour in another group of leaves belongs to the top separator set T; the others
  // follow their own leaf's group.  Separators are numbered group by group, T last; in the separator graph (edges inside S
  // and a clique per part's separator list) no edge joins two groups, so (groups, T) is a separator partition of it.
  std::vector<char> top(n, 0);
  for (int e = 0; e < m; ++e) {
    const int u = src[e], v = dst[e];
    if (b.leaf_super[part[u]] != b.leaf_super[part[v]]) { if (sep[u]) top[u] = 1; if (sep[v]) top[v] = 1; }
  }
  const int sep0 = k;
  int nsp = 0;
  std::vector<std::vector<int32_t>> group(std::max(b.nsuper, 1));
  for (int j = 0; j < (int)b.leaves.size(); ++j)
    for (int i = b.leaves[j].first; i < b.leaves[j].second; ++i) {
      const int v = b.idx[i];
      if (sep[v] && !top[v]) group[std::max(b.leaf_super[j], 0)].push_back(v);
    }
  if (sep_part_ptr_out) sep_part_ptr_out[0] = 0;
  for (const auto& grp : group) {
    if (grp.empty()) continue;
    for (int v : grp) order_out[k++] = v;
    if (sep_part_ptr_out && nsp < max_parts) sep_part_ptr_out[++nsp] = k - sep0;
  }
  for (const auto& lf : b.leaves)
    for (int i = lf.first; i < lf.second; ++i)
      if (sep[b.idx[i]] && top[b.idx[i]]) order_out[k++] = b.idx[i];
  if (n_sep_parts_out) *n_sep_parts_out = sep_part_ptr_out ? nsp : 0;
  tm.lap("ordering");
  return DRS_OK;
}
