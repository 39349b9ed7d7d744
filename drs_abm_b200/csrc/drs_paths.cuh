// Canonical shortest-path predecessors, shared by the query kernels (drs_api.cu) and the insertion commit
// (insertion_kernels.cu).  Rule (DESIGN.md "canonical predecessor"): among in-arcs (u -> t, w, eid) minimise
// (dist[s,u] + w, dist[s,u], eid).
#pragma once
#include "drs_common.cuh"

__device__ __forceinline__ int prev_node(int e, int t, const int32_t* esrc, const int32_t* edst) {
  const int a = esrc[e], b = edst[e];
  return (b == t) ? a : b;   // undirected edges are walked against either orientation
}

// Where the last edge of the canonical path s -> t comes from: the materialised predecessor matrix, or --
// when it was not built (DRS_PRED_LAZY) -- the canonical rule evaluated on the spot from the distance row
// of s and t's in-arcs (same rule as pred_kernel: minimise (dist[s,u] + w, dist[s,u], eid)).
struct PredSource {
  const int32_t* pred;      // may be null
  const void* dist;
  int64_t ld;
  int mode;
  const int32_t *in_ptr, *in_src, *in_eid;
  const float* in_w;

  template <typename T> __device__ int rule(const T* row, T inf, int t) const {
    int best_e = -1;
    T best_c = inf, best_du = inf;
    for (int a = in_ptr[t]; a < in_ptr[t + 1]; ++a) {
      const T du = row[in_src[a]];
      if (!(du < inf)) continue;
      const T cand = du + (T)in_w[a];
      const int e = in_eid[a];
      if (cand < best_c || (cand == best_c && (du < best_du || (du == best_du && e < best_e)))) {
        best_c = cand; best_du = du; best_e = e;
      }
    }
    return best_e;
  }
  __device__ int last_edge(int s, int t) const {
    if (pred) return pred[(int64_t)s * ld + t];
    if (s == t) return -1;
    if (mode == DRS_MODE_F32) {
      const float* row = (const float*)dist + (int64_t)s * ld;
      const float inf = __int_as_float(0x7f800000);
      return (row[t] < inf) ? rule<float>(row, inf, t) : -1;
    }
    const int* row = (const int*)dist + (int64_t)s * ld;
    return (row[t] < DRS_I32_INF) ? rule<int>(row, DRS_I32_INF, t) : -1;
  }
};


inline PredSource drs_pred_source(const drs_graph* g) {
  return PredSource{g->d_pred, g->d_dist, g->ld, g->mode, g->d_in_ptr, g->d_in_src, g->d_in_eid, g->d_in_w};
}
