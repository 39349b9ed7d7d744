// RTV trip enumeration, one level for the whole fleet (host side of the Alonso-Mora tick).
//
// Replaces the per-vehicle Python loops of AlonsoMora.generateRTVGraph (reference lib/Optimization.py:265-347): which
// request sets of size k become candidate trips of a vehicle, given its feasible trips of size k-1 and the request-request
// graph.  Set logic only -- the candidates are then costed on the GPU by drs_tick_trip_eval.  No device work here: at
// BASELINE config C4 (10 000 vehicles, ~3*10^5 pair trips) the enumeration is a few 10^7 set operations.
#include <cstdint>
#include <cstring>
#include <unordered_set>
#include <vector>

#include "drs_b200.h"
#include "drs_common.cuh"

namespace {

constexpr int MAXK = DRS_MAX_TRIP_REQS;

struct TripKey {
    int32_t m[MAXK];
    bool operator==(const TripKey& o) const { return std::memcmp(m, o.m, sizeof(m)) == 0; }
};
struct TripHash {
    size_t operator()(const TripKey& k) const {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (int i = 0; i < MAXK; ++i) {
            h ^= (uint64_t)(uint32_t)k.m[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
            h *= 0xBF58476D1CE4E5B9ull;
        }
        return (size_t)(h ^ (h >> 31));
    }
};
using TripSet = std::unordered_set<TripKey, TripHash>;

inline bool rr_edge(const uint64_t* rr, int64_t words, int32_t a, int32_t b) {
    return (rr[(int64_t)a * words + (b >> 6)] >> (b & 63)) & 1ull;
}

struct Sink {
    int64_t cap, n = 0;
    int32_t k;
    int32_t* veh;
    int32_t* req;
    void push(int32_t v, const int32_t* members) {
        if (n < cap) {
            veh[n] = v;
            for (int i = 0; i < k; ++i) req[n * k + i] = members[i];
        }
        ++n;
    }
};

}  // namespace

extern "C" int drs_rtv_level(int32_t n_veh, int32_t k, const int32_t* key_off, const int32_t* key_req, const uint8_t* veh_on,
                             int32_t n_req, const uint64_t* rr_bits, int64_t cap_tasks, int32_t* task_veh, int32_t* task_req,
                             int64_t* n_tasks, int64_t* n_saved) {
    DRS_REQUIRE(n_veh >= 0 && k >= 2 && key_off && n_tasks && n_saved && (cap_tasks <= 0 || (task_veh && task_req)) &&
                    (n_req == 0 || rr_bits),
                DRS_ERR_ARG, "drs_rtv_level: bad argument");
    DRS_REQUIRE(k <= MAXK, DRS_ERR_CAPACITY, "drs_rtv_level: a trip holds at most %d new requests", MAXK);
    const int64_t words = ((int64_t)n_req + 63) / 64;
    const int q = k - 1;                       // size of the keys
    Sink out{cap_tasks, 0, k, task_veh, task_req};
    int64_t saved = 0;
    TripSet keys, tested;
    for (int32_t v = 0; v < n_veh; ++v) {
        const int32_t lo = key_off[v], hi = key_off[v + 1];
        if (hi <= lo || (veh_on && !veh_on[v])) continue;
        const int32_t* K = key_req + (int64_t)lo * q;
        const int32_t g = hi - lo;
        if (k == 2) {
            // lib/Optimization.py:265-288: every pair of the vehicle's single requests that share a request-request edge
            for (int32_t x = 0; x < g; ++x)
                for (int32_t y = x + 1; y < g; ++y) {
                    int32_t a = K[x], b = K[y];
                    if (rr_edge(rr_bits, words, a, b)) {
                        int32_t m[2] = {a, b};
                        out.push(v, m);
                    } else {
                        ++saved;
                    }
                }
            continue;
        }
        // lib/Optimization.py:291-347: unions of two (k-1)-trips with k members, each tested once; every (k-1)-subset
        // must itself be one of the vehicle's trips, otherwise the reference LEAVES the inner loop (:325-327), and
        // the requests must be pairwise joined in the request-request graph
        keys.clear();
        tested.clear();
        for (int32_t x = 0; x < g; ++x) {
            TripKey t{};
            for (int i = 0; i < MAXK; ++i) t.m[i] = i < q ? K[(int64_t)x * q + i] : -1;
            keys.insert(t);
        }
        for (int32_t x = 0; x < g; ++x) {
            const int32_t* t1 = K + (int64_t)x * q;
            for (int32_t y = 0; y < g; ++y) {
                const int32_t* t2 = K + (int64_t)y * q;
                // sorted union; give up as soon as it exceeds k members
                int32_t u[MAXK + 1];
                int nu = 0, i = 0, j = 0;
                while ((i < q || j < q) && nu <= k) {
                    if (j >= q || (i < q && t1[i] < t2[j])) u[nu++] = t1[i++];
                    else if (i >= q || t2[j] < t1[i]) u[nu++] = t2[j++];
                    else { u[nu++] = t1[i]; ++i; ++j; }
                }
                if (nu != k || i < q || j < q) continue;
                TripKey trip{};
                for (int a = 0; a < MAXK; ++a) trip.m[a] = a < k ? u[a] : -1;
                if (!tested.insert(trip).second) continue;
                bool subsets_ok = true;
                for (int r = 0; r < k && subsets_ok; ++r) {
                    TripKey sub{};
                    int n = 0;
                    for (int a = 0; a < k; ++a)
                        if (a != r) sub.m[n++] = u[a];
                    for (; n < MAXK; ++n) sub.m[n] = -1;
                    subsets_ok = keys.count(sub) != 0;
                }
                if (!subsets_ok) {
                    ++saved;
                    break;
                }
                bool clique = true;
                for (int a = 0; a < k && clique; ++a)
                    for (int b = a + 1; b < k && clique; ++b) clique = rr_edge(rr_bits, words, u[a], u[b]);
                if (clique) out.push(v, u);
                else ++saved;
            }
        }
    }
    *n_tasks = out.n;
    *n_saved = saved;
    return DRS_OK;
}
