// ref_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-callable wrapper around the UNMODIFIED reference headers, compiled where they lie
// (-I/root/reference; see oracle/Makefile).  Nothing from the reference is copied into this
// repository: the build output (oracle/_ref/*.so) is git-ignored.
//
// Two flavours are built from this one file:
//   libbnref_pure.so   the headers exactly as they are (belief_propagation.hpp:31-159);
//                      max_sweeps is ignored, a non-converging case never returns.
//   libbnref.so        -DBNREF_CAPPED: the single `while(true)` at belief_propagation.hpp:75 is
//                      turned into a counted loop by a preprocessor shim (the header text itself
//                      is still the file under /root/reference), which gives the sweep count and
//                      a sweep cap the public API lacks.
// tests/test_oracle.py cross-checks the two flavours against each other.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iterator>
#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#include <cassert>

#include <bayesian/graph.hpp>
#include <bayesian/matrix.hpp>

#ifdef BNREF_CAPPED
static thread_local int g_cap = 1 << 30;
static thread_local int g_iter = 0;
// every standard header the reference pulls in is already included above (include guards),
// so the only `while` the preprocessor can still meet is belief_propagation.hpp:75.
#define while(cond) for (g_iter = 0; g_iter < g_cap; ++g_iter)
#include <bayesian/inference/belief_propagation.hpp>
#undef while
#else
#include <bayesian/inference/belief_propagation.hpp>
#endif

namespace {

struct built {
    bn::graph_t graph;
    std::vector<bn::vertex_type> v;
};

void build(built& b, int32_t n, const int32_t* card, const int32_t* poff, const int32_t* par,
           const int64_t* coff, const double* cpt)
{
    for (int32_t i = 0; i < n; ++i) {
        auto v = b.graph.add_vertex();
        v->id = i;
        v->selectable_num = static_cast<std::size_t>(card[i]);
        b.v.push_back(v);
    }
    // ascending child order: the new edge's target has no out-edges yet whenever node ids are
    // topologically sorted, which keeps graph_t::is_able_trace (graph.hpp:437-449) cheap.
    for (int32_t x = 0; x < n; ++x)
        for (int32_t e = poff[x]; e < poff[x + 1]; ++e)
            b.graph.add_edge(b.v[par[e]], b.v[x]);
    for (int32_t x = 0; x < n; ++x) {
        std::vector<bn::vertex_type> parents;
        for (int32_t e = poff[x]; e < poff[x + 1]; ++e) parents.push_back(b.v[par[e]]);
        b.v[x]->cpt.assign(parents, b.v[x]);
        int32_t k = poff[x + 1] - poff[x], r = card[x];
        std::vector<int> cfg(static_cast<std::size_t>(k), 0);
        int64_t Q = 1;
        for (int32_t j = 0; j < k; ++j) Q *= card[par[poff[x] + j]];
        for (int64_t q = 0; q < Q; ++q) {
            bn::condition_t cond;
            for (int32_t j = 0; j < k; ++j) cond[parents[j]] = cfg[j];
            std::vector<double> row(cpt + coff[x] + q * r, cpt + coff[x] + (q + 1) * r);
            b.v[x]->cpt[cond].second = row;
            for (int32_t j = k - 1; j >= 0; --j) {
                if (++cfg[j] < card[par[poff[x] + j]]) break;
                cfg[j] = 0;
            }
        }
    }
}

} // namespace

extern "C" {

int bnref_is_capped(void)
{
#ifdef BNREF_CAPPED
    return 1;
#else
    return 0;
#endif
}

// out_marginals [n_cases][sum card]; out_sweeps / out_converged are filled only by the capped
// flavour (the pure one writes -1 / 1).  *bp_seconds = wall time spent inside operator().
int bnref_run(int32_t n, const int32_t* card, const int32_t* poff, const int32_t* par,
              const int64_t* coff, const double* cpt,
              int64_t n_cases, const int64_t* ev_off, const int32_t* ev_node,
              const int32_t* ev_state, const int64_t* ev_val_off, const double* ev_values,
              double eps, int32_t max_sweeps,
              double* out_marginals, int32_t* out_sweeps, uint8_t* out_converged,
              double* bp_seconds)
{
    built b;
    build(b, n, card, poff, par, coff, cpt);
    int64_t V = 0;
    std::vector<int64_t> voff(static_cast<std::size_t>(n) + 1, 0);
    for (int32_t i = 0; i < n; ++i) { voff[i + 1] = voff[i] + card[i]; }
    V = voff[n];
    bn::inference::belief_propagation bp(b.graph);
    double secs = 0;
    for (int64_t c = 0; c < n_cases; ++c) {
        std::unordered_map<bn::vertex_type, bn::matrix_type> pre;
        int64_t e0 = ev_off ? ev_off[c] : 0, e1 = ev_off ? ev_off[c + 1] : 0;
        for (int64_t e = e0; e < e1; ++e) {
            int32_t node = ev_node[e], r = card[node];
            bn::matrix_type m(1, static_cast<std::size_t>(r), 0.0);
            for (int32_t i = 0; i < r; ++i)
                m[0][i] = ev_values ? ev_values[ev_val_off[e] + i] : (ev_state[e] == i ? 1.0 : 0.0);
            pre[b.v[node]] = m;
        }
#ifdef BNREF_CAPPED
        g_cap = max_sweeps > 0 ? max_sweeps : (1 << 30);
#else
        (void)max_sweeps;
#endif
        auto t0 = std::chrono::steady_clock::now();
        auto result = bp(pre, eps);
        auto t1 = std::chrono::steady_clock::now();
        secs += std::chrono::duration<double>(t1 - t0).count();
        for (int32_t i = 0; i < n; ++i) {
            auto const& m = result.at(b.v[i]);
            for (int32_t s = 0; s < card[i]; ++s) out_marginals[c * V + voff[i] + s] = m[0][s];
        }
#ifdef BNREF_CAPPED
        bool broke = g_iter < g_cap;
        if (out_sweeps) out_sweeps[c] = broke ? g_iter + 1 : g_cap;
        if (out_converged) out_converged[c] = broke ? 1 : 0;
#else
        if (out_sweeps) out_sweeps[c] = -1;
        if (out_converged) out_converged[c] = 1;
#endif
    }
    if (bp_seconds) *bp_seconds = secs;
    return 0;
}

} // extern "C"
