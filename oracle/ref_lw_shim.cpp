// ref_lw_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-callable wrapper around the UNMODIFIED reference header bayesian/inference/likelihood_weighting.hpp,
// compiled where it lies (-I/root/reference; oracle/Makefile -> oracle/_ref/libbnref_lw.so, git-ignored).
// The reference seeds its generator from std::random_device, so its output differs run to run: the
// tests compare against it within sampling error only.
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <functional>
#include <memory>
#include <random>
#include <unordered_map>
#include <vector>

#include <bayesian/graph.hpp>
#include <bayesian/matrix.hpp>
#include <bayesian/inference/likelihood_weighting.hpp>

extern "C" int bnref_lw(int32_t n, const int32_t* card, const int32_t* poff, const int32_t* par,
                        const int64_t* coff, const double* cpt,
                        int64_t n_cases, const int64_t* ev_off, const int32_t* ev_node, const int32_t* ev_state,
                        int64_t n_samples, double* out)
{
    bn::graph_t graph;
    std::vector<bn::vertex_type> v;
    for (int32_t i = 0; i < n; ++i) {
        auto x = graph.add_vertex();
        x->id = i;
        x->selectable_num = static_cast<std::size_t>(card[i]);
        v.push_back(x);
    }
    for (int32_t i = 0; i < n; ++i)
        for (int32_t e = poff[i]; e < poff[i + 1]; ++e)
            if (!graph.add_edge(v[par[e]], v[i])) return 2;
    for (int32_t i = 0; i < n; ++i) {
        std::vector<bn::vertex_type> parents;
        for (int32_t e = poff[i]; e < poff[i + 1]; ++e) parents.push_back(v[par[e]]);
        v[i]->cpt.assign(parents, v[i]);
        int64_t rows = (coff[i + 1] - coff[i]) / card[i];
        for (int64_t q = 0; q < rows; ++q) {
            bn::condition_t cond;
            int64_t rem = q;
            for (int32_t e = poff[i + 1] - 1; e >= poff[i]; --e) {       // first parent slowest
                cond[v[par[e]]] = static_cast<int>(rem % card[par[e]]);
                rem /= card[par[e]];
            }
            std::vector<double> row(cpt + coff[i] + q * card[i], cpt + coff[i] + (q + 1) * card[i]);
            v[i]->cpt[cond].second = row;
        }
    }
    int64_t V = 0;
    std::vector<int64_t> voff(n + 1, 0);
    for (int32_t i = 0; i < n; ++i) voff[i + 1] = voff[i] + card[i];
    V = voff[n];
    bn::inference::likelihood_weighting lw(graph);
    for (int64_t c = 0; c < n_cases; ++c) {
        bn::inference::likelihood_weighting::evidence_list ev;
        for (int64_t e = ev_off[c]; e < ev_off[c + 1]; ++e) ev[v[ev_node[e]]] = ev_state[e];
        auto const res = lw(ev, static_cast<std::uint64_t>(n_samples));
        for (int32_t i = 0; i < n; ++i)
            for (int32_t s = 0; s < card[i]; ++s) out[c * V + voff[i] + s] = res.at(v[i])[0][s];
    }
    return 0;
}
