// ref_sampler_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-callable wrapper around the reference's own bayesian/sampler.hpp (sampler::load_sample :42-76,
// sampler::make_cpt :81-163), compiled where it lies (-I/root/reference; oracle/Makefile `ref`).  The header
// needs two Boost headers this image lacks; the stand-ins under tests/cpp/boost/ provide exactly the calls it
// makes (boost::algorithm::split with token_compress_on; boost::optional is included but unused).  Nothing from
// the reference is copied: the output is oracle/_ref/libbnref_sampler.so (git-ignored).
// Pins bnbp_estimate_cpt / include/bayesian/sampler.hpp against a reference binary.
#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

#include <bayesian/graph.hpp>
#include <bayesian/sampler.hpp>

namespace {

struct built {
    bn::graph_t graph;
    std::vector<bn::vertex_type> v;
};

void build_topology(built& b, int32_t n, const int32_t* card, const int32_t* poff, const int32_t* par)
{
    for (int32_t i = 0; i < n; ++i) {
        auto v = b.graph.add_vertex();
        v->id = i;
        v->selectable_num = static_cast<std::size_t>(card[i]);
        b.v.push_back(v);
    }
    for (int32_t x = 0; x < n; ++x)
        for (int32_t e = poff[x]; e < poff[x + 1]; ++e) b.graph.add_edge(b.v[par[e]], b.v[x]);
}

// CPTs of the graph into the flat layout of include/bnbp.h (first parent slowest, parents = in_vertexes order)
void read_back(built const& b, int32_t n, const int32_t* card, const int32_t* poff, const int32_t* par, const int64_t* coff,
               double* out)
{
    for (int32_t x = 0; x < n; ++x) {
        int32_t const k = poff[x + 1] - poff[x], r = card[x];
        std::vector<int> cfg(static_cast<std::size_t>(k), 0);
        int64_t Q = 1;
        for (int32_t j = 0; j < k; ++j) Q *= card[par[poff[x] + j]];
        for (int64_t q = 0; q < Q; ++q) {
            bn::condition_t cond;
            for (int32_t j = 0; j < k; ++j) cond[b.v[par[poff[x] + j]]] = cfg[j];
            auto const& row = b.v[x]->cpt[cond].second;
            for (int32_t i = 0; i < r; ++i) out[coff[x] + q * r + i] = row.at(static_cast<std::size_t>(i));
            for (int32_t j = k - 1; j >= 0; --j) {
                if (++cfg[j] < card[par[poff[x] + j]]) break;
                cfg[j] = 0;
            }
        }
    }
}

} // namespace

extern "C" {

// samples[row][node], multiplicity[row] -> sampler::load_sample(table) + sampler::make_cpt(graph)
int bnref_make_cpt(int32_t n, const int32_t* card, const int32_t* poff, const int32_t* par, const int64_t* coff,
                   const int32_t* samples, const int64_t* mult, int64_t n_rows, double* out_cpt)
{
    built b;
    build_topology(b, n, card, poff, par);
    std::unordered_map<bn::condition_t, std::size_t> table;
    for (int64_t r = 0; r < n_rows; ++r) {
        bn::condition_t s;
        for (int32_t i = 0; i < n; ++i) s[b.v[i]] = samples[r * n + i];
        table[s] += static_cast<std::size_t>(mult ? mult[r] : 1);
    }
    bn::sampler smp;
    if (!smp.load_sample(table)) return 1;
    if (!smp.make_cpt(b.graph)) return 2;
    read_back(b, n, card, poff, par, coff, out_cpt);
    return 0;
}

// the reference's own file reader: "count v1 v2 ... vN" per line (sampler.hpp:42-76), then make_cpt
int bnref_make_cpt_from_file(int32_t n, const int32_t* card, const int32_t* poff, const int32_t* par, const int64_t* coff,
                             const char* path, int64_t* sampling_size, double* out_cpt)
{
    built b;
    build_topology(b, n, card, poff, par);
    bn::sampler smp{std::string(path)};
    if (!smp.load_sample(b.v)) return 1;
    if (sampling_size) *sampling_size = static_cast<int64_t>(smp.sampling_size());
    if (!smp.make_cpt(b.graph)) return 2;
    read_back(b, n, card, poff, par, coff, out_cpt);
    return 0;
}

} // extern "C"
