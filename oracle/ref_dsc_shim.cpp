// ref_dsc_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-callable wrapper around the reference's UNMODIFIED DSC loader (bayesian/serializer/dsc.hpp, needs
// no Boost), compiled where it lies (-I/root/reference; oracle/Makefile target `ref`).  It parses a
// DSC text with the reference and reads the resulting graph back through the reference's own
// accessors (vertex_list, in_vertexes, cpt[condition]) into the flat layout of include/bnbp.h, so
// tests can compare the new loader (include/bayesian/serializer/dsc.hpp) with the real thing.
// The reference's BIF loader cannot be built here (Boost.Spirit is absent).
#include <cassert>
#include <cstdint>
#include <string>
#include <vector>

#include <bayesian/graph.hpp>
#include <bayesian/serializer/dsc.hpp>

extern "C" {

// Two-pass: call with NULL arrays to get the sizes (n_nodes, n_edges, n_cpt), then with buffers.
int bnref_dsc_flatten(const char* text, int32_t* n_nodes, int32_t* n_edges, int64_t* n_cpt,
                      int32_t* card, int32_t* parent_off, int32_t* parents, int64_t* cpt_off, double* cpt)
{
    bn::serializer::dsc reader;
    bn::graph_t graph = reader.from_data(std::string(text));
    auto const& nodes = graph.vertex_list();
    int32_t edges = 0;
    int64_t values = 0;
    if (parent_off) parent_off[0] = 0;
    if (cpt_off) cpt_off[0] = 0;
    for (std::size_t i = 0; i < nodes.size(); ++i) {
        auto const ps = graph.in_vertexes(nodes[i]);
        std::size_t q_total = 1;
        for (auto const& p : ps) q_total *= p->selectable_num;
        if (card) card[i] = static_cast<int32_t>(nodes[i]->selectable_num);
        std::vector<int> cfg(ps.size(), 0);
        for (std::size_t q = 0; q < q_total; ++q) {
            bn::condition_t cond;
            for (std::size_t j = 0; j < ps.size(); ++j) cond[ps[j]] = cfg[j];
            auto const& row = nodes[i]->cpt[cond].second;
            if (cpt)
                for (std::size_t x = 0; x < nodes[i]->selectable_num; ++x) cpt[values + static_cast<int64_t>(x)] = row.at(x);
            values += static_cast<int64_t>(nodes[i]->selectable_num);
            for (std::size_t j = ps.size(); j-- > 0;) {
                if (++cfg[j] < static_cast<int>(ps[j]->selectable_num)) break;
                cfg[j] = 0;
            }
        }
        for (auto const& p : ps) {
            if (parents) {
                for (std::size_t k = 0; k < nodes.size(); ++k)
                    if (nodes[k] == p) parents[edges] = static_cast<int32_t>(k);
            }
            ++edges;
        }
        if (parent_off) parent_off[i + 1] = edges;
        if (cpt_off) cpt_off[i + 1] = values;
    }
    *n_nodes = static_cast<int32_t>(nodes.size());
    *n_edges = edges;
    *n_cpt = values;
    return 0;
}

} // extern "C"
