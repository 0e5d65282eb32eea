// bayesian/sampler.hpp — bn::sampler on B200: the CPT-estimation step of the reference workflow
// (sample table -> CPTs -> inference), with the counting done by the CUDA kernels behind
// bnbp_estimate_cpt (<bnbp.h>, csrc/bnbp_lw.cuh).  SURVEY 8 f3.
//
// Drop-in for the part of the reference class that feeds the inference path (godai0519/BayesianNetwork,
// bayesian/sampler.hpp:16-225):
//   reference member (sampler.hpp)                                   here
//   ---------------------------------------------------------------  --------------------------------------
//   sampler() / sampler(std::string const& filename)        :18-27   same
//   load_sample(unordered_map<condition_t, size_t> const&)  :29-37   same (table of distinct samples)
//   load_sample(std::vector<vertex_type> const& node_list)  :42-76   same file format: one line per distinct
//                                                                    sample, "count s_0 s_1 ...", whitespace
//                                                                    separated (plain iostreams, no Boost)
//   make_cpt(graph_t const&) const                          :81-163  counts on the GPU; rows written back
//                                                                    into vertex->cpt exactly as the reference
//                                                                    does (count / row total, uniform 1/r for
//                                                                    parent configurations never seen)
//   filename / set_filename / table / sampling_size         :166-193 same
// Link with -lbnbp; there is no CPU fallback for make_cpt.
#ifndef BNB200_BAYESIAN_SAMPLER_HPP
#define BNB200_BAYESIAN_SAMPLER_HPP

#include <cstddef>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include <bayesian/graph.hpp>
#include <bnbp.h>

namespace bn {

class sampler {
public:
    sampler() : filename_(), table_(), sampling_size_(0) {}
    sampler(std::string const& filename) : filename_(filename), table_(), sampling_size_(0) {}

    bool load_sample(std::unordered_map<condition_t, std::size_t> const& table)
    {
        table_ = table;
        sampling_size_ = 0;
        for (auto const& p : table) sampling_size_ += p.second;
        return true;
    }

    bool load_sample(std::vector<vertex_type> const& node_list)
    {
        std::ifstream ifs(filename_);
        if (!ifs.is_open()) return false;
        std::size_t total = 0;
        std::unordered_map<condition_t, std::size_t> table;
        std::string line;
        while (std::getline(ifs, line)) {
            std::istringstream in(line);
            long long count = 0;
            if (!(in >> count)) continue;                      // blank line
            condition_t sample;
            for (vertex_type const& v : node_list) {
                int s = 0;
                if (!(in >> s)) throw std::invalid_argument("sampler::load_sample: short line in " + filename_);
                sample[v] = s;
            }
            table[sample] += static_cast<std::size_t>(count);
            total += static_cast<std::size_t>(count);
        }
        sampling_size_ = total;
        table_ = std::move(table);
        return true;
    }

    bool make_cpt(graph_t const& graph) const
    {
        if (sampling_size() == 0) return false;
        std::vector<vertex_type> const& nodes = graph.vertex_list();
        for (vertex_type const& node : nodes) node->cpt.assign(graph.in_vertexes(node), node);
        flat_network flat = flatten_topology(graph);
        std::size_t const N = nodes.size();
        std::vector<std::int32_t> rows;
        std::vector<std::int64_t> mult;
        rows.reserve(table_.size() * N);
        mult.reserve(table_.size());
        for (auto const& sample : table_) {
            for (vertex_type const& v : nodes) rows.push_back(static_cast<std::int32_t>(sample.first.at(v)));
            mult.push_back(static_cast<std::int64_t>(sample.second));
        }
        bnbp_flat_network net;
        net.n_nodes = static_cast<std::int32_t>(N);
        net.card = flat.card.data();
        net.parent_off = flat.parent_off.data();
        net.parents = flat.parents.data();
        net.cpt_off = flat.cpt_off.data();
        net.cpt = nullptr;
        std::vector<double> cpt(static_cast<std::size_t>(flat.cpt_off.back()));
        if (bnbp_estimate_cpt(&net, rows.data(), mult.data(), static_cast<std::int64_t>(mult.size()), -1, cpt.data()) != BNBP_OK)
            throw std::runtime_error(std::string("bnbp_estimate_cpt: ") + bnbp_last_error());
        // write the rows back: configuration q of node i, first parent slowest (the arena's order)
        for (std::size_t i = 0; i < N; ++i) {
            std::vector<std::size_t> const pidx = graph.parent_indices(i);
            std::size_t const r = nodes[i]->selectable_num;
            std::size_t const n_rows = static_cast<std::size_t>(flat.cpt_off[i + 1] - flat.cpt_off[i]) / r;
            for (std::size_t q = 0; q < n_rows; ++q) {
                condition_t cond;
                std::size_t rem = q;
                for (std::size_t j = pidx.size(); j-- > 0;) {
                    std::size_t const rp = nodes[pidx[j]]->selectable_num;
                    cond[nodes[pidx[j]]] = static_cast<int>(rem % rp);
                    rem /= rp;
                }
                double const* src = cpt.data() + flat.cpt_off[i] + q * r;
                nodes[i]->cpt[cond].second.assign(src, src + r);
            }
        }
        return true;
    }

    std::string filename() const { return filename_; }
    void set_filename(std::string const& filename)
    {
        filename_ = filename;
        sampling_size_ = 0;
        table_.clear();
    }
    std::unordered_map<condition_t, std::size_t> table() const { return table_; }
    std::size_t sampling_size() const { return sampling_size_; }

private:
    // topology + CPT offsets only (the CPT values are what make_cpt is about to produce)
    static flat_network flatten_topology(graph_t const& graph)
    {
        std::vector<vertex_type> const& nodes = graph.vertex_list();
        flat_network f;
        f.parent_off.push_back(0);
        f.cpt_off.push_back(0);
        for (std::size_t i = 0; i < nodes.size(); ++i) {
            if (nodes[i]->selectable_num == 0) throw std::invalid_argument("sampler::make_cpt: vertex with selectable_num == 0");
            f.card.push_back(static_cast<std::int32_t>(nodes[i]->selectable_num));
            std::size_t configurations = 1;
            for (std::size_t const p : graph.parent_indices(i)) {
                f.parents.push_back(static_cast<std::int32_t>(p));
                configurations *= nodes[p]->selectable_num;
            }
            f.parent_off.push_back(static_cast<std::int32_t>(f.parents.size()));
            f.cpt_off.push_back(f.cpt_off.back() + static_cast<std::int64_t>(configurations * nodes[i]->selectable_num));
        }
        return f;
    }

    std::string filename_;
    std::unordered_map<condition_t, std::size_t> table_;
    std::size_t sampling_size_;
};

} // namespace bn

#endif // BNB200_BAYESIAN_SAMPLER_HPP
