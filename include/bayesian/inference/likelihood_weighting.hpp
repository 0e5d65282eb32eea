// bayesian/inference/likelihood_weighting.hpp — bn::inference::likelihood_weighting on B200.
//
// Drop-in for the reference class (godai0519/BayesianNetwork,
// bayesian/inference/likelihood_weighting.hpp:13-250): same constructor, same
// operator()(evidence_list, sample_num) and return type.  The samples are drawn and weighted by the
// CUDA kernel behind bnbp_lw_run_batch (<bnbp.h>, csrc/bnbp_lw.cuh); nothing is computed on the host.
// It exists here as the independent statistical check of the belief_propagation marginals on loopy
// networks.  Link with -lbnbp; there is no CPU fallback.
//
//   reference member (likelihood_weighting.hpp)                here
//   ---------------------------------------------------------  ------------------------------------------
//   likelihood_weighting(graph_t const&)               :20-23   keeps a (shallow) graph copy; flattens lazily
//   operator()(evidence_list const&, uint64 n = 10000) :28-59   one-case batch through bnbp_lw_run_batch
//   evidence_list = unordered_map<vertex_type, int>    :15      same
//   probability_generator (mt19937 <- random_device)   :227-246 counter-based variates keyed by `seed`
//                                                               (reproducible; set_seed / a fresh seed per call)
//   make_samples(...)                                  :62-118  not provided: it returns the host-side map of
//                                                               every distinct sample pattern, which has no
//                                                               device counterpart on this path
// New: operator()(std::vector<evidence_list> const&, n) -> std::vector<return_type> (one launch).
#ifndef BNB200_BAYESIAN_INFERENCE_LIKELIHOOD_WEIGHTING_HPP
#define BNB200_BAYESIAN_INFERENCE_LIKELIHOOD_WEIGHTING_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include <bayesian/graph.hpp>
#include <bayesian/matrix.hpp>
#include <bnbp.h>

namespace bn {
namespace inference {

class likelihood_weighting {
public:
    typedef std::unordered_map<vertex_type, int> evidence_list;
    typedef std::unordered_map<vertex_type, int> pattern_list;
    typedef std::unordered_map<vertex_type, matrix_type> return_type;

    explicit likelihood_weighting(graph_t const& graph) : graph_(graph) {}
    virtual ~likelihood_weighting() { release(); }
    likelihood_weighting(likelihood_weighting const&) = delete;
    likelihood_weighting& operator=(likelihood_weighting const&) = delete;

    // The reference draws a fresh random_device seed per object; here the seed is explicit and every
    // call advances it, so two calls on one object are independent and a re-run of a program repeats.
    void set_seed(std::uint64_t seed) { seed_ = seed; }

    return_type operator()(evidence_list const& evidence, std::uint64_t const sample_num = 10000)
    {
        std::vector<evidence_list> const one(1, evidence);
        return std::move(operator()(one, sample_num).front());
    }

    std::vector<return_type> operator()(std::vector<evidence_list> const& cases, std::uint64_t const sample_num = 10000)
    {
        sync_network();
        std::vector<vertex_type> const& nodes = graph_.vertex_list();
        std::unordered_map<vertex_type, std::int32_t> index;
        for (std::size_t i = 0; i < nodes.size(); ++i) index[nodes[i]] = static_cast<std::int32_t>(i);
        std::vector<std::int64_t> off(1, 0);
        std::vector<std::int32_t> node, state;
        for (evidence_list const& one : cases) {
            for (auto const& kv : one) {
                auto const it = index.find(kv.first);
                if (it == index.end()) throw std::invalid_argument("likelihood_weighting: evidence vertex is not in the graph");
                node.push_back(it->second);
                state.push_back(static_cast<std::int32_t>(kv.second));
            }
            off.push_back(static_cast<std::int64_t>(node.size()));
        }
        bnbp_evidence ev = bnbp_evidence();
        ev.n_cases = static_cast<std::int64_t>(cases.size());
        ev.ev_off = off.data();
        ev.ev_node = node.data();
        ev.ev_state = state.data();
        std::size_t V = 0;
        std::vector<std::size_t> offset(nodes.size() + 1, 0);
        for (std::size_t i = 0; i < nodes.size(); ++i) offset[i + 1] = offset[i] + static_cast<std::size_t>(flat_.card[i]);
        V = offset.back();
        std::vector<double> marg(cases.size() * V);
        if (bnbp_lw_run_batch(handle_, &ev, static_cast<std::int64_t>(sample_num), seed_++, marg.data(), nullptr) != BNBP_OK)
            throw std::runtime_error(std::string("bnbp_lw_run_batch: ") + bnbp_last_error());
        std::vector<return_type> out(cases.size());
        for (std::size_t c = 0; c < cases.size(); ++c)
            for (std::size_t i = 0; i < nodes.size(); ++i) {
                matrix_type& m = out[c][nodes[i]];
                m.resize(1, static_cast<std::size_t>(flat_.card[i]));
                m.assign(marg.data() + c * V + offset[i], marg.data() + c * V + offset[i + 1]);
            }
        return out;
    }

private:
    void release()
    {
        if (handle_) bnbp_destroy(handle_);
        handle_ = nullptr;
    }

    // the reference reads vertex->cpt at call time: re-flatten, re-upload what changed
    void sync_network()
    {
        flat_network now = flatten(graph_);
        if (handle_ && now.same_topology(flat_)) {
            if (now.cpt != flat_.cpt) {
                if (bnbp_refresh_cpt(handle_, now.cpt.data(), static_cast<std::int64_t>(now.cpt.size())) != BNBP_OK)
                    throw std::runtime_error(std::string("bnbp_refresh_cpt: ") + bnbp_last_error());
                flat_.cpt.swap(now.cpt);
            }
            return;
        }
        release();
        flat_ = std::move(now);
        bnbp_flat_network net;
        net.n_nodes = static_cast<std::int32_t>(flat_.card.size());
        net.card = flat_.card.data();
        net.parent_off = flat_.parent_off.data();
        net.parents = flat_.parents.data();
        net.cpt_off = flat_.cpt_off.data();
        net.cpt = flat_.cpt.data();
        bnbp_options bo = bnbp_options();
        bo.device = -1;
        bo.specialize = BNBP_SPEC_NEVER;          // no sweep kernels are needed for sampling
        if (bnbp_create(&net, &bo, &handle_) != BNBP_OK) {
            handle_ = nullptr;
            throw std::runtime_error(std::string("bnbp_create: ") + bnbp_last_error());
        }
    }

    graph_t const graph_;
    flat_network flat_;
    bnbp_handle* handle_ = nullptr;
    std::uint64_t seed_ = 0x5EED5EED5EEDull;
};

} // namespace inference
} // namespace bn

#endif // BNB200_BAYESIAN_INFERENCE_LIKELIHOOD_WEIGHTING_HPP
