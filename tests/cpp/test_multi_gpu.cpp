// The drop-in class over SEVERAL GPUs of one box (options::devices -> bnbp_create_multi): one run_flat call shards
// its cases over every visible device and must return, bit for bit, what one device returns; the convergence
// summary comes back all-reduced over the library's NCCL communicator.  Also the "what leaves the device"
// options of run_flat (query vertices, float marginals) and the page-locked result storage.
// Runs on a B200 box (any GPU count: with one GPU the group has one member and the same code path runs).
#define BOOST_TEST_MAIN
#include <boost/test/included/unit_test.hpp>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "bayesian/graph.hpp"
#include "bayesian/inference/belief_propagation.hpp"

namespace {

struct rng64 {                                       // splitmix64
    unsigned long long s;
    unsigned long long next() { unsigned long long z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// a loopy 30-node network: node i has up to 3 parents among its predecessors, cardinalities 2..4
bn::graph_t make_network(std::vector<bn::vertex_type>* nodes)
{
    bn::graph_t g;
    rng64 r{20261018};
    int const n = 30;
    for (int i = 0; i < n; ++i) { nodes->push_back(g.add_vertex()); (*nodes)[i]->selectable_num = 2 + r.next() % 3; }
    for (int i = 1; i < n; ++i) {
        int const k = (int)(r.next() % 4);
        for (int j = 0; j < k && j < i; ++j) g.add_edge((*nodes)[r.next() % i], (*nodes)[i]);   // duplicates are refused by add_edge
    }
    for (int i = 0; i < n; ++i) {
        bn::vertex_type const& v = (*nodes)[i];
        std::vector<bn::vertex_type> const parents = g.in_vertexes(v);
        v->cpt.assign(parents, v);
        for (auto const& cond : v->cpt.pattern()) {
            std::vector<double> row(v->selectable_num);
            double sum = 0;
            for (double& x : row) { x = 0.05 + 0.95 * r.unit(); sum += x; }
            for (double& x : row) x /= sum;
            v->cpt[cond].second = row;
        }
    }
    return g;
}

struct evidence_csr {
    std::vector<std::int64_t> off;
    std::vector<std::int32_t> node, state;
    bnbp_evidence view() const
    {
        bnbp_evidence ev;
        ev.n_cases = (std::int64_t)off.size() - 1;
        ev.ev_off = off.data(); ev.ev_node = node.data(); ev.ev_state = state.data();
        ev.ev_val_off = nullptr; ev.ev_values = nullptr;
        return ev;
    }
};

evidence_csr make_evidence(std::vector<bn::vertex_type> const& nodes, std::size_t n_cases)
{
    evidence_csr e;
    rng64 r{7};
    e.off.push_back(0);
    for (std::size_t c = 0; c < n_cases; ++c) {
        for (std::size_t i = 0; i < nodes.size(); ++i)
            if (r.next() % 10 == 0) { e.node.push_back((std::int32_t)i); e.state.push_back((std::int32_t)(r.next() % nodes[i]->selectable_num)); }
        e.off.push_back((std::int64_t)e.node.size());
    }
    return e;
}

} // namespace

BOOST_AUTO_TEST_CASE(all_devices_equal_one_device_bit_for_bit)
{
    std::vector<bn::vertex_type> nodes;
    bn::graph_t g = make_network(&nodes);
    evidence_csr const e = make_evidence(nodes, 20011);             // not a multiple of anything
    bnbp_evidence const ev = e.view();
    int const n_dev = bnbp_device_count();
    BOOST_CHECK(n_dev >= 1);
    std::printf("devices visible: %d\n", n_dev);

    for (int mode = 0; mode < 2; ++mode) {                          // fixed sweep count, then the reference's stopping rule
        bn::inference::belief_propagation::options one;
        one.epsilon = mode ? 1e-7 : 0.0;
        one.max_sweeps = mode ? 200 : 15;
        bn::inference::belief_propagation bp1(g);
        auto const a = bp1.run_flat(ev, one);

        bn::inference::belief_propagation::options all = one;
        all.devices = {-1};                                         // every visible device
        bn::inference::belief_propagation bpn(g);
        auto const b = bpn.run_flat(ev, all);

        BOOST_CHECK(a.n_cases == b.n_cases && a.values_per_case == b.values_per_case);
        BOOST_CHECK(a.marginals.size() == b.marginals.size());
        BOOST_CHECK(std::memcmp(a.marginals.data(), b.marginals.data(), a.marginals.size() * sizeof(double)) == 0);
        BOOST_CHECK(std::memcmp(a.sweeps.data(), b.sweeps.data(), a.sweeps.size() * sizeof(std::int32_t)) == 0);
        BOOST_CHECK(std::memcmp(a.converged.data(), b.converged.data(), a.converged.size()) == 0);
        // the all-reduced summary equals the totals of the per-case arrays
        long long sweeps = 0, open = 0, mx = 0;
        for (std::size_t c = 0; c < a.n_cases; ++c) { sweeps += a.sweeps[c]; open += a.converged[c] ? 0 : 1; mx = a.sweeps[c] > mx ? a.sweeps[c] : mx; }
        BOOST_CHECK(b.summary.n_cases == (long long)a.n_cases);
        BOOST_CHECK(b.summary.case_sweeps == sweeps);
        BOOST_CHECK(b.summary.not_converged == open);
        BOOST_CHECK(b.summary.max_sweeps == mx);
        std::printf("mode %d: %zu cases on %d device(s), %lld case-sweeps, %lld not converged, max %lld sweeps\n", mode, a.n_cases,
                    n_dev, (long long)b.summary.case_sweeps, (long long)b.summary.not_converged, (long long)b.summary.max_sweeps);

        if (n_dev >= 2) {                                           // an explicit subset, fewer cases than devices * 32
            bn::inference::belief_propagation::options two = one;
            two.devices = {1, 0};
            bn::inference::belief_propagation bp2(g);
            evidence_csr const small = make_evidence(nodes, 37);
            auto const s1 = bp1.run_flat(small.view(), one);
            auto const s2 = bp2.run_flat(small.view(), two);
            BOOST_CHECK(std::memcmp(s1.marginals.data(), s2.marginals.data(), s1.marginals.size() * sizeof(double)) == 0);
        }
    }
}

BOOST_AUTO_TEST_CASE(query_vertices_and_float_marginals)
{
    std::vector<bn::vertex_type> nodes;
    bn::graph_t g = make_network(&nodes);
    evidence_csr const e = make_evidence(nodes, 9000);              // >= 4096: the on-chip kernel where the network is eligible
    bnbp_evidence const ev = e.view();
    bn::inference::belief_propagation::options opt;
    opt.epsilon = 0.0;
    opt.max_sweeps = 12;
    bn::inference::belief_propagation bp(g);
    auto const full = bp.run_flat(ev, opt);

    bn::inference::belief_propagation::options q = opt;
    q.query = {nodes[17], nodes[3], nodes[29]};                     // any order
    q.devices = {-1};
    bn::inference::belief_propagation bpq(g);
    auto const part = bpq.run_flat(ev, q);
    BOOST_CHECK(part.vertex.size() == 3 && part.vertex[0] == 17 && part.vertex[1] == 3 && part.vertex[2] == 29);
    BOOST_CHECK(part.values_per_case == nodes[17]->selectable_num + nodes[3]->selectable_num + nodes[29]->selectable_num);
    bool same = true;
    for (std::size_t c = 0; c < full.n_cases && same; ++c)
        for (std::size_t i = 0; i < part.vertex.size() && same; ++i)
            for (std::size_t x = 0; x < nodes[part.vertex[i]]->selectable_num; ++x)
                same = same && part.marginals[c * part.values_per_case + part.offset[i] + x] ==
                                   full.marginals[c * full.values_per_case + full.offset[part.vertex[i]] + x];
    BOOST_CHECK(same);

    // float marginals need an fp32 handle; they equal the double marginals of the same handle rounded to float
    bn::inference::belief_propagation::options f = opt;
    f.precision = BNBP_FP32;
    bn::inference::belief_propagation bpf(g);
    auto const wide = bpf.run_flat(ev, f);
    f.float_marginals = true;
    auto const narrow = bpf.run_flat(ev, f);
    BOOST_CHECK(narrow.marginals.empty() && narrow.marginals_f32.size() == wide.marginals.size());
    bool close = true;
    for (std::size_t i = 0; i < wide.marginals.size() && close; ++i)
        close = std::fabs((double)narrow.marginals_f32[i] - wide.marginals[i]) <= 1e-6 * std::fabs(wide.marginals[i]) + 1e-7;
    BOOST_CHECK(close);
    for (std::size_t i = 0; i < wide.marginals.size() && close; ++i)
        close = std::fabs(full.marginals[i] - wide.marginals[i]) <= 1e-5 * std::fabs(full.marginals[i]) + 1e-7;
    BOOST_CHECK(close);
    opt.float_marginals = true;                                     // fp64 handle: refused
    bool threw = false;
    try { bp.run_flat(ev, opt); } catch (std::runtime_error const&) { threw = true; }
    BOOST_CHECK(threw);
}
