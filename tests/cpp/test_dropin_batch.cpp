// Own tests of the drop-in headers beyond what the reference's test files cover: full-precision
// golden vectors (produced by the unmodified reference), the batched
// overloads, condition_t evidence, CPT edits between calls, topology edits, error behaviour.
// Runs on a B200 (calls libbnbp).  Uses the same Boost.Test stand-in as the reference tests.
#define BOOST_TEST_MAIN
#include <boost/test/included/unit_test.hpp>

#include <cmath>
#include <stdexcept>

#include "bayesian/graph.hpp"
#include "bayesian/inference/belief_propagation.hpp"

namespace {

bn::graph_t pearl()
{
    bn::graph_t g;
    auto r = g.add_vertex(), s = g.add_vertex(), w = g.add_vertex(), h = g.add_vertex();
    g.add_edge(r, w); g.add_edge(r, h); g.add_edge(s, h);
    for (auto const& v : g.vertex_list()) v->selectable_num = 2;
    r->cpt.assign({}, r); r->cpt[bn::condition_t()].second = {0.2, 0.8};
    s->cpt.assign({}, s); s->cpt[bn::condition_t()].second = {0.1, 0.9};
    w->cpt.assign({r}, w);
    w->cpt[{{r, 0}}].second = {1.0, 0.0};
    w->cpt[{{r, 1}}].second = {0.2, 0.8};
    h->cpt.assign({s, r}, h);                         // parents listed in the OTHER order on purpose
    h->cpt[{{r, 0}, {s, 0}}].second = {1.0, 0.0};
    h->cpt[{{r, 0}, {s, 1}}].second = {1.0, 0.0};
    h->cpt[{{r, 1}, {s, 0}}].second = {0.9, 0.1};
    h->cpt[{{r, 1}, {s, 1}}].second = {0.0, 1.0};
    return g;
}

bn::graph_t chain()
{
    bn::graph_t g;
    auto a = g.add_vertex(), b = g.add_vertex(), c = g.add_vertex(), d = g.add_vertex();
    g.add_edge(a, b); g.add_edge(b, c); g.add_edge(c, d);
    a->selectable_num = 3; b->selectable_num = 3; c->selectable_num = 2; d->selectable_num = 3;
    a->cpt.assign({}, a); a->cpt[bn::condition_t()].second = {0.30, 0.60, 0.10};
    b->cpt.assign({a}, b);
    b->cpt[{{a, 0}}].second = {0.20, 0.30, 0.50};
    b->cpt[{{a, 1}}].second = {0.30, 0.30, 0.40};
    b->cpt[{{a, 2}}].second = {0.80, 0.10, 0.10};
    c->cpt.assign({b}, c);
    c->cpt[{{b, 0}}].second = {0.50, 0.50};
    c->cpt[{{b, 1}}].second = {0.70, 0.30};
    c->cpt[{{b, 2}}].second = {0.40, 0.60};
    d->cpt.assign({c}, d);
    d->cpt[{{c, 0}}].second = {0.40, 0.30, 0.30};
    d->cpt[{{c, 1}}].second = {0.20, 0.60, 0.20};
    return g;
}

void check_row(bn::matrix_type const& m, std::vector<double> const& want, double rel = 1e-12)
{
    BOOST_CHECK(m.height() == 1);
    BOOST_CHECK(m.width() == want.size());
    for (std::size_t i = 0; i < want.size() && i < m.width(); ++i)
        BOOST_CHECK(std::fabs(m[0][i] - want[i]) <= rel * std::fabs(want[i]) + 1e-15);
}

} // namespace

BOOST_AUTO_TEST_CASE(golden_pearl_full_precision)
{
    bn::graph_t g = pearl();
    auto const v = g.vertex_list();
    bn::inference::belief_propagation bp(g);
    auto const none = bp();
    check_row(none.at(v[0]), {0.2, 0.8});
    check_row(none.at(v[1]), {0.1, 0.9});
    check_row(none.at(v[2]), {0.36, 0.64});
    check_row(none.at(v[3]), {0.272, 0.728});
    std::unordered_map<bn::vertex_type, bn::matrix_type> pre;
    pre[v[3]].resize(1, 2, 0);
    pre[v[3]][0][0] = 1;
    auto const res = bp(pre);                         // belief_propagation.cpp:92-121 at 17 digits
    check_row(res.at(v[0]), {0.73529411764705888, 0.26470588235294118});
    check_row(res.at(v[1]), {0.33823529411764708, 0.66176470588235292});
    check_row(res.at(v[2]), {0.78823529411764715, 0.21176470588235297});
    check_row(res.at(v[3]), {1.0, 0.0});
}

BOOST_AUTO_TEST_CASE(alias_and_condition_evidence)
{
    bn::graph_t g = chain();
    auto const v = g.vertex_list();
    bn::inference::loopy_belief_propagation bp(g);    // the north-star name
    bn::condition_t hard = {{v[3], 2}};               // D = 2  (resume_sample3)
    auto const res = bp(hard, 0.001);
    check_row(res.at(v[0]), {0.2992845786963434, 0.60095389507154218, 0.099761526232114497});
    check_row(res.at(v[1]), {0.31796502384737679, 0.30047694753577109, 0.38155802861685217});
    check_row(res.at(v[2]), {0.61526232114467405, 0.38473767885532589});
    check_row(res.at(v[3]), {0.0, 0.0, 1.0});
}

BOOST_AUTO_TEST_CASE(soft_evidence_goes_into_pi_and_lambda)
{
    bn::graph_t g = chain();
    auto const v = g.vertex_list();
    bn::inference::belief_propagation bp(g);
    std::unordered_map<bn::vertex_type, bn::matrix_type> pre;
    pre[v[2]].resize(1, 2);
    pre[v[2]][0] = {0.3, 0.7};
    auto const res = bp(pre);                         // belief = normalize(e .* e), :69-73,:151-158
    check_row(res.at(v[2]), {0.15517241379310345, 0.84482758620689646});
}

BOOST_AUTO_TEST_CASE(batched_overloads_match_single_calls)
{
    bn::graph_t g = chain();
    auto const v = g.vertex_list();
    bn::inference::belief_propagation bp(g);
    std::vector<bn::condition_t> cases = {{}, {{v[0], 0}}, {{v[2], 1}}, {{v[0], 1}, {v[2], 1}}, {{v[3], 2}}};
    bn::inference::belief_propagation::options opt;
    auto const batch = bp(cases, opt);
    BOOST_CHECK(batch.size() == cases.size());
    for (std::size_t c = 0; c < cases.size(); ++c) {
        auto const single = bp(cases[c], opt.epsilon);
        for (auto const& node : v) check_row(batch[c].at(node), single.at(node)[0], 0.0);
    }
    check_row(batch[1].at(v[1]), {0.2, 0.3, 0.5});                                 // resume_sample4
    check_row(batch[3].at(v[1]), {0.3125, 0.1875, 0.5});                           // resume_sample2
    // flat form: sweeps and convergence flags per case
    std::vector<std::int64_t> off = {0, 0, 1};
    std::vector<std::int32_t> node = {3}, state = {2};
    bnbp_evidence ev = {2, off.data(), node.data(), state.data(), nullptr, nullptr};
    auto const flat = bp.run_flat(ev, opt);
    BOOST_CHECK(flat.n_cases == 2 && flat.values_per_case == 11);
    BOOST_CHECK(flat.sweeps[0] == 6 && flat.sweeps[1] == 6);                       // resume_sample3: SURVEY section 4 table
    BOOST_CHECK(flat.converged[0] == 1 && flat.converged[1] == 1);
}

BOOST_AUTO_TEST_CASE(cpt_and_topology_edits_are_seen_by_the_next_call)
{
    bn::graph_t g = pearl();
    auto const v = g.vertex_list();
    bn::inference::belief_propagation bp(g);
    check_row(bp().at(v[2]), {0.36, 0.64});
    v[0]->cpt[bn::condition_t()].second = {0.5, 0.5};            // edit a prior: shared vertex objects
    check_row(bp().at(v[2]), {0.6, 0.4});
    check_row(bp().at(v[0]), {0.5, 0.5});
}

BOOST_AUTO_TEST_CASE(clone_is_usable_for_inference)
{
    bn::graph_t g = pearl();
    bn::graph_t c = g.clone();
    BOOST_CHECK(c.vertex_list().size() == 4 && c.edge_list().size() == 3);
    BOOST_CHECK(c.vertex_list()[0] != g.vertex_list()[0]);
    bn::inference::belief_propagation bp(c);
    check_row(bp().at(c.vertex_list()[3]), {0.272, 0.728});
}

BOOST_AUTO_TEST_CASE(errors_are_exceptions_not_undefined_behaviour)
{
    bn::graph_t g = pearl();
    auto const v = g.vertex_list();
    bn::inference::belief_propagation bp(g);
    std::unordered_map<bn::vertex_type, bn::matrix_type> bad;
    bad[v[3]].resize(1, 3, 1.0);                                  // wrong width
    bool threw = false;
    try { bp(bad); } catch (std::invalid_argument const&) { threw = true; }
    BOOST_CHECK(threw);
    bn::condition_t out_of_range = {{v[3], 7}};
    threw = false;
    try { bp(out_of_range, 0.001); } catch (std::runtime_error const&) { threw = true; }
    BOOST_CHECK(threw);
    check_row(bp().at(v[3]), {0.272, 0.728});                     // still usable afterwards
    // a vertex without a CPT row for some parent configuration: the reference has UB (graph.hpp:120-124)
    bn::graph_t h;
    auto a = h.add_vertex(), b = h.add_vertex();
    h.add_edge(a, b);
    a->selectable_num = 2; b->selectable_num = 2;
    a->cpt.assign({}, a); a->cpt[bn::condition_t()].second = {0.5, 0.5};
    bn::inference::belief_propagation broken(h);                  // b has no CPT at all
    threw = false;
    try { broken(); } catch (std::invalid_argument const&) { threw = true; }
    BOOST_CHECK(threw);
}

BOOST_AUTO_TEST_CASE(impossible_evidence_gives_nan_like_the_reference)
{
    bn::graph_t g = pearl();
    auto const v = g.vertex_list();
    bn::inference::belief_propagation bp(g);
    bn::inference::belief_propagation::options opt;
    opt.epsilon = 0.0;                                            // fixed sweep count: with eps = 1e-3 the case stops
    opt.max_sweeps = 6;                                           // after 2 sweeps, before the NaN reaches a belief
    std::vector<bn::condition_t> cases = {{{v[0], 0}, {v[2], 1}}};   // R = 0 and W = 1: P = 0
    auto const res = bp(cases, opt);
    bool any_nan = false;
    for (auto const& node : v)
        for (double x : res[0].at(node)[0]) any_nan = any_nan || std::isnan(x);
    BOOST_CHECK(any_nan);
}
