// e2e_timing — the reference-facing boundary timed from C++: bn::inference::belief_propagation::run_flat on a
// network and an evidence batch that bench.py wrote to a binary file, K timed calls, one JSON line on stdout.
// This is the call a user of the drop-in makes: the graph is built through the public graph_t / cpt_t API,
// flattened and uploaded by the class, and the results come back in page-locked memory (pinned_buffer).
//   e2e_timing <file> <calls> <warmup> <precision fp64|fp32> <float_marginals 0|1> <devices: -1 all | 0 one> [query count]
// File layout (little endian): int64 N, E, n_cpt, n_cases, nnz, sweeps; int32 card[N], parent_off[N+1], parents[E];
// int64 cpt_off[N+1]; double cpt[n_cpt]; int64 ev_off[n_cases+1]; int32 ev_node[nnz], ev_state[nnz].
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bayesian/graph.hpp"
#include "bayesian/inference/belief_propagation.hpp"

template <class T> static bool read_vec(FILE* f, std::vector<T>& v, std::size_t n)
{
    v.resize(n);
    return n == 0 || std::fread(v.data(), sizeof(T), n, f) == n;
}

int main(int argc, char** argv)
{
    if (argc < 7) { std::fprintf(stderr, "usage: e2e_timing file calls warmup fp64|fp32 float_marginals devices [n_query]\n"); return 2; }
    int const calls = std::atoi(argv[2]), warmup = std::atoi(argv[3]);
    bool const fp32 = std::string(argv[4]) == "fp32";
    bool const float_out = std::atoi(argv[5]) != 0;
    int const devices = std::atoi(argv[6]);
    int const n_query = argc > 7 ? std::atoi(argv[7]) : 0;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    std::int64_t hdr[6];
    if (std::fread(hdr, 8, 6, f) != 6) return 2;
    std::size_t const N = hdr[0], E = hdr[1], n_cpt = hdr[2], n_cases = hdr[3], nnz = hdr[4];
    int const sweeps = (int)hdr[5];
    std::vector<std::int32_t> card, poff, par, ev_node, ev_state;
    std::vector<std::int64_t> coff, ev_off;
    std::vector<double> cpt;
    if (!read_vec(f, card, N) || !read_vec(f, poff, N + 1) || !read_vec(f, par, E) || !read_vec(f, coff, N + 1) ||
        !read_vec(f, cpt, n_cpt) || !read_vec(f, ev_off, n_cases + 1) || !read_vec(f, ev_node, nnz) || !read_vec(f, ev_state, nnz)) {
        std::fprintf(stderr, "short file\n");
        return 2;
    }
    std::fclose(f);

    // the network through the reference's public API (graph.hpp:251-291, cpt_t::assign :490-505)
    bn::graph_t g;
    std::vector<bn::vertex_type> v;
    for (std::size_t i = 0; i < N; ++i) { v.push_back(g.add_vertex()); v[i]->selectable_num = (std::size_t)card[i]; }
    for (std::size_t i = 0; i < N; ++i)
        for (std::int32_t e = poff[i]; e < poff[i + 1]; ++e) g.add_edge(v[(std::size_t)par[e]], v[i]);
    for (std::size_t i = 0; i < N; ++i) {
        std::vector<bn::vertex_type> parents;
        for (std::int32_t e = poff[i]; e < poff[i + 1]; ++e) parents.push_back(v[(std::size_t)par[e]]);
        v[i]->cpt.assign(parents, v[i]);
        std::size_t const k = parents.size(), r = (std::size_t)card[i];
        std::size_t const Q = (std::size_t)(coff[i + 1] - coff[i]) / r;
        for (std::size_t q = 0; q < Q; ++q) {                       // first parent slowest (all_combination_pattern :269-295)
            bn::condition_t cond;
            std::size_t rest = q;
            for (std::size_t j = k; j-- > 0;) { cond[parents[j]] = (int)(rest % (std::size_t)card[(std::size_t)par[poff[i] + (std::int32_t)j]]); rest /= (std::size_t)card[(std::size_t)par[poff[i] + (std::int32_t)j]]; }
            v[i]->cpt[cond].second.assign(cpt.begin() + coff[i] + (std::int64_t)(q * r), cpt.begin() + coff[i] + (std::int64_t)((q + 1) * r));
        }
    }

    bn::inference::belief_propagation::options opt;
    opt.epsilon = 0.0;
    opt.max_sweeps = sweeps;
    opt.precision = fp32 ? BNBP_FP32 : BNBP_FP64;
    opt.float_marginals = float_out;
    if (devices < 0) opt.devices = {-1};
    for (int i = 0; i < n_query && (std::size_t)i < N; ++i) opt.query.push_back(v[(std::size_t)i * (N / (std::size_t)n_query)]);
    bnbp_evidence ev;
    ev.n_cases = (std::int64_t)n_cases;
    ev.ev_off = ev_off.data(); ev.ev_node = ev_node.data(); ev.ev_state = ev_state.data();
    ev.ev_val_off = nullptr; ev.ev_values = nullptr;

    bn::inference::belief_propagation bp(g);
    double checksum = 0.0;
    std::size_t row = 0;
    // the first call with a fresh result pays for pinning the result buffers (reported separately); a hot loop keeps
    // one flat_result and hands it back in (run_flat(ev, opt, result))
    auto const f0 = std::chrono::steady_clock::now();
    bn::inference::belief_propagation::flat_result res = bp.run_flat(ev, opt);
    double const first_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - f0).count();
    for (int i = 0; i < warmup; ++i) bp.run_flat(ev, opt, res);
    std::vector<double> ms;
    auto const t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < calls; ++i) {
        auto const a = std::chrono::steady_clock::now();
        bp.run_flat(ev, opt, res);
        ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count());
        row = res.values_per_case;
        checksum = float_out ? (double)res.marginals_f32[res.marginals_f32.size() - 1] : res.marginals[res.marginals.size() - 1];
    }
    double const total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double mn = ms[0], mx = ms[0];
    for (double x : ms) { mn = x < mn ? x : mn; mx = x > mx ? x : mx; }
    std::printf("{\"calls\": %d, \"n_cases\": %zu, \"sweeps\": %d, \"value\": %.6e, \"unit\": \"case-sweeps/s\", \"ms_per_call\": %.4f, "
                "\"ms_per_call_min_max\": [%.4f, %.4f], \"values_per_case\": %zu, \"bytes_per_value\": %d, \"devices\": %d, "
                "\"d2h_bytes_per_call\": %.0f, \"last_value\": %.17g, \"first_call_ms_incl_upload_compile_and_pinning\": %.3f, "
                "\"what\": \"bn::inference::belief_propagation::run_flat(ev, opt, result) (C++ drop-in header over the C ABI), host buffers in, "
                "results in the caller's reused pinned_buffer; every call re-flattens the graph (CPT edits stay visible, like the reference)\"}\n",
                calls, n_cases, sweeps, (double)n_cases * sweeps * calls / total_s, 1e3 * total_s / calls, mn, mx, row, float_out ? 4 : 8,
                devices < 0 ? bnbp_device_count() : 1, (double)n_cases * (double)row * (float_out ? 4 : 8) + 5.0 * (double)n_cases, checksum, first_ms);
    return 0;
}
