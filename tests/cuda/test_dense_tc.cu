// Stand-alone check of the tensor-core dense contraction kernel (csrc/bnbp_dense_tc.cuh) against a
// double-precision host product: descriptor layouts, hi/lo split, ragged K / N, odd tile counts,
// both state-tile widths.  Built by tests/cuda/Makefile, run on a B200 (tests/test_gpu_dense_tc.py).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#define BNBP_DENSE_TC_KERNEL
#include "bnbp_dense_tc.cuh"

using namespace bnbp;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static double urand()
{
    rng_state += 0x9E3779B97F4A7C15ull;
    unsigned long long z = rng_state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

struct JobSpec { int nf; int card[DENSE_MAXF]; int N; };

static int run(int TBC, int n_tiles, const std::vector<JobSpec>& specs)
{
    const int64_t n_cases = (int64_t)TBC * n_tiles;
    // message slots: the factors of all jobs one after another
    std::vector<DenseJob> jobs;
    std::vector<unsigned long long> dig;
    std::vector<int32_t> ytab;
    std::vector<std::vector<double>> Bs;
    int M = 0, TS = 0, max_rows = 0;
    int64_t tc_values = 0;
    for (const JobSpec& sp : specs) {
        DenseJob jb;
        memset(&jb, 0, sizeof jb);
        jb.nf = sp.nf;
        jb.K = 1;
        for (int f = 0; f < sp.nf; ++f) { jb.f_slot[f] = M; jb.f_card[f] = sp.card[f]; M += sp.card[f]; jb.n_rows += sp.card[f]; jb.K *= sp.card[f]; }
        jb.N = sp.N;
        jb.arena = 2;
        jb.b_off = tc_values;
        tc_values += tc_job_floats(jb.K, jb.N);
        jb.t_off = TS;
        TS += jb.N;
        jb.y0 = (int)ytab.size();
        jb.dig_off = (int64_t)dig.size();
        const int kpad = (jb.K + TC_K - 1) / TC_K * TC_K + TC_K;
        std::vector<int> stride(jb.nf), base(jb.nf);
        int st = 1, rows = 0;
        for (int f = jb.nf - 1; f >= 0; --f) { stride[f] = st; st *= jb.f_card[f]; }
        for (int f = 0; f < jb.nf; ++f) { base[f] = rows; rows += jb.f_card[f]; }
        for (int kk = 0; kk < kpad; ++kk) {
            unsigned long long v = 0;
            for (int f = 0; f < jb.nf; ++f) {
                const int row = kk < jb.K ? base[f] + (kk / stride[f]) % jb.f_card[f] : rows;
                v |= (unsigned long long)row << (8 * f);
            }
            dig.push_back(v);
        }
        for (int y = 0; y < (jb.N + TC_N - 1) / TC_N; ++y) ytab.push_back((int32_t)jobs.size());
        max_rows = std::max(max_rows, (int)jb.n_rows);
        std::vector<double> B((size_t)jb.K * jb.N);
        for (double& b : B) b = 0.05 + 0.95 * urand();
        Bs.push_back(B);
        jobs.push_back(jb);
    }
    std::vector<float> arena((size_t)tc_values);
    for (size_t j = 0; j < jobs.size(); ++j)
        tc_pack_job(arena.data() + jobs[j].b_off, Bs[j].data(), jobs[j].K, jobs[j].N, jobs[j].N, 1);
    // messages: [tiles][M][TBC]
    std::vector<float> msg((size_t)n_tiles * M * TBC);
    for (float& v : msg) v = (float)(0.01 + 0.99 * urand());
    const int stages = tc_stages_for(max_rows, 227 * 1024);
    const size_t smem = tc_smem_bytes(max_rows, stages);
    printf("TBC %d tiles %d jobs %zu grid.y %zu M %d TS %d stages %d smem %zu\n", TBC, n_tiles, jobs.size(), ytab.size(), M, TS, stages, smem);

    DenseJob* d_jobs; int32_t* d_ytab; unsigned long long* d_dig; float *d_arena, *d_msg, *d_t;
    CK(cudaMalloc(&d_jobs, jobs.size() * sizeof(DenseJob)));
    CK(cudaMalloc(&d_ytab, ytab.size() * 4));
    CK(cudaMalloc(&d_dig, dig.size() * 8));
    CK(cudaMalloc(&d_arena, arena.size() * 4));
    CK(cudaMalloc(&d_msg, msg.size() * 4));
    CK(cudaMalloc(&d_t, (size_t)n_tiles * TS * TBC * 4));
    CK(cudaMemcpy(d_jobs, jobs.data(), jobs.size() * sizeof(DenseJob), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ytab, ytab.data(), ytab.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_dig, dig.data(), dig.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_arena, arena.data(), arena.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_msg, msg.data(), msg.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_t, 0xFF, (size_t)n_tiles * TS * TBC * 4));
    CK(cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DenseTcArgs a;
    memset(&a, 0, sizeof a);
    a.jobs = d_jobs; a.ytab = d_ytab; a.dig = d_dig; a.arena_tc = d_arena; a.pl = nullptr; a.msg_cur = d_msg; a.tscr = d_t;
    a.PL = 0; a.M = M; a.TS = TS; a.TBC = TBC; a.n_cases = n_cases; a.stages = stages; a.status = nullptr;
    dense_tc_kernel<<<dim3((unsigned)((n_cases + TC_M - 1) / TC_M), (unsigned)ytab.size()), TC_THREADS, smem>>>(a);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> T((size_t)n_tiles * TS * TBC);
    CK(cudaMemcpy(T.data(), d_t, T.size() * 4, cudaMemcpyDeviceToHost));

    // Accuracy model.  The tensor core ADDS into the fp32 accumulator with truncation (measured on B200:
    // about -0.9 * 2^-24 relative per accumulation), so a product of depth K (3 MMAs per 8 rows) comes out
    // low by up to (3K/8) * 2^-23 -- a factor COMMON to the columns of a case (they share the operand row),
    // which every consumer of these tables removes again (pi, lambda-messages and beliefs are normalised).
    // What is left is the random part of the truncation, a random walk of 3K/8 steps of variance 1/12 ulp^2 --
    // the same order as the rounding noise of a depth-K fp32 FMA chain (the CUDA-core products).
    // Checked: (1) the common factor s_c (median of got / want over the columns) is within that bound,
    // (2) with s_c removed every entry agrees with the double-precision product to
    //     1e-6 + 1e-7 * sqrt(3K)  (K = 1024: 6.5e-6; measured 4.3e-6).
    int bad_total = 0;
    for (size_t j = 0; j < jobs.size(); ++j) {
        const DenseJob& jb = jobs[j];
        double worst = 0.0, worst_bias = 0.0;
        int bad = 0;
        const double bias_bound = (3.0 * ((jb.K + 7) / 8)) * ldexp(1.0, -23) + 1e-6;
        const double resid_bound = 1e-6 + 1e-7 * sqrt(3.0 * jb.K);
        std::vector<double> A((size_t)jb.K), want((size_t)jb.N), ratio((size_t)jb.N);
        for (int64_t c = 0; c < n_cases; ++c) {
            const int64_t tile = c / TBC, ln = c % TBC;
            for (int k = 0; k < jb.K; ++k) {
                double v = 1.0;
                int st = 1;
                for (int f = jb.nf - 1; f >= 0; --f) {
                    const int d = (k / st) % jb.f_card[f];
                    st *= jb.f_card[f];
                    v *= (double)msg[((size_t)tile * M + jb.f_slot[f] + d) * TBC + ln];
                }
                A[k] = v;
            }
            for (int n = 0; n < jb.N; ++n) {
                double w = 0.0;
                for (int k = 0; k < jb.K; ++k) w += A[k] * (double)(float)Bs[j][(size_t)k * jb.N + n];
                want[n] = w;
                ratio[n] = (double)T[((size_t)tile * TS + jb.t_off + n) * TBC + ln] / w;
            }
            std::vector<double> r2 = ratio;
            std::nth_element(r2.begin(), r2.begin() + r2.size() / 2, r2.end());
            const double sc = r2[r2.size() / 2];
            const double bias = fabs(sc - 1.0);
            if (bias > worst_bias || bias != bias) worst_bias = bias;
            if (!(bias <= bias_bound)) {
                if (bad < 6) printf("  job %zu case %lld: common factor %.9g outside 1 +- %.3g\n", j, (long long)c, sc, bias_bound);
                ++bad;
            }
            for (int n = 0; n < jb.N; ++n) {
                const double err = fabs(ratio[n] / sc - 1.0);
                if (!(err <= resid_bound)) {
                    if (bad < 6) printf("  job %zu case %lld col %d: got/want %.9g, common factor %.9g (residual %.3g)\n", j, (long long)c, n, ratio[n], sc, err);
                    ++bad;
                }
                if (err > worst || err != err) worst = err;
            }
        }
        printf("job %zu  K %d N %d nf %d : accumulator truncation %.3g (bound %.3g), residual after the common factor %.3g (bound %.3g), %d bad\n",
               j, jb.K, jb.N, jb.nf, worst_bias, bias_bound, worst, resid_bound, bad);
        bad_total += bad;
    }
    cudaFree(d_jobs); cudaFree(d_ytab); cudaFree(d_dig); cudaFree(d_arena); cudaFree(d_msg); cudaFree(d_t);
    return bad_total ? 1 : 0;
}

int main()
{
    std::vector<JobSpec> specs = {
        {2, {16, 16}, 256},       // one full tile, K = 256
        {2, {5, 7}, 300},         // ragged K (35) and N (two column tiles)
        {1, {8}, 9},              // one MMA step, 9 columns
        {3, {4, 3, 32}, 130},     // three factors, K = 384
        {2, {32, 32}, 1024},      // the card32 shape: K = 1024, four column tiles
    };
    int rc = 0;
    rc |= run(128, 3, specs);     // odd number of 128-case tiles: the last CTA owns one
    rc |= run(256, 2, specs);
    rc |= run(128, 1, specs);
    printf(rc ? "FAILED\n" : "dense_tc ok\n");
    return rc;
}
