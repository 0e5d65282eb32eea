// bnbp_lw.cuh — batched likelihood weighting, the independent statistical check of the
// BP marginals on loopy networks, where BP is only approximate.
//
// Reference: bayesian/inference/likelihood_weighting.hpp (godai0519/BayesianNetwork)
//   operator()        :28-59    n samples; ret[node][state of the sample] += weight; normalise
//   weighted_sample   :122-173  parents before children; an observed node multiplies the weight by
//                               its CPT entry and takes the observed state, any other node draws
//                               its state from its CPT row
//   make_random_by_weight :177-194  first i with  sum_{j<i} w_j <= u < sum_{j<=i} w_j, else the last state
//   normalize         :198-224  sum < 1e-20 -> uniform row
//
// The reference seeds a std::mt19937 from std::random_device, so no two of its runs agree; parity with it
// is statistical.  Here the uniform variate of (case, sample, node) is a pure function of the seed
// (splitmix64 finaliser, see lw_uniform): results do not depend on the order nodes or samples are
// visited in, on the GPU count, or on the block shape -- and oracle/bp_oracle.c (bp_oracle_lw)
// reproduces them draw for draw.
//
// Work decomposition: one block per evidence case, one sample per thread and iteration.  The sampled
// configuration lives in shared memory ([node][thread] bytes), the weighted histogram of the case in
// shared memory too (fp64 atomics), so HBM sees the CPT rows (L2-resident) and one row of marginals
// per case.
#pragma once
#include "bnbp_kernels.cuh"

namespace bnbp {

__host__ __device__ __forceinline__ uint64_t lw_mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t lw_case_key(uint64_t seed, uint64_t case_index) { return lw_mix(seed ^ lw_mix(case_index)); }
__host__ __device__ __forceinline__ uint64_t lw_sample_key(uint64_t case_key, uint64_t sample) { return lw_mix(case_key + sample); }
// uniform in [0, 1) with 53 random bits
__host__ __device__ __forceinline__ double lw_uniform(uint64_t sample_key, uint32_t node)
{
    return (double)(lw_mix(sample_key ^ (uint64_t)node) >> 11) * (1.0 / 9007199254740992.0);
}

struct LwArgs {
    const NodeMeta* nodes;
    const int32_t* order;       // [N] a topological order (parents first)
    const int32_t* par;         // [E] parent node of in-edge e (in_vertexes order: first parent slowest)
    const double* cpt;          // reference-layout CPT arena, fp64
    int32_t n_nodes, V;
    int64_t n_cases;            // cases of this launch
    int64_t case_base;          // global index of the first one (keys the variates)
    const int64_t* ev_off; int64_t ev_base;
    const int32_t* ev_node; const int32_t* ev_state;
    int64_t n_samples;
    uint64_t seed;
    double* out;                // [n_cases][V]
    double* out_wsum;           // [n_cases] or nullptr: total weight of the case's samples
    int32_t* error_flag;
};

// dynamic shared memory: double hist[V] | int32 obs[N] (observed state or -1) | uint8 st[N][blockDim.x]
static __global__ void lw_kernel(const LwArgs a)
{
    extern __shared__ __align__(16) unsigned char lw_smem[];
    double* const hist = reinterpret_cast<double*>(lw_smem);
    int32_t* const obs = reinterpret_cast<int32_t*>(hist + a.V);
    uint8_t* const st = reinterpret_cast<uint8_t*>(obs + a.n_nodes);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int64_t c = blockIdx.x; c < a.n_cases; c += gridDim.x) {
        for (int j = tid; j < a.V; j += nt) hist[j] = 0.0;
        for (int x = tid; x < a.n_nodes; x += nt) obs[x] = -1;
        __syncthreads();
        if (tid == 0) {
            const int64_t e0 = a.ev_off[c] - a.ev_base, e1 = a.ev_off[c + 1] - a.ev_base;
            for (int64_t e = e0; e < e1; ++e) {
                const int node = a.ev_node[e];
                if (node < 0 || node >= a.n_nodes) { *a.error_flag = 1; continue; }
                const int s = a.ev_state[e];
                if (s < 0 || s >= a.nodes[node].card) { *a.error_flag = 3; continue; }
                obs[node] = s;                                    // a node listed twice: the last entry wins
            }
        }
        __syncthreads();
        const uint64_t ckey = lw_case_key(a.seed, (uint64_t)(a.case_base + c));
        for (int64_t s = tid; s < a.n_samples; s += nt) {
            const uint64_t skey = lw_sample_key(ckey, (uint64_t)s);
            double w = 1.0;
            for (int i = 0; i < a.n_nodes; ++i) {
                const int x = a.order[i];
                const NodeMeta nd = a.nodes[x];
                int64_t q = 0;                                    // parent configuration, first parent slowest
                for (int j = 0; j < nd.k; ++j) {
                    const int p = a.par[nd.e0 + j];
                    q = q * a.nodes[p].card + st[(size_t)p * nt + tid];
                }
                const double* const row = a.cpt + nd.cpt_off + q * nd.card;
                int sel;
                if (obs[x] >= 0) {
                    sel = obs[x];
                    w *= row[sel];                                // :152-156
                } else {
                    const double u = lw_uniform(skey, (uint32_t)x);
                    sel = nd.card - 1;                            // :193
                    double total = 0.0;
                    for (int v = 0; v < nd.card; ++v) {
                        const double old_total = total;
                        total += row[v];
                        if (old_total <= u && u < total) { sel = v; break; }
                    }
                }
                st[(size_t)x * nt + tid] = (uint8_t)sel;
            }
            for (int x = 0; x < a.n_nodes; ++x)                   // :44-48
                atomicAdd(&hist[a.nodes[x].bel_off + st[(size_t)x * nt + tid]], w);
        }
        __syncthreads();
        for (int x = tid; x < a.n_nodes; x += nt) {               // :52-55, normalize :198-224
            const NodeMeta nd = a.nodes[x];
            double sum = 0.0;
            for (int v = 0; v < nd.card; ++v) sum += hist[nd.bel_off + v];
            double* const o = a.out + (size_t)c * a.V + nd.bel_off;
            if (sum < 1.0e-20) for (int v = 0; v < nd.card; ++v) o[v] = 1.00 / nd.card;
            else               for (int v = 0; v < nd.card; ++v) o[v] = hist[nd.bel_off + v] / sum;
            if (x == 0 && a.out_wsum) a.out_wsum[c] = sum;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// CPT estimation from a table of samples.
// Reference: bayesian/sampler.hpp  make_cpt :81-163 -- every distinct sample (a full assignment with
// a multiplicity) adds its multiplicity to counter[(node, parent configuration)][state of the node]
// (:104-128); a CPT row is count / row total, or uniform 1/r when the configuration never occurred
// (:131-160).  Counts are integers, so the result is independent of the summation order.
struct CountArgs {
    const NodeMeta* nodes;
    const int32_t* par;         // [E]
    const int32_t* samples;     // [n_rows][N] state of every node
    const int64_t* mult;        // [n_rows] multiplicity or nullptr (1 each)
    int64_t n_rows;
    int32_t n_nodes;
    unsigned long long* counts; // CPT arena layout: counts[cpt_off[X] + q*r + x]
    int32_t* error_flag;
};

// one thread per (row, node): consecutive threads take consecutive nodes of one row (coalesced reads)
static __global__ void cpt_count_kernel(const CountArgs a)
{
    const int64_t total = a.n_rows * (int64_t)a.n_nodes;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / a.n_nodes;
        const int x = (int)(i - row * a.n_nodes);
        const int32_t* const s = a.samples + row * a.n_nodes;
        const NodeMeta nd = a.nodes[x];
        const int sx = s[x];
        bool ok = sx >= 0 && sx < nd.card;
        int64_t q = 0;
        for (int j = 0; j < nd.k; ++j) {
            const int p = a.par[nd.e0 + j];
            const int sp = s[p];
            const int rp = a.nodes[p].card;
            ok = ok && sp >= 0 && sp < rp;
            q = q * rp + sp;
        }
        if (!ok) { *a.error_flag = 3; continue; }
        const long long m = a.mult ? a.mult[row] : 1;
        if (m < 0) { *a.error_flag = 4; continue; }
        atomicAdd(a.counts + nd.cpt_off + q * nd.card + sx, (unsigned long long)m);
    }
}

// one thread per CPT row; row_start[x] = number of CPT rows of the nodes before x
static __global__ void cpt_normalize_kernel(const NodeMeta* nodes, const int64_t* row_start, int32_t n_nodes, int64_t n_cpt_rows,
                                            const unsigned long long* counts, double* cpt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cpt_rows) return;
    int lo = 0, hi = n_nodes;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (row_start[mid] <= i) lo = mid; else hi = mid;
    }
    const NodeMeta nd = nodes[lo];
    const int r = nd.card;
    const int64_t o = nd.cpt_off + (i - row_start[lo]) * r;
    unsigned long long tot = 0;
    for (int v = 0; v < r; ++v) tot += counts[o + v];
    const double parameter = (double)tot;                          // :142
    for (int v = 0; v < r; ++v)
        cpt[o + v] = tot == 0 ? 1.0 / (double)r : (double)counts[o + v] / parameter;   // :147-157
}

} // namespace bnbp
