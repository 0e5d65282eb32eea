/* bp_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp64) of the reference's loopy belief propagation
 *   bayesian/inference/belief_propagation.hpp  (godai0519/BayesianNetwork)
 * over a flat CSR network.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product path (libbnbp) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (1) the reference's own test vectors (libs/bayesian/test/belief_propagation.cpp:64-301,
 *       17-digit table in SURVEY.md section 4, committed as tests/golden/reference_tests.json), and
 *   (2) the unmodified reference header compiled in place (oracle/_ref/libbnref.so, built by
 *       oracle/Makefile from /root/reference) on random polytrees, loopy DAGs, soft evidence and
 *       the NaN case, to 1e-12, plus fixtures generated from it under tests/golden/.
 *
 * Each function cites the reference lines it follows.  Loop orders are the reference's
 * (first parent slowest; x outer / configuration inner in the lambda message) so that the
 * only rounding differences left are the reference's own run-dependent unordered_map orders.
 *
 * Extensions that the reference does not have (all default to reference behaviour):
 *   max_sweeps      cap on the while(true) of :75 (the reference never stops on oscillation)
 *   damping         msg <- (1-d)*msg_new + d*msg_old on both message kinds
 *   check_interval  evaluate the :105-131 delta only every n-th sweep
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t n;
    const int32_t* card;
    const int32_t* poff;     /* parents CSR (graph.hpp:389-413: ascending vertex index) */
    const int32_t* par;
    const int64_t* coff;
    const double*  cpt;
    int32_t* choff;          /* children CSR (graph.hpp:362-386: ascending vertex index) */
    int32_t* chd;            /* child node id */
    int32_t* ched;           /* in-edge index (position in par[]) of that child edge */
    int64_t* voff;           /* per node offset into pi/lambda vectors, [n+1] */
    int64_t* moff;           /* per in-edge offset into message vectors (length card[parent]), [E+1] */
    int32_t semiring;        /* 0 = the reference's sum-product; 1 = max-product (extension: every += over parent
                                configurations / child states below becomes a maximum; fmax drops NaN) */
} net_t;

static int build_net(net_t* g)
{
    int32_t n = g->n, E = g->poff[n];
    g->choff = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
    g->chd = (int32_t*)malloc(sizeof(int32_t) * (size_t)(E > 0 ? E : 1));
    g->ched = (int32_t*)malloc(sizeof(int32_t) * (size_t)(E > 0 ? E : 1));
    g->voff = (int64_t*)malloc(sizeof(int64_t) * ((size_t)n + 1));
    g->moff = (int64_t*)malloc(sizeof(int64_t) * ((size_t)E + 1));
    if (!g->choff || !g->chd || !g->ched || !g->voff || !g->moff) return -1;
    g->voff[0] = 0;
    for (int32_t i = 0; i < n; ++i) g->voff[i + 1] = g->voff[i] + g->card[i];
    g->moff[0] = 0;
    for (int32_t e = 0; e < E; ++e) g->moff[e + 1] = g->moff[e] + g->card[g->par[e]];
    for (int32_t e = 0; e < E; ++e) g->choff[g->par[e] + 1]++;
    for (int32_t i = 0; i < n; ++i) g->choff[i + 1] += g->choff[i];
    int32_t* fill = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
    if (!fill) return -1;
    /* children in ascending child index: iterate child nodes in order */
    for (int32_t x = 0; x < n; ++x)
        for (int32_t e = g->poff[x]; e < g->poff[x + 1]; ++e) {
            int32_t u = g->par[e];
            int32_t slot = g->choff[u] + fill[u]++;
            g->chd[slot] = x;
            g->ched[slot] = e;
        }
    free(fill);
    return 0;
}

static void free_net(net_t* g)
{
    free(g->choff); free(g->chd); free(g->ched); free(g->voff); free(g->moff);
}

/* belief_propagation.hpp:298-311 — divide by the plain sum, no zero guard (0/0 -> NaN). */
static void normalize(double* v, int32_t r)
{
    double sum = 0;
    for (int32_t i = 0; i < r; ++i) sum += v[i];
    for (int32_t i = 0; i < r; ++i) v[i] /= sum;
}

typedef struct {
    double *pi, *lam, *pmsg, *lmsg;          /* current  (:323-327) */
    double *npi, *nlam, *npmsg, *nlmsg;      /* future   (:330-333) */
    uint8_t* is_ev;                          /* preconditional_node_ (:321) */
    int32_t* cfg;                            /* current parent configuration */
} state_t;

/* :174-200 calculate_pi */
static void calc_pi(const net_t* g, const state_t* s, int32_t x, double* out)
{
    int32_t r = g->card[x], e0 = g->poff[x], k = g->poff[x + 1] - e0;
    const double* cpt = g->cpt + g->coff[x];
    for (int32_t i = 0; i < r; ++i) out[i] = 0.0;
    int64_t Q = 1;
    for (int32_t j = 0; j < k; ++j) { Q *= g->card[g->par[e0 + j]]; s->cfg[j] = 0; }
    for (int64_t q = 0; q < Q; ++q) {
        for (int32_t i = 0; i < r; ++i) {
            double value = cpt[q * r + i];
            for (int32_t j = 0; j < k; ++j) value *= s->pmsg[g->moff[e0 + j] + s->cfg[j]];
            if (g->semiring) out[i] = fmax(out[i], value); else out[i] += value;
        }
        for (int32_t j = k - 1; j >= 0; --j) {       /* odometer, last parent fastest (:286-290) */
            if (++s->cfg[j] < g->card[g->par[e0 + j]]) break;
            s->cfg[j] = 0;
        }
    }
    normalize(out, r);
}

/* :202-218 calculate_pi_i(from = child x via in-edge e, target = parent u) */
static void calc_pi_msg(const net_t* g, const state_t* s, int32_t e, double* out)
{
    int32_t u = g->par[e], r = g->card[u];
    for (int32_t i = 0; i < r; ++i) out[i] = s->pi[g->voff[u] + i];
    for (int32_t i = 0; i < r; ++i)
        for (int32_t c = g->choff[u]; c < g->choff[u + 1]; ++c) {
            if (g->ched[c] == e) continue;           /* every child of u except `from` */
            out[i] *= s->lmsg[g->moff[g->ched[c]] + i];
        }
    normalize(out, r);
}

/* :220-238 calculate_lambda */
static void calc_lambda(const net_t* g, const state_t* s, int32_t x, double* out)
{
    int32_t r = g->card[x];
    for (int32_t i = 0; i < r; ++i) out[i] = 1.0;
    for (int32_t i = 0; i < r; ++i)
        for (int32_t c = g->choff[x]; c < g->choff[x + 1]; ++c)
            out[i] *= s->lmsg[g->moff[g->ched[c]] + i];
    normalize(out, r);
}

/* :240-266 calculate_lambda_k(from = child x, target = its jt-th parent) */
static void calc_lambda_msg(const net_t* g, const state_t* s, int32_t x, int32_t jt, double* out)
{
    int32_t r = g->card[x], e0 = g->poff[x], k = g->poff[x + 1] - e0;
    int32_t rt = g->card[g->par[e0 + jt]];
    const double* cpt = g->cpt + g->coff[x];
    int64_t Q = 1;
    for (int32_t j = 0; j < k; ++j) Q *= g->card[g->par[e0 + j]];
    for (int32_t a = 0; a < rt; ++a) out[a] = 0.0;
    for (int32_t i = 0; i < r; ++i) {
        double times = s->lam[g->voff[x] + i];
        for (int32_t j = 0; j < k; ++j) s->cfg[j] = 0;
        for (int64_t q = 0; q < Q; ++q) {
            double value = times * cpt[q * r + i];
            for (int32_t j = 0; j < k; ++j)
                if (j != jt) value *= s->pmsg[g->moff[e0 + j] + s->cfg[j]];
            if (g->semiring) out[s->cfg[jt]] = fmax(out[s->cfg[jt]], value); else out[s->cfg[jt]] += value;
            for (int32_t j = k - 1; j >= 0; --j) {
                if (++s->cfg[j] < g->card[g->par[e0 + j]]) break;
                s->cfg[j] = 0;
            }
        }
    }
    normalize(out, rt);
}

/* One evidence case: belief_propagation.hpp:31-159. */
static void run_case(const net_t* g, state_t* s,
                     int64_t nev, const int32_t* ev_node, const int32_t* ev_state,
                     const int64_t* ev_val_off, const double* ev_values,
                     double eps, int32_t max_sweeps, double damping, int32_t check_interval,
                     double* out, int32_t* out_sweeps, uint8_t* out_conv)
{
    int32_t n = g->n, E = g->poff[n];
    int64_t V = g->voff[n], M = g->moff[E];
    /* :35-65 init */
    for (int64_t i = 0; i < V; ++i) { s->pi[i] = 1.0; s->lam[i] = 1.0; }
    for (int64_t i = 0; i < M; ++i) { s->pmsg[i] = 1.0; s->lmsg[i] = 1.0; }
    for (int32_t x = 0; x < n; ++x)
        if (g->poff[x + 1] == g->poff[x])                     /* :58-64 root: raw prior row */
            for (int32_t i = 0; i < g->card[x]; ++i) s->pi[g->voff[x] + i] = g->cpt[g->coff[x] + i];
    /* :68-73 evidence into both pi and lambda */
    memset(s->is_ev, 0, (size_t)n);
    for (int64_t t = 0; t < nev; ++t) {
        int32_t v = ev_node[t], r = g->card[v];
        s->is_ev[v] = 1;
        for (int32_t i = 0; i < r; ++i) {
            double val = ev_values ? ev_values[ev_val_off[t] + i] : (ev_state[t] == i ? 1.0 : 0.0);
            s->pi[g->voff[v] + i] = val;
            s->lam[g->voff[v] + i] = val;
        }
    }
    int32_t sweeps = 0;
    uint8_t conv = 0;
    for (;;) {                                                /* :75 while(true) */
        /* :78-88 messages, all from time-t state */
        for (int32_t x = 0; x < n; ++x)
            for (int32_t e = g->poff[x]; e < g->poff[x + 1]; ++e) {
                calc_pi_msg(g, s, e, s->npmsg + g->moff[e]);
                calc_lambda_msg(g, s, x, e - g->poff[x], s->nlmsg + g->moff[e]);
            }
        /* :91-101 node updates, evidence nodes keep their vectors (:177,:223) */
        for (int32_t x = 0; x < n; ++x) {
            if (s->is_ev[x]) {
                memcpy(s->npi + g->voff[x], s->pi + g->voff[x], sizeof(double) * (size_t)g->card[x]);
                memcpy(s->nlam + g->voff[x], s->lam + g->voff[x], sizeof(double) * (size_t)g->card[x]);
            } else {
                calc_pi(g, s, x, s->npi + g->voff[x]);
                calc_lambda(g, s, x, s->nlam + g->voff[x]);
            }
        }
        if (damping != 0.0)                                   /* extension */
            for (int64_t i = 0; i < M; ++i) {
                s->npmsg[i] = (1.0 - damping) * s->npmsg[i] + damping * s->pmsg[i];
                s->nlmsg[i] = (1.0 - damping) * s->nlmsg[i] + damping * s->lmsg[i];
            }
        ++sweeps;
        int test = eps > 0.0 && (sweeps % check_interval == 0 || sweeps >= max_sweeps);
        double maxdiff = DBL_MIN;                             /* :105 */
        if (test)
            for (int64_t i = 0; i < M; ++i) {                 /* :106-131; std::max keeps lhs on NaN */
                double d1 = fabs(s->npmsg[i] - s->pmsg[i]);
                double d2 = fabs(s->nlmsg[i] - s->lmsg[i]);
                if (maxdiff < d1) maxdiff = d1;
                if (maxdiff < d2) maxdiff = d2;
            }
        /* :135-143 commit */
        double* t;
        t = s->pi; s->pi = s->npi; s->npi = t;
        t = s->lam; s->lam = s->nlam; s->nlam = t;
        t = s->pmsg; s->pmsg = s->npmsg; s->npmsg = t;
        t = s->lmsg; s->lmsg = s->nlmsg; s->nlmsg = t;
        if (test && maxdiff < eps) { conv = 1; break; }       /* :147 */
        if (sweeps >= max_sweeps) break;                      /* extension */
    }
    /* :151-158 belief = normalize(pi % lambda) */
    for (int32_t x = 0; x < n; ++x) {
        int32_t r = g->card[x];
        double* o = out + g->voff[x];
        for (int32_t i = 0; i < r; ++i) o[i] = s->pi[g->voff[x] + i] * s->lam[g->voff[x] + i];
        normalize(o, r);
    }
    if (out_sweeps) *out_sweeps = sweeps;
    if (out_conv) *out_conv = conv;
}

static int alloc_state(const net_t* g, state_t* s)
{
    int32_t n = g->n, E = g->poff[n];
    size_t V = (size_t)g->voff[n] + 1, M = (size_t)g->moff[E] + 1;
    int32_t kmax = 1;
    for (int32_t x = 0; x < n; ++x)
        if (g->poff[x + 1] - g->poff[x] > kmax) kmax = g->poff[x + 1] - g->poff[x];
    s->pi = (double*)malloc(sizeof(double) * V);   s->npi = (double*)malloc(sizeof(double) * V);
    s->lam = (double*)malloc(sizeof(double) * V);  s->nlam = (double*)malloc(sizeof(double) * V);
    s->pmsg = (double*)malloc(sizeof(double) * M); s->npmsg = (double*)malloc(sizeof(double) * M);
    s->lmsg = (double*)malloc(sizeof(double) * M); s->nlmsg = (double*)malloc(sizeof(double) * M);
    s->is_ev = (uint8_t*)malloc((size_t)n + 1);
    s->cfg = (int32_t*)malloc(sizeof(int32_t) * (size_t)kmax);
    return (s->pi && s->npi && s->lam && s->nlam && s->pmsg && s->npmsg && s->lmsg && s->nlmsg &&
            s->is_ev && s->cfg) ? 0 : -1;
}

static void free_state(state_t* s)
{
    free(s->pi); free(s->npi); free(s->lam); free(s->nlam);
    free(s->pmsg); free(s->npmsg); free(s->lmsg); free(s->nlmsg);
    free(s->is_ev); free(s->cfg);
}

int bp_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Returns 0 on success.  out_marginals is [n_cases][sum card] row-major.  semiring: see net_t. */
int bp_oracle_run_semiring(int32_t n_nodes, const int32_t* card, const int32_t* parent_off,
                  const int32_t* parents, const int64_t* cpt_off, const double* cpt,
                  int64_t n_cases, const int64_t* ev_off, const int32_t* ev_node,
                  const int32_t* ev_state, const int64_t* ev_val_off, const double* ev_values,
                  double eps, int32_t max_sweeps, double damping, int32_t check_interval,
                  int32_t n_threads, int32_t semiring,
                  double* out_marginals, int32_t* out_sweeps, uint8_t* out_converged)
{
    net_t g;
    memset(&g, 0, sizeof g);
    g.n = n_nodes; g.card = card; g.poff = parent_off; g.par = parents; g.coff = cpt_off; g.cpt = cpt;
    g.semiring = semiring;
    if (build_net(&g)) return -1;
    if (max_sweeps <= 0) max_sweeps = 1 << 30;
    if (check_interval <= 0) check_interval = 1;
    int64_t V = g.voff[n_nodes];
    int err = 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
#pragma omp parallel num_threads(n_threads)
    {
        state_t s;
        memset(&s, 0, sizeof s);
        if (alloc_state(&g, &s)) {
#pragma omp atomic write
            err = -1;
        } else {
#pragma omp for schedule(dynamic, 4)
            for (int64_t c = 0; c < n_cases; ++c) {
                int64_t a = ev_off ? ev_off[c] : 0, b = ev_off ? ev_off[c + 1] : 0;
                run_case(&g, &s, b - a, ev_node + a, ev_state ? ev_state + a : NULL,
                         ev_val_off ? ev_val_off + a : NULL, ev_values,
                         eps, max_sweeps, damping, check_interval,
                         out_marginals + c * V,
                         out_sweeps ? out_sweeps + c : NULL,
                         out_converged ? out_converged + c : NULL);
            }
        }
        free_state(&s);
    }
    free_net(&g);
    return err;
}

int bp_oracle_run(int32_t n_nodes, const int32_t* card, const int32_t* parent_off,
                  const int32_t* parents, const int64_t* cpt_off, const double* cpt,
                  int64_t n_cases, const int64_t* ev_off, const int32_t* ev_node,
                  const int32_t* ev_state, const int64_t* ev_val_off, const double* ev_values,
                  double eps, int32_t max_sweeps, double damping, int32_t check_interval,
                  int32_t n_threads,
                  double* out_marginals, int32_t* out_sweeps, uint8_t* out_converged)
{
    return bp_oracle_run_semiring(n_nodes, card, parent_off, parents, cpt_off, cpt, n_cases, ev_off, ev_node, ev_state,
                                  ev_val_off, ev_values, eps, max_sweeps, damping, check_interval, n_threads, 0,
                                  out_marginals, out_sweeps, out_converged);
}

/* ------------------------------------------------------------------------------------------------
 * Likelihood weighting -- restatement of bayesian/inference/likelihood_weighting.hpp
 *   operator()             :28-59
 *   weighted_sample        :122-173  (parents before children; observed node: w *= CPT entry)
 *   make_random_by_weight  :177-194
 *   normalize              :198-224  (sum < 1e-20 -> uniform)
 * The reference draws from a std::mt19937 seeded by std::random_device (:229-237), so its output is not
 * reproducible: parity with it is statistical (tests/test_lw.py compares against oracle/_ref/libbnref_lw.so
 * within sampling error, and against exact enumeration).  The variate of (case, sample, node) used here is
 * the counter-based one of csrc/bnbp_lw.cuh, restated below, so the CUDA kernel can be checked draw for draw. */
static uint64_t lw_mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int bp_oracle_lw(int32_t n_nodes, const int32_t* card, const int32_t* parent_off,
                 const int32_t* parents, const int64_t* cpt_off, const double* cpt,
                 int64_t n_cases, const int64_t* ev_off, const int32_t* ev_node, const int32_t* ev_state,
                 int64_t n_samples, uint64_t seed, int64_t case_base, double* out, double* out_wsum)
{
    int32_t n = n_nodes;
    int64_t V = 0;
    int64_t* voff = (int64_t*)malloc(sizeof(int64_t) * ((size_t)n + 1));
    int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* indeg = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* obs = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* st = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    if (!voff || !order || !indeg || !obs || !st) return -1;
    voff[0] = 0;
    for (int32_t i = 0; i < n; ++i) voff[i + 1] = voff[i] + card[i];
    V = voff[n];
    double* hist = (double*)malloc(sizeof(double) * (size_t)V);
    if (!hist) return -1;
    /* any topological order gives the same sample: the variates are keyed by node, not by draw order */
    int32_t done = 0;
    for (int32_t i = 0; i < n; ++i) indeg[i] = parent_off[i + 1] - parent_off[i];
    while (done < n) {
        int32_t before = done;
        for (int32_t i = 0; i < n; ++i) {
            if (indeg[i] != 0) continue;
            order[done++] = i;
            indeg[i] = -1;
            for (int32_t j = 0; j < n; ++j)
                for (int32_t e = parent_off[j]; e < parent_off[j + 1]; ++e)
                    if (parents[e] == i) indeg[j]--;
        }
        if (done == before) return -2;          /* cycle */
    }
    for (int64_t c = 0; c < n_cases; ++c) {
        for (int64_t j = 0; j < V; ++j) hist[j] = 0.0;
        for (int32_t i = 0; i < n; ++i) obs[i] = -1;
        for (int64_t e = ev_off[c]; e < ev_off[c + 1]; ++e) obs[ev_node[e]] = ev_state[e];
        uint64_t ckey = lw_mix(seed ^ lw_mix((uint64_t)(case_base + c)));
        for (int64_t s = 0; s < n_samples; ++s) {
            uint64_t skey = lw_mix(ckey + (uint64_t)s);
            double w = 1.0;
            for (int32_t i = 0; i < n; ++i) {
                int32_t x = order[i];
                int64_t q = 0;
                for (int32_t e = parent_off[x]; e < parent_off[x + 1]; ++e) q = q * card[parents[e]] + st[parents[e]];
                const double* row = cpt + cpt_off[x] + q * card[x];
                if (obs[x] >= 0) {                              /* :152-156 */
                    w *= row[obs[x]];
                    st[x] = obs[x];
                } else {                                        /* :157-161, :177-194 */
                    double u = (double)(lw_mix(skey ^ (uint64_t)(uint32_t)x) >> 11) * (1.0 / 9007199254740992.0);
                    int32_t sel = card[x] - 1;
                    double total = 0.0;
                    for (int32_t v = 0; v < card[x]; ++v) {
                        double old_total = total;
                        total += row[v];
                        if (old_total <= u && u < total) { sel = v; break; }
                    }
                    st[x] = sel;
                }
            }
            for (int32_t x = 0; x < n; ++x) hist[voff[x] + st[x]] += w;      /* :44-48 */
        }
        for (int32_t x = 0; x < n; ++x) {                       /* :52-55 */
            double sum = 0.0;
            for (int32_t v = 0; v < card[x]; ++v) sum += hist[voff[x] + v];
            for (int32_t v = 0; v < card[x]; ++v)
                out[c * V + voff[x] + v] = sum < 1.0e-20 ? 1.00 / card[x] : hist[voff[x] + v] / sum;
            if (x == 0 && out_wsum) out_wsum[c] = sum;
        }
    }
    free(voff); free(order); free(indeg); free(obs); free(st); free(hist);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * CPT estimation -- restatement of bayesian/sampler.hpp make_cpt :81-163 over flat arrays:
 * counter[(node, parent configuration)][state] += multiplicity (:104-128); row = count / total,
 * uniform 1/r for a configuration that never occurred (:131-160).
 * Parity status: the reference's sampler.hpp needs Boost (absent here), so this restatement is pinned by
 * hand-computed known answers and by the reference's own bayesian_network/sampler test data where they
 * apply (tests/test_cpt_estimation.py), not by a compiled reference. */
int bp_oracle_make_cpt(int32_t n_nodes, const int32_t* card, const int32_t* parent_off, const int32_t* parents,
                       const int64_t* cpt_off, const int32_t* samples, const int64_t* mult, int64_t n_rows, double* out_cpt)
{
    int64_t total = cpt_off[n_nodes];
    uint64_t* cnt = (uint64_t*)calloc((size_t)(total > 0 ? total : 1), sizeof(uint64_t));
    if (!cnt) return -1;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int32_t* s = samples + r * n_nodes;
        for (int32_t x = 0; x < n_nodes; ++x) {
            int64_t q = 0;
            for (int32_t e = parent_off[x]; e < parent_off[x + 1]; ++e) q = q * card[parents[e]] + s[parents[e]];
            cnt[cpt_off[x] + q * card[x] + s[x]] += (uint64_t)(mult ? mult[r] : 1);
        }
    }
    for (int32_t x = 0; x < n_nodes; ++x) {
        int64_t rows = (cpt_off[x + 1] - cpt_off[x]) / card[x];
        for (int64_t q = 0; q < rows; ++q) {
            uint64_t tot = 0;
            for (int32_t v = 0; v < card[x]; ++v) tot += cnt[cpt_off[x] + q * card[x] + v];
            double parameter = (double)tot;
            for (int32_t v = 0; v < card[x]; ++v)
                out_cpt[cpt_off[x] + q * card[x] + v] = tot == 0 ? 1.0 / (double)card[x]
                                                                 : (double)cnt[cpt_off[x] + q * card[x] + v] / parameter;
        }
    }
    free(cnt);
    return 0;
}
