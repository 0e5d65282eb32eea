// onchip_emul.cpp -- host emulation of the ON-CHIP multi-sweep kernel (bnbp_onchip.cuh behind bnbp_spec.cuh), one OS thread
// per CUDA thread of ONE CTA: pthread barriers stand for `bar.sync` and for the convergence points of the warp collectives
// (__ballot_sync / __shfl_sync are called by all 32 lanes of a warp in this kernel), the shared memory of the CTA is a static
// buffer, the ticket counter an atomic.  Test infrastructure only (tests/test_onchip_emul.py): the case hand-out, the two-phase
// sweep over one message buffer, the stopping rule with retire / refill batches and the belief write of the headline kernel
// run against the oracle WITHOUT a GPU -- and under ThreadSanitizer the barrier placement is checked for data races.
//
// The test replaces the three inline-PTX statements of the source textually before it is compiled here:
//   bar.sync 1, OC_THREADS  -> emul_cta_barrier();      mov.u32 smid -> 0;      rcp.approx.ftz.f64 -> 1.0 / s (then the Newton steps)
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <pthread.h>
#include <vector>

#define __device__
#define __global__
#define __constant__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__
#define __align__(n) __attribute__((aligned(n)))

struct emul_dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local emul_dim3 threadIdx;
static emul_dim3 blockIdx, gridDim;

template <typename U> static inline U __ldg(const U* p) { return *p; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned __activemask() { return 0xffffffffu; }

constexpr int EMUL_MAX_WARPS = 16;
static pthread_barrier_t g_cta_barrier;
static pthread_barrier_t g_warp_barrier[EMUL_MAX_WARPS];
static unsigned g_ballot[EMUL_MAX_WARPS][32];
static unsigned long long g_shfl[EMUL_MAX_WARPS][32];

static inline void emul_cta_barrier() { pthread_barrier_wait(&g_cta_barrier); }
static inline void __syncwarp() { pthread_barrier_wait(&g_warp_barrier[threadIdx.x >> 5]); }

static inline unsigned __ballot_sync(unsigned, bool pred)
{
    const int w = (int)(threadIdx.x >> 5), l = (int)(threadIdx.x & 31);
    g_ballot[w][l] = pred ? 1u : 0u;
    pthread_barrier_wait(&g_warp_barrier[w]);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= g_ballot[w][i] << i;
    pthread_barrier_wait(&g_warp_barrier[w]);             // nobody overwrites the slots before everybody has read them
    return m;
}

static inline unsigned long long __shfl_sync(unsigned, unsigned long long v, int src)
{
    const int w = (int)(threadIdx.x >> 5), l = (int)(threadIdx.x & 31);
    g_shfl[w][l] = v;
    pthread_barrier_wait(&g_warp_barrier[w]);
    const unsigned long long r = g_shfl[w][src & 31];
    pthread_barrier_wait(&g_warp_barrier[w]);
    return r;
}

static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v)
{
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}

alignas(16) unsigned char oc_smem[232 * 1024];            // the CTA's dynamic shared memory

#include BNBP_GENERATED

namespace {
struct Launch { bnbp_spec::OcArgs args; int tid; };
void* lane_main(void* p)
{
    Launch* l = static_cast<Launch*>(p);
    threadIdx.x = (unsigned)l->tid;
    bnbp_onchip_run(l->args);
    return nullptr;
}
} // namespace

#pragma GCC visibility push(default)
extern "C" {

int emul_threads() { return bnbp_spec::OC_THREADS; }
int emul_value_bytes() { return (int)sizeof(T); }
int emul_out_bytes() { return (int)sizeof(OUT); }
int emul_smem_bytes() { return (int)((BNBP_PL + BNBP_M) * 32 * sizeof(T) + bnbp_spec::ROLES * 32 * sizeof(T) + 32 * 8); }
void emul_set_cpt(const T* cpt, long long n) { memcpy(bnbp_cpt, cpt, (size_t)n * sizeof(T)); }

// One CTA of <<<1, OC_THREADS, smem>>> bnbp_onchip_run(args): a persistent group that takes cases from the ticket counter
// until none is left -- the whole batch flows through its 32 lanes, retire / refill included.
int emul_run(const long long* ev_off, const int* ev_node, const int* ev_state, long long n_cases, OUT* out,
             const int* bel_col, long long out_stride, int* out_sweeps, unsigned char* out_conv, T eps, T damping,
             int max_sweeps, int interval)
{
    if (emul_smem_bytes() > (int)sizeof oc_smem) return -1;
    static unsigned long long ticket;
    static int error_flag;
    ticket = 0;
    error_flag = 0;
    bnbp_spec::OcArgs a;
    memset(&a, 0, sizeof a);
    a.ev_off = ev_off; a.ev_base = 0; a.ev_node = ev_node; a.ev_state = ev_state; a.n_cases = n_cases;
    a.out = out; a.bel_col = bel_col; a.out_stride = out_stride; a.out_sweeps = out_sweeps; a.out_conv = out_conv;
    a.ticket = &ticket; a.error_flag = &error_flag;
    a.eps = eps; a.damping = damping; a.max_sweeps = max_sweeps; a.interval = interval;
    a.first_reserved_sm = 1 << 30;
    const int nt = bnbp_spec::OC_THREADS;
    pthread_barrier_init(&g_cta_barrier, nullptr, (unsigned)nt);
    for (int w = 0; w < nt / 32; ++w) pthread_barrier_init(&g_warp_barrier[w], nullptr, 32);
    std::vector<pthread_t> th((size_t)nt);
    std::vector<Launch> ln((size_t)nt);
    for (int t = 0; t < nt; ++t) {
        ln[(size_t)t].args = a;
        ln[(size_t)t].tid = t;
        pthread_create(&th[(size_t)t], nullptr, lane_main, &ln[(size_t)t]);
    }
    for (int t = 0; t < nt; ++t) pthread_join(th[(size_t)t], nullptr);
    pthread_barrier_destroy(&g_cta_barrier);
    for (int w = 0; w < nt / 32; ++w) pthread_barrier_destroy(&g_warp_barrier[w]);
    return error_flag;
}

}
#pragma GCC visibility pop

#ifdef EMUL_MAIN
// Stand-alone form for ThreadSanitizer (an instrumented shared object cannot be loaded into an uninstrumented Python):
//   onchip_emul <in.bin> <out.bin>
// in.bin : int64 n_cases, nnz, n_nodes, n_cpt, stride, max_sweeps, interval; double eps, damping;
//          int64 ev_off[n_cases+1]; int32 ev_node[nnz], ev_state[nnz], bel_col[n_nodes]; double cpt[n_cpt]
// out.bin: OUT out[n_cases*stride]; int32 sweeps[n_cases]; uint8 conv[n_cases]
#include <cstdio>
int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    long long h[7];
    double e[2];
    if (fread(h, 8, 7, f) != 7 || fread(e, 8, 2, f) != 2) return 2;
    const long long n = h[0], nnz = h[1], nn = h[2], nc = h[3], stride = h[4];
    std::vector<long long> ev_off((size_t)n + 1);
    std::vector<int> ev_node((size_t)nnz + 1), ev_state((size_t)nnz + 1), bel_col((size_t)nn);
    std::vector<double> cpt((size_t)nc);
    bool ok = fread(ev_off.data(), 8, (size_t)n + 1, f) == (size_t)n + 1;
    ok = ok && fread(ev_node.data(), 4, (size_t)nnz, f) == (size_t)nnz && fread(ev_state.data(), 4, (size_t)nnz, f) == (size_t)nnz;
    ok = ok && fread(bel_col.data(), 4, (size_t)nn, f) == (size_t)nn && fread(cpt.data(), 8, (size_t)nc, f) == (size_t)nc;
    fclose(f);
    if (!ok) return 2;
    std::vector<T> cpt_t(cpt.begin(), cpt.end());
    emul_set_cpt(cpt_t.data(), nc);
    std::vector<OUT> out((size_t)(n * stride), (OUT)-7);
    std::vector<int> sweeps((size_t)n, 0);
    std::vector<unsigned char> conv((size_t)n, 0);
    const int rc = emul_run(ev_off.data(), ev_node.data(), ev_state.data(), n, out.data(), bel_col.data(), stride, sweeps.data(),
                            conv.data(), (T)e[0], (T)e[1], (int)h[5], (int)h[6]);
    if (rc) return 3;
    f = fopen(argv[2], "wb");
    if (!f) return 2;
    fwrite(out.data(), sizeof(OUT), out.size(), f);
    fwrite(sweeps.data(), 4, sweeps.size(), f);
    fwrite(conv.data(), 1, conv.size(), f);
    fclose(f);
    return 0;
}
#endif
