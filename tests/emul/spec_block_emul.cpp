// spec_block_emul.cpp -- host emulation of a streaming sweep variant that uses the WARP: the last sweep fused with the
// beliefs (variants 6 / 7 of bnbp_spec.cuh: every thread puts its cases' marginals into a per-warp shared-memory tile,
// __syncwarp, then lane j streams column j of all 32 rows to the case-major output).  One OS thread per CUDA thread of a
// block, a pthread barrier per warp for __syncwarp, the block's __shared__ tile a function-local static; blocks run one
// after the other.  Test infrastructure only (tests/test_netcompiler_emul.py).
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
#define __shared__ static

struct emul_dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local emul_dim3 threadIdx;
static emul_dim3 blockIdx, gridDim;

template <typename U> static inline U __ldg(const U* p) { return *p; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static pthread_barrier_t g_warp_barrier[4];
static inline void __syncwarp() { pthread_barrier_wait(&g_warp_barrier[threadIdx.x >> 5]); }
#define asm(...) ((void)0)          // inline PTX exists only in the on-chip code paths (discarded branches here)

#include BNBP_GENERATED

namespace {
struct Launch { T* pl; const T* cur; T* nxt; const unsigned* evbits; bnbp_spec::Aux aux; int tid; };
void* thread_main(void* p)
{
    Launch* l = static_cast<Launch*>(p);
    threadIdx.x = (unsigned)l->tid;
    bnbp_spec_sweep(l->pl, l->cur, l->nxt, l->evbits, l->aux);
    return nullptr;
}
} // namespace

#pragma GCC visibility push(default)
extern "C" {

int emul_v() { return BNBP_V; }
int emul_out_bytes() { return (int)sizeof(OUT); }
void emul_set_cpt(const T* cpt, long long n) { memcpy(bnbp_cpt, cpt, (size_t)n * sizeof(T)); }

// <<<tiles, 128>>> bnbp_spec_sweep(pl, cur, nxt, evbits, aux) with aux.out / aux.n_valid of the fused-belief variants
void emul_launch_last(T* pl, const T* cur, T* nxt, const unsigned* evbits, int tiles, OUT* out, long long n_valid)
{
    bnbp_spec::Aux a;
    memset(&a, 0, sizeof a);
    a.n_inner = 1;
    a.out = out;
    a.n_valid = n_valid;
    gridDim.x = (unsigned)tiles;
    gridDim.y = 1;
    for (int w = 0; w < 4; ++w) pthread_barrier_init(&g_warp_barrier[w], nullptr, 32);
    for (int t = 0; t < tiles; ++t) {
        blockIdx.x = (unsigned)t;
        pthread_t th[128];
        Launch ln[128];
        for (int i = 0; i < 128; ++i) {
            ln[i] = Launch{pl, cur, nxt, evbits, a, i};
            pthread_create(&th[i], nullptr, thread_main, &ln[i]);
        }
        for (int i = 0; i < 128; ++i) pthread_join(th[i], nullptr);
    }
    for (int w = 0; w < 4; ++w) pthread_barrier_destroy(&g_warp_barrier[w]);
}

}
#pragma GCC visibility pop
