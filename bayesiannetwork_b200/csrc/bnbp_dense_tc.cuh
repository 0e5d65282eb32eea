// bnbp_dense_tc.cuh — tensor-core (tcgen05 / TMEM) variant of the dense contraction path, fp32 handles.
//
// Same two matrix products per dense node and sweep as bnbp_dense.cuh (calculate_pi / calculate_lambda_k,
// belief_propagation.hpp:174-200, :240-266, regrouped as CPT x batch), but the products run on the 5th
// generation tensor cores:  T[c][n] = sum_k A[c][k] * B[k][n]  with
//   M = 256 cases per CTA (two 128-row accumulators), N = 256 result columns, K step 8 (kind::tf32).
//
// Precision.  The fp32 mode of the library promises 1e-5 relative on the posterior marginals; one
// tf32 product (10-bit mantissa) does not hold that, so each operand is split  v = hi + lo  with
// hi = v with the low 13 mantissa bits cleared (exact in tf32) and lo = v - hi (exact in fp32, at most
// 13 significant bits), and three MMAs per step  lo*hi + hi*lo + hi*hi  accumulate in fp32 in TMEM
// ("3xTF32": error ~2^-21 relative per product instead of 2^-11).  The CPT operand (B) is split ONCE
// at bnbp_create / bnbp_refresh_cpt and stored pre-tiled in exactly the shared-memory image the MMA
// reads (K-major, no swizzle, 8x16-byte core matrices), so a stage of B is ONE 16 KB bulk-async copy
// (cp.async.bulk + mbarrier complete_tx, no tensor map).  The batch operand (A) never exists in
// memory: its rows are products of message entries, formed on the fly by 256 producer threads from the
// staged factor rows (as in bnbp_dense.cuh) and written hi/lo straight into the MMA layout.
//
// Roles (320 threads, 1 CTA per SM, all 512 TMEM columns):
//   warp 0  lane 0 : B loader      wait empty[s] -> expect_tx + bulk copy            -> full_b[s]
//   warp 1  lane 0 : MMA issuer    wait full_a[s], full_b[s] -> 6 x tcgen05.mma -> tcgen05.commit -> empty[s]
//   warps 2..9     : A producers   wait empty[s] -> 8 operand values / thread, hi/lo -> full_a[s];
//                    then epilogue: tcgen05.ld of the accumulators -> per-case tables in HBM
//                    (a warp stores 32 consecutive cases of one column: 128-byte segments).
// Every mbarrier wait is bounded (2 s of %globaltimer, then trap) so a protocol bug is an error, not a hang.
#pragma once
#include "bnbp_dense.h"

namespace bnbp {

constexpr int TC_M = 256;                    // cases per CTA
constexpr int TC_N = 256;                    // result columns per CTA
constexpr int TC_K = 8;                      // contraction depth per stage (one tf32 MMA)
constexpr int TC_THREADS = 320;
constexpr int TC_A_PART = 128 * TC_K * 4;    // bytes of one operand part (hi or lo) of one 128-case tile
constexpr int TC_B_PART = TC_N * TC_K * 4;   // bytes of the hi (or lo) part of a B stage
constexpr int TC_STAGE = 4 * TC_A_PART + 2 * TC_B_PART;   // 32 KB
constexpr int TC_HDR = 1024;                 // barriers + TMEM slot
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_B_STAGE_FLOATS = 2 * TC_B_PART / 4;      // floats of one pre-tiled (hi | lo) B stage

struct DenseTcArgs {
    const DenseJob* jobs;
    const int32_t* ytab;                     // [grid.y] job index
    const unsigned long long* dig;
    const float* arena_tc;                   // pre-tiled hi/lo CPT stages (see tc_pack_job)
    const float* pl;
    const float* msg_cur;
    float* tscr;                             // [tiles][TS][TBC]
    int32_t PL, M, TS, TBC;
    int64_t n_cases;                         // resident cases (a multiple of 128)
    int32_t stages;
    const uint8_t* status;
};

inline size_t tc_smem_bytes(int n_rows, int stages)
{
    return (size_t)TC_HDR + (size_t)stages * TC_STAGE + (size_t)(n_rows + 1) * TC_M * sizeof(float);
}

// stages that fit next to the factor rows of a job (0: the job cannot take this path)
inline int tc_stages_for(int n_rows, size_t smem_limit)
{
    const size_t fixed = (size_t)TC_HDR + (size_t)(n_rows + 1) * TC_M * sizeof(float);
    if (fixed + 2 * (size_t)TC_STAGE > smem_limit) return 0;
    const int s = (int)((smem_limit - fixed) / TC_STAGE);
    return s > TC_MAX_STAGES ? TC_MAX_STAGES : s;
}

#ifdef BNBP_DENSE_TC_KERNEL   // the kernel itself: only bnbp_dense_tc_inst.cu

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool tc_mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity)
{
    if (tc_mbar_try(bar, parity)) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (uint32_t it = 1;; ++it) {
        if (tc_mbar_try(bar, parity)) return;
        if ((it & 1023u) == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 2000000000ull) __trap();             // protocol bug: fail, do not hang the GPU
        }
    }
}
__device__ __forceinline__ void tc_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices of 128 contiguous
// bytes; SBO = distance between core matrices along M/N, LBO = distance between the two along K
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// D[tmem] (+)= A[smem] * B[smem], one CTA, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = 256
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((128u >> 4) << 24);

__global__ void __launch_bounds__(TC_THREADS, 1) dense_tc_kernel(const DenseTcArgs a)
{
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int S = a.stages;
    const uint32_t sbase = tc_smem_u32(tc_smem);
    const uint32_t bar_full_a = sbase, bar_full_b = sbase + 64, bar_empty = sbase + 128, bar_accum = sbase + 192;
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(tc_smem + 200);
    int* const any_active = reinterpret_cast<int*>(tc_smem + 208);
    const uint32_t stage0 = sbase + TC_HDR;
    float (*Fs)[TC_M] = reinterpret_cast<float (*)[TC_M]>(tc_smem + TC_HDR + (size_t)S * TC_STAGE);

    const DenseJob* __restrict__ jb = a.jobs + a.ytab[blockIdx.y];
    const int K = jb->K, N = jb->N, nf = jb->nf;
    const int ctile = (int)blockIdx.y - jb->y0;
    const int n0 = ctile * TC_N;
    const int KS = (K + TC_K - 1) / TC_K;
    const size_t c0 = (size_t)blockIdx.x * TC_M;
    const bool valid1 = c0 + 128 < (size_t)a.n_cases;            // the last CTA of an odd tile count owns one tile
    const size_t TBC = (size_t)a.TBC;

    if (a.status) {                                               // epsilon mode: skip CTAs whose cases are all frozen
        if (t == 0) *any_active = 0;
        __syncthreads();
        if (t < TC_M && c0 + t < (size_t)a.n_cases && a.status[c0 + t] == 0) *any_active = 1;   // benign race
        __syncthreads();
        if (!*any_active) return;
    }

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < S; ++s) {
            tc_mbar_init(bar_full_a + 8 * s, 8);                  // one arrival per producer warp
            tc_mbar_init(bar_full_b + 8 * s, 1);                  // the loader's expect_tx arrival
            tc_mbar_init(bar_empty + 8 * s, 1);                   // tcgen05.commit
        }
        tc_mbar_init(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // producer / epilogue identity: thread p owns case row m of case tile ct
    const int p = t - 64;
    const int ct = (p >> 7) & 1, m = p & 127;
    const size_t cc = c0 + ((ct && valid1) ? 128 : 0);            // an absent second tile repeats the first (never stored)
    const size_t tile = cc / TBC, lane0 = cc % TBC;
    if (warp >= 2) {
        // stage the factor vectors of this CTA's 256 cases: Fs[row][case]
        int row = 0;
        for (int f = 0; f < nf; ++f) {
            const int s = jb->f_slot[f], card = jb->f_card[f];
            const float* __restrict__ src = s >= 0 ? a.msg_cur + ((size_t)tile * a.M + (size_t)s) * TBC + lane0 + m
                                                   : a.pl + ((size_t)tile * a.PL + (size_t)(-(s + 1))) * TBC + lane0 + m;
            for (int d = 0; d < card; ++d) Fs[row + d][p] = src[(size_t)d * TBC];
            row += card;
        }
        Fs[row][p] = 0.0f;                                        // the zero row: operand rows >= K
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---- B loader: one 16 KB bulk copy per stage out of the pre-tiled arena -------------------
            const float* __restrict__ src = a.arena_tc + jb->b_off + (size_t)ctile * KS * TC_B_STAGE_FLOATS;
            for (int ks = 0; ks < KS; ++ks) {
                const int s = ks % S;
                const uint32_t ph = (uint32_t)(ks / S) & 1u;
                tc_mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                tc_mbar_expect_tx(bar_full_b + 8 * s, 2 * TC_B_PART);
                tc_bulk_g2s(stage0 + s * TC_STAGE + 4 * TC_A_PART, src + (size_t)ks * TC_B_STAGE_FLOATS, 2 * TC_B_PART,
                            bar_full_b + 8 * s);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer: per stage and case tile  lo*hi + hi*lo + hi*hi  into the fp32 accumulator --
            for (int ks = 0; ks < KS; ++ks) {
                const int s = ks % S;
                const uint32_t ph = (uint32_t)(ks / S) & 1u;
                tc_mbar_wait(bar_full_b + 8 * s, ph);
                tc_mbar_wait(bar_full_a + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = stage0 + s * TC_STAGE, sb = sa + 4 * TC_A_PART;
                const uint64_t b_hi = tc_desc(sb, TC_N * 16, 128), b_lo = tc_desc(sb + TC_B_PART, TC_N * 16, 128);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c == 1 && !valid1) break;
                    const uint64_t a_hi = tc_desc(sa + c * 2 * TC_A_PART, 128 * 16, 128);
                    const uint64_t a_lo = tc_desc(sa + c * 2 * TC_A_PART + TC_A_PART, 128 * 16, 128);
                    const uint32_t d = tmem + c * TC_N;
                    tc_mma_tf32(d, a_lo, b_hi, TC_IDESC, ks > 0 ? 1u : 0u);
                    tc_mma_tf32(d, a_hi, b_lo, TC_IDESC, 1u);
                    tc_mma_tf32(d, a_hi, b_hi, TC_IDESC, 1u);
                }
                tc_commit(bar_empty + 8 * s);                     // frees the stage when these MMAs have read it
            }
            tc_commit(bar_accum);                                 // accumulators complete
        }
        __syncwarp();
    } else {
        // ---- A producers: operand rows k0..k0+7 of this thread's case, split hi/lo, MMA layout ---------
        const unsigned long long* __restrict__ dig = a.dig + jb->dig_off;
        const float* const Fcol = &Fs[0][p];
        for (int ks = 0; ks < KS; ++ks) {
            const int s = ks % S;
            const uint32_t ph = (uint32_t)(ks / S) & 1u;
            float v[TC_K];
#pragma unroll
            for (int i = 0; i < TC_K; ++i) {
                const unsigned long long dg = __ldg(dig + ks * TC_K + i);          // warp-uniform
                float va = Fcol[(unsigned)(dg & 255ull) * TC_M];
                for (int f = 1; f < nf; ++f) va *= Fcol[(unsigned)((dg >> (8 * f)) & 255ull) * TC_M];
                v[i] = va;                                                         // rows >= K: the zero row
            }
            float hi[TC_K], lo[TC_K];
#pragma unroll
            for (int i = 0; i < TC_K; ++i) {
                hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
                lo[i] = v[i] - hi[i];
            }
            tc_mbar_wait(bar_empty + 8 * s, ph ^ 1u);
            // element (m, k): + (k / 4) * 2048 + m * 16 + (k % 4) * 4 within a part
            unsigned char* const dst = tc_smem + TC_HDR + (size_t)s * TC_STAGE + (size_t)ct * 2 * TC_A_PART + (size_t)m * 16;
            *reinterpret_cast<float4*>(dst) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(dst + 2048) = make_float4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<float4*>(dst + TC_A_PART) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4*>(dst + TC_A_PART + 2048) = make_float4(lo[4], lo[5], lo[6], lo[7]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic writes -> async proxy (MMA)
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(bar_full_a + 8 * s);
        }

        // ---- epilogue: accumulator rows (TMEM lanes) = cases, columns = result slots -------------------
        tc_mbar_wait(bar_accum, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ct == 0 || valid1) {
            const int q = warp & 3;                               // the TMEM lane quarter this warp may read
            const int rowc = 32 * q + lane;                       // case row within the 128-case tile
            const size_t ce = c0 + (size_t)ct * 128;
            float* const tb = a.tscr + ((size_t)(ce / TBC) * a.TS + (size_t)jb->t_off) * TBC + ce % TBC + rowc;
            for (int cb = 0; cb < TC_N / 32; ++cb) {
                const int nb = n0 + cb * 32;
                if (nb >= N) break;                               // warp-uniform
                float v[32];
                tc_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(ct * TC_N + cb * 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (nb + j < N) tb[(size_t)(nb + j) * TBC] = v[j];
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

#endif // BNBP_DENSE_TC_KERNEL

// Host side: the (hi | lo) stage images of one job's B matrix.  B[k][n] = src[k * sk + n * sn]
// (GEMM 1: the reference CPT layout, sk = N, sn = 1; GEMM 2: the same table read transposed).
// Stage (column tile j, step ks) starts at float  (j * KS + ks) * TC_B_STAGE_FLOATS; element (n, k)
// of its hi part sits at  ((k % 8) / 4) * 1024 + (n % 256) * 4 + k % 4, the lo part 2048 floats later.
inline int64_t tc_job_floats(int K, int N)
{
    return (int64_t)((N + TC_N - 1) / TC_N) * ((K + TC_K - 1) / TC_K) * TC_B_STAGE_FLOATS;
}

inline void tc_split(float v, float& hi, float& lo)
{
    uint32_t u;
    memcpy(&u, &v, 4);
    u &= 0xFFFFE000u;
    memcpy(&hi, &u, 4);
    lo = v - hi;
}

inline void tc_pack_job(float* out, const double* src, int K, int N, int64_t sk, int64_t sn)
{
    const int KS = (K + TC_K - 1) / TC_K, NT = (N + TC_N - 1) / TC_N;
    for (int j = 0; j < NT; ++j)
        for (int ks = 0; ks < KS; ++ks) {
            float* st = out + ((size_t)j * KS + ks) * TC_B_STAGE_FLOATS;
            for (int kk = 0; kk < TC_K; ++kk) {
                const int k = ks * TC_K + kk;
                for (int nn = 0; nn < TC_N; ++nn) {
                    const int n = j * TC_N + nn;
                    float hi = 0.0f, lo = 0.0f;
                    if (k < K && n < N) tc_split((float)src[(int64_t)k * sk + (int64_t)n * sn], hi, lo);
                    const size_t e = (size_t)(kk / 4) * 1024 + (size_t)nn * 4 + (size_t)(kk % 4);
                    st[e] = hi;
                    st[e + TC_B_PART / 4] = lo;
                }
            }
        }
}

cudaError_t launch_dense_tc(const DenseTcArgs& a, dim3 grid, size_t smem, cudaStream_t st);
cudaError_t set_dense_tc_smem(int bytes);

} // namespace bnbp
