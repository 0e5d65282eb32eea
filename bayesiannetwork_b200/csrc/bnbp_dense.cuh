// bnbp_dense.cuh — dense contraction path for nodes with LARGE conditional probability tables.
//
// For a node X with parents U_1..U_k the parent side of a sweep (calculate_pi, belief_propagation.hpp
// :174-200, and calculate_lambda_k :240-266) is a contraction of the CPT with the k incoming
// pi-messages and lambda_X.  The per-thread kernels walk the whole CPT once per CASE with broadcast
// loads; for a 32^4-entry CPT (config "card32") that is 1 M loads per case and node.  Here the
// contraction is regrouped so that the CPT meets the whole BATCH in a matrix product:
//
//   split the parents into A = (U_1..U_s) [slow CPT digits] and B = (U_s+1..U_k) [fast digits];
//   the reference CPT layout cpt[q*r + x] (q mixed radix, first parent slowest, :269-295) IS the
//   row-major matrix  P[uA][(uB,x)]  of shape  QA x (QB*r)  -- no permutation needed.
//
//   GEMM 1:  T1[c][(uB,x)] = sum_uA  WA[c][uA] * P[uA][(uB,x)],   WA[c][uA]     = prod_{j<=s} m_j[c](u_j)
//   GEMM 2:  T2[c][uA]     = sum_(uB,x) WB[c][(uB,x)] * P[uA][(uB,x)], WB[c][(uB,x)] = lambda_X[c](x) prod_{j>s} m_j[c](u_j)
//
// T1 is a per-case CPT of X over the parents in B only, T2 a per-case "lambda table" over the
// parents in A; the sweep kernel finishes both with its ordinary recursion (bnbp_sweep.cuh, dense
// branch), so normalisation, evidence, freezing, damping and the delta scan stay in ONE place.
// Flops: 4*|CPT| per case and node (the one-pass recursion needs (2k+2)*|CPT|, the reference
// (k^2+k+1)*|CPT|), and the CPT is read once per 128-case tile out of L2 instead of once per case.
//
// The operand rows WA / WB are never materialised: each CTA stages the message vectors of its 128
// cases in shared memory once and forms every A-tile on the fly (a product of NF <= DENSE_MAXF
// factors per element, factor rows looked up through a per-job digit table).  Arithmetic is the
// handle's precision on the CUDA cores (fp64 parity bound 1e-9: no tensor-core format holds it).
#pragma once
#include "bnbp_dense.h"

namespace bnbp {

// Tile shapes: 128 cases x (16 * TNT) result columns per CTA, 256 threads as 16 (cases) x 16 (columns),
// 8 x TNT results per thread.  TNT = 8 for wide products; 4 / 2 for the small products of
// mid-sized CPTs (N = 9..64 columns), where a 128-column tile would be mostly padding.
// Operand elements of one thread are interleaved in 16-byte groups (2 doubles / 4 floats) so that a
// quarter-warp's 128-bit shared-memory loads hit consecutive addresses (conflict-free).
template <typename T, int TNT, bool MMA> struct DenseShape {
    static constexpr int TN = 16 * TNT;                            // columns per CTA
    // row strides of the staged operand tiles; the DMMA fragments read 4 rows x 8 columns per
    // request, +8 elements of padding spread the 4 rows over both halves of the banks
    static constexpr int SA = DT_M + (MMA ? 8 : 0);
    static constexpr int SB = TN + (MMA ? 8 : 0);
    static constexpr int VW = 16 / (int)sizeof(T);                 // elements per 128-bit access
    static constexpr int VN = TNT < VW ? TNT : VW;                 // vector width on the column side
    static constexpr int LDB = DT_K * TN / DT_THREADS;             // B elements staged per thread and chunk
    static_assert(LDB >= 1, "column tile too narrow for the loader");
    static constexpr size_t tiles_bytes = (size_t)2 * DT_K * (SA + SB) * sizeof(T);
};

// D(8x8) += A(8x4, row) * B(4x8, col) in fp64 on the tensor pipe (DMMA).  Fragments (PTX ISA,
// mma.m8n8k4 .f64): a = A[lane/4][lane%4], b = B[lane%4][lane/4], d0/d1 = D[lane/4][2*(lane%4) + 0/1].
__device__ __forceinline__ void dmma_m8n8k4(double (&d)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <typename T, int NF, int TNT, bool MMA>
__global__ void __launch_bounds__(DT_THREADS, (sizeof(T) == 8 && TNT == 8) ? 1 : 2)
dense_gemm_kernel(const DenseArgs<T> a)
{
    static_assert(!MMA || sizeof(T) == 8, "the tensor-pipe variant is the fp64 DMMA path");
    using SH = DenseShape<T, TNT, MMA>;
    constexpr int TN = SH::TN, VW = SH::VW, VN = SH::VN, LDB = SH::LDB, SA = SH::SA, SB = SH::SB;
    extern __shared__ __align__(16) unsigned char dense_smem_raw[];
    T (*As)[DT_K][SA] = reinterpret_cast<T (*)[DT_K][SA]>(dense_smem_raw);                           // [2]
    T (*Bs)[DT_K][SB] = reinterpret_cast<T (*)[DT_K][SB]>(dense_smem_raw + 2 * DT_K * SA * sizeof(T)); // [2]
    T (*Fs)[DT_M] = reinterpret_cast<T (*)[DT_M]>(dense_smem_raw + SH::tiles_bytes);                 // [n_rows + 1]
    __shared__ int any_active;

    const int t = threadIdx.x;
    const DenseJob* __restrict__ jb = a.jobs + a.ytab[blockIdx.y];
    const int K = jb->K, N = jb->N;
    const int n0 = ((int)blockIdx.y - jb->y0) * TN;
    const size_t c0 = (size_t)blockIdx.x * DT_M;
    const size_t TBC = (size_t)a.TBC;
    const size_t tile = c0 / TBC, lane0 = c0 % TBC;

    if (a.status) {
        if (t == 0) any_active = 0;
        __syncthreads();
        if (t < DT_M && a.status[c0 + t] == 0) any_active = 1;     // benign race: same value
        __syncthreads();
        if (!any_active) return;
    }

    const int lm = t & (DT_M - 1);         // A loader: case ...
    const int lk = t >> 7;                 // ... and row parity within the stage
    {   // stage the factor vectors of this tile's 128 cases: Fs[row][case]
        int row = 0;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const int s = jb->f_slot[f], card = jb->f_card[f];
            const T* __restrict__ src = s >= 0 ? a.msg_cur + ((size_t)tile * a.M + (size_t)s) * TBC + lane0
                                               : a.pl + ((size_t)tile * a.PL + (size_t)(-(s + 1))) * TBC + lane0;
            for (int d = lk; d < card; d += 2) Fs[row + d][lm] = src[(size_t)d * TBC + lm];
            row += card;
        }
        if (lk == 0) Fs[row][lm] = T(0);
    }

    const T* __restrict__ Bg = (jb->arena ? a.arena1 : a.arena0) + jb->b_off;
    constexpr int LDA = DT_K / 2;          // A elements this thread stages per chunk

    const unsigned long long* __restrict__ dig = a.dig + jb->dig_off + lk;
    const T* const Fcol = &Fs[0][lm];
    // B loader: element e = t + 256*i of the [DT_K][TN] stage
    const T* bsrc[LDB];
    int brow[LDB], bcol[LDB];
#pragma unroll
    for (int i = 0; i < LDB; ++i) {
        const int e = t + DT_THREADS * i;
        brow[i] = e / TN;
        bcol[i] = e % TN;
        bsrc[i] = Bg + min(n0 + bcol[i], N - 1);                   // columns >= N: computed, never stored
    }
    T ra[LDA], rb[LDB];
    // operand rows of the next stage: issued before the FMA block of the current stage, consumed
    // (stash) after it, so the L2 latency of the CPT loads hides behind the arithmetic
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < LDA; ++i) {
            const unsigned long long dg = __ldg(dig + k0 + 2 * i);     // warp-uniform (a warp shares lk)
            T va = Fcol[(unsigned)(dg & 255ull) * DT_M];
#pragma unroll
            for (int f = 1; f < NF; ++f) va *= Fcol[(unsigned)((dg >> (8 * f)) & 255ull) * DT_M];
            ra[i] = va;                                                // rows >= K: the zero row
        }
#pragma unroll
        for (int i = 0; i < LDB; ++i) rb[i] = __ldg(bsrc[i] + (size_t)min(k0 + brow[i], K - 1) * N);
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < LDA; ++i) As[buf][lk + 2 * i][lm] = ra[i];
#pragma unroll
        for (int i = 0; i < LDB; ++i) Bs[buf][brow[i]][bcol[i]] = rb[i];
    };
    const int n_chunks = (K + DT_K - 1) / DT_K;
    T* const tb = a.tscr + ((size_t)tile * a.TS + (size_t)jb->t_off) * TBC + lane0;

    if constexpr (!MMA) {
        // ---- CUDA-core variant: 16 x 16 threads, 8 x TNT results each ---------------------------
        const int tx = t & 15, ty = t >> 4;
        T acc[8][TNT];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < TNT; ++j) acc[i][j] = T(0);
        // element i of this thread along the case / column axis
        auto m_of = [&](int i) { return (i / VW) * (16 * VW) + tx * VW + i % VW; };
        auto n_of = [&](int j) { return (j / VN) * (16 * VN) + ty * VN + j % VN; };

        __syncthreads();                       // Fs complete
        fetch(0);
        stash(0);
        __syncthreads();
        for (int c = 0; c < n_chunks; ++c) {
            const int buf = c & 1;
            fetch((c + 1) * DT_K);             // past the end: zero rows (the digit table is padded)
#pragma unroll
            for (int kk = 0; kk < DT_K; ++kk) {
                T av[8], bv[TNT];
#pragma unroll
                for (int i = 0; i < 8; ++i) av[i] = As[buf][kk][m_of(i)];
#pragma unroll
                for (int j = 0; j < TNT; ++j) bv[j] = Bs[buf][kk][n_of(j)];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < TNT; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
            }
            stash(buf ^ 1);
            __syncthreads();
        }

        // T[tile][t_off + n][TBC]: for a fixed column the 16 tx-lanes write runs of consecutive cases
#pragma unroll
        for (int j = 0; j < TNT; ++j) {
            const int n = n0 + n_of(j);
            if (n < N) {
                T* const col = tb + (size_t)n * TBC;
#pragma unroll
                for (int g = 0; g < 8 / VW; ++g) {
                    Pk<T, VW> o;
#pragma unroll
                    for (int i = 0; i < VW; ++i) o.v[i] = acc[g * VW + i][j];
                    stp<T, VW>(col + m_of(g * VW), o);
                }
            }
        }
    } else {
        // ---- tensor-pipe variant (fp64 DMMA m8n8k4): 8 warps as 2 (cases) x 4 (columns), warp tile
        //      64 cases x TN/4 columns = 8 x NJ fragments of 8 x 8 ------------------------------------
        constexpr int NJ = TN / 32;
        const int lane = t & 31, w = t >> 5;
        const int g = lane >> 2, q = lane & 3;
        const int wm = (w & 1) * 64, wn = (w >> 1) * (8 * NJ);
        double acc[8][NJ][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        __syncthreads();                       // Fs complete
        fetch(0);
        stash(0);
        __syncthreads();
        for (int c = 0; c < n_chunks; ++c) {
            const int buf = c & 1;
            fetch((c + 1) * DT_K);
#pragma unroll
            for (int k4 = 0; k4 < DT_K / 4; ++k4) {
                double av[8], bv[NJ];
#pragma unroll
                for (int i = 0; i < 8; ++i) av[i] = As[buf][k4 * 4 + q][wm + 8 * i + g];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bv[j] = Bs[buf][k4 * 4 + q][wn + 8 * j + g];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j], av[i], bv[j]);
            }
            stash(buf ^ 1);
            __syncthreads();
        }
        // fragment (i, j): cases wm + 8i + g, columns wn + 8j + 2q + {0,1}
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = n0 + wn + 8 * j + 2 * q + e;
                if (n < N) {
                    T* const col = tb + (size_t)n * TBC + wm + g;
#pragma unroll
                    for (int i = 0; i < 8; ++i) col[8 * i] = acc[i][j][e];
                }
            }
    }
}

// one launch per (factor count NF, column-tile class TNT): the A-tile product is unrolled over NF.
// mma: the fp64 DMMA variant (ignored for float).
template <typename T>
cudaError_t launch_dense(const DenseArgs<T>& a, int nf, int tnt, bool mma, dim3 grid, size_t smem, cudaStream_t st)
{
    constexpr bool F64 = sizeof(T) == 8;
#define BNBP_DENSE_CASE(F, W)                                                                               \
    if (nf == F && tnt == W) {                                                                              \
        if (F64 && mma) dense_gemm_kernel<T, F, W, F64><<<grid, DT_THREADS, smem, st>>>(a);                 \
        else dense_gemm_kernel<T, F, W, false><<<grid, DT_THREADS, smem, st>>>(a);                          \
        return cudaGetLastError();                                                                          \
    }
    BNBP_DENSE_VARIANTS(BNBP_DENSE_CASE)
#undef BNBP_DENSE_CASE
    return cudaErrorInvalidValue;
}

template <typename T> cudaError_t set_dense_smem(int bytes)
{
    constexpr bool F64 = sizeof(T) == 8;
    cudaError_t e = cudaSuccess;
#define BNBP_DENSE_CASE(F, W)                                                                                                   \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dense_gemm_kernel<T, F, W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    if (e == cudaSuccess && F64) e = cudaFuncSetAttribute(dense_gemm_kernel<T, F, W, F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    BNBP_DENSE_VARIANTS(BNBP_DENSE_CASE)
#undef BNBP_DENSE_CASE
    return e;
}

} // namespace bnbp
