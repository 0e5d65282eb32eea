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

template <typename T> struct DenseSmem {
    T As[2][DT_K][DT_M];
    T Bs[2][DT_K][DT_N];
    // followed by Fs[n_rows + 1][DT_M]
};

template <typename T, int NF>
__global__ void __launch_bounds__(DT_THREADS, sizeof(T) == 8 ? 1 : 2)
dense_gemm_kernel(const DenseArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char dense_smem_raw[];
    DenseSmem<T>& sm = *reinterpret_cast<DenseSmem<T>*>(dense_smem_raw);
    T (*Fs)[DT_M] = reinterpret_cast<T (*)[DT_M]>(dense_smem_raw + sizeof(DenseSmem<T>));
    __shared__ int any_active;

    const int t = threadIdx.x;
    const DenseJob* __restrict__ jb = a.jobs + a.ytab[blockIdx.y];
    const int K = jb->K, N = jb->N;
    const int n0 = ((int)blockIdx.y - jb->y0) * DT_N;
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

    const int lm = t & (DT_M - 1);         // loader: column (case / result column) ...
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
    constexpr int LD = DT_K / 2;           // elements of each operand this thread stages per chunk
    const int tx = t & 15, ty = t >> 4;

    T acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = T(0);

    const unsigned long long* __restrict__ dig = a.dig + jb->dig_off + lk;
    const T* __restrict__ Bcol = Bg + min(n0 + lm, N - 1);     // columns >= N: computed, never stored
    const T* const Fcol = &Fs[0][lm];
    T ra[LD], rb[LD];
    // operand rows of the next stage: issued before the FMA block of the current stage, consumed
    // (stash) after it, so the L2 latency of the CPT loads hides behind the arithmetic
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < LD; ++i) {
            const int kk = k0 + lk + 2 * i;
            const unsigned long long dg = __ldg(dig + k0 + 2 * i);     // warp-uniform (a warp shares lk)
            T va = Fcol[(unsigned)(dg & 255ull) * DT_M];
#pragma unroll
            for (int f = 1; f < NF; ++f) va *= Fcol[(unsigned)((dg >> (8 * f)) & 255ull) * DT_M];
            ra[i] = va;                                                // rows >= K: the zero row
            rb[i] = __ldg(Bcol + (size_t)min(kk, K - 1) * N);
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < LD; ++i) {
            sm.As[buf][lk + 2 * i][lm] = ra[i];
            sm.Bs[buf][lk + 2 * i][lm] = rb[i];
        }
    };

    const int n_chunks = (K + DT_K - 1) / DT_K;
    __syncthreads();                       // Fs complete
    fetch(0);
    stash(0);
    __syncthreads();
    for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        fetch((c + 1) * DT_K);             // past the end: zero rows (the digit table is padded)
#pragma unroll
        for (int kk = 0; kk < DT_K; ++kk) {
            T av[8], bv[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                av[i] = sm.As[buf][kk][tx * 4 + i];
                av[4 + i] = sm.As[buf][kk][64 + tx * 4 + i];
                bv[i] = sm.Bs[buf][kk][ty * 4 + i];
                bv[4 + i] = sm.Bs[buf][kk][64 + ty * 4 + i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        stash(buf ^ 1);
        __syncthreads();
    }

    // T[tile][t_off + n][TBC]: for a fixed column the 16 tx-lanes write 64 consecutive cases
    T* const tb = a.tscr + ((size_t)tile * a.TS + (size_t)jb->t_off) * TBC + lane0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = n0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
        if (n < N) {
            T* const col = tb + (size_t)n * TBC;
            Pk<T, 4> lo, hi;
#pragma unroll
            for (int i = 0; i < 4; ++i) { lo.v[i] = acc[i][j]; hi.v[i] = acc[4 + i][j]; }
            stp<T, 4>(col + tx * 4, lo);
            stp<T, 4>(col + 64 + tx * 4, hi);
        }
    }
}

// one launch per distinct factor count NF (the A-tile product is unrolled over it)
template <typename T>
cudaError_t launch_dense(const DenseArgs<T>& a, int nf, dim3 grid, size_t smem, cudaStream_t st)
{
    switch (nf) {
#define BNBP_DENSE_CASE(F) case F: dense_gemm_kernel<T, F><<<grid, DT_THREADS, smem, st>>>(a); break;
        BNBP_DENSE_CASE(1) BNBP_DENSE_CASE(2) BNBP_DENSE_CASE(3) BNBP_DENSE_CASE(4)
        BNBP_DENSE_CASE(5) BNBP_DENSE_CASE(6) BNBP_DENSE_CASE(7) BNBP_DENSE_CASE(8)
#undef BNBP_DENSE_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <typename T> cudaError_t set_dense_smem(int bytes)
{
    cudaError_t e = cudaSuccess;
#define BNBP_DENSE_CASE(F) \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dense_gemm_kernel<T, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    BNBP_DENSE_CASE(1) BNBP_DENSE_CASE(2) BNBP_DENSE_CASE(3) BNBP_DENSE_CASE(4)
    BNBP_DENSE_CASE(5) BNBP_DENSE_CASE(6) BNBP_DENSE_CASE(7) BNBP_DENSE_CASE(8)
#undef BNBP_DENSE_CASE
    return e;
}

} // namespace bnbp
