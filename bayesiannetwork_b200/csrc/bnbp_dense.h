// bnbp_dense.h — host-visible part of the dense contraction path (kernel: bnbp_dense.cuh).
#pragma once
#include "bnbp_kernels.cuh"

namespace bnbp {

constexpr int DENSE_MAXF = 8;      // factors of one operand row: parents of one group (+ lambda_X)
constexpr int DENSE_MAXROWS = 160; // staged factor rows per job (sum of the factor cardinalities), + 1 zero row
constexpr int DT_M = 128;          // cases per CTA
constexpr int DT_N = 128;          // result columns per CTA of the widest tile class (16 * TNT, TNT = 8 | 4 | 2)
constexpr int DT_K = 8;            // contraction depth per shared-memory stage
constexpr int DT_THREADS = 256;    // 16 x 16 threads, 8 x 8 results each

struct DenseJob {                  // one GEMM: T[c][n] = sum_k (prod_f F_f[c][digit_f(k)]) * B[k][n]
    int64_t b_off;                 // offset of the row-major [K][N] matrix in its arena
    int64_t dig_off;               // first entry of the job in the digit table (see DenseArgs::dig)
    int32_t arena;                 // 0: reference-layout CPT arena (GEMM 1), 1: transposed copies (GEMM 2)
    int32_t K, N;
    int32_t t_off;                 // slot of T[.][0] in the per-case contraction scratch
    int32_t y0;                    // first blockIdx.y of this job within its launch (one per 16*TNT columns)
    int32_t nf;
    int32_t n_rows;                // staged factor rows = sum_f f_card[f]; row n_rows is all zeros
    int32_t f_slot[DENSE_MAXF];    // >= 0: message slot in msg_cur; < 0: slot -(v+1) of the pi/lambda region
    int32_t f_card[DENSE_MAXF];
};

template <typename T> struct DenseArgs {
    const DenseJob* jobs;
    const int32_t* ytab;           // [grid.y] job index
    // byte f of dig[dig_off + k] = staged row of factor f for operand row k (first row of the factor
    // + its mixed-radix digit); rows K .. (padded to DT_K) + DT_K point every factor at the zero row
    const unsigned long long* dig;
    const T* arena0;
    const T* arena1;
    const T* pl;
    const T* msg_cur;
    T* tscr;                       // [tiles][TS][TBC]
    int32_t PL, M, TS, TBC;
    const uint8_t* status;         // != nullptr: skip 128-case tiles whose cases are all frozen
};

// bytes of dynamic shared memory of one CTA whose job stages n_rows factor rows
inline size_t dense_smem_bytes(int n_rows, int tnt, size_t tsize)
{
    // operand stages (row strides padded by 8 for the DMMA variant; the bound covers both) + factor rows
    return (size_t)2 * DT_K * (DT_M + 16 * tnt + 16) * tsize + (size_t)(n_rows + 1) * DT_M * tsize;
}

// column-tile class of a product with N result columns: 8 x TNT results per thread
inline int dense_tnt(int N) { return N > 64 ? 8 : (N > 32 ? 4 : 2); }

// (NF, TNT) instantiations
#define BNBP_DENSE_VARIANTS(X) \
    X(1, 8) X(2, 8) X(3, 8) X(4, 8) X(5, 8) X(6, 8) X(7, 8) X(8, 8) \
    X(1, 4) X(2, 4) X(3, 4) X(4, 4) X(5, 4) X(6, 4) X(7, 4) X(8, 4) \
    X(1, 2) X(2, 2) X(3, 2) X(4, 2) X(5, 2) X(6, 2) X(7, 2) X(8, 2)

// one launch per (factor count NF, column-tile class TNT); defined in
// bnbp_dense_inst.cu for T = double, float
template <typename T>
cudaError_t launch_dense(const DenseArgs<T>& a, int nf, int tnt, bool mma, dim3 grid, size_t smem, cudaStream_t st);
template <typename T> cudaError_t set_dense_smem(int bytes);

} // namespace bnbp
