// Explicit instantiations of the dense contraction kernels (bnbp_dense.cuh), one translation unit.
#include "bnbp_dense.cuh"
namespace bnbp {
static_assert(DenseShape<double, 8, true>::tiles_bytes <= (size_t)2 * DT_K * (DT_M + DT_N + 16) * 8, "dense_smem_bytes out of sync");
template cudaError_t launch_dense<double>(const DenseArgs<double>&, int, int, bool, dim3, size_t, cudaStream_t);
template cudaError_t launch_dense<float>(const DenseArgs<float>&, int, int, bool, dim3, size_t, cudaStream_t);
template cudaError_t set_dense_smem<double>(int);
template cudaError_t set_dense_smem<float>(int);
} // namespace bnbp
