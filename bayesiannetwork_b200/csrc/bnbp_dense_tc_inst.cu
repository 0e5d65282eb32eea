// Translation unit of the tensor-core dense contraction kernel (bnbp_dense_tc.cuh).
#include <cstring>
#define BNBP_DENSE_TC_KERNEL
#include "bnbp_dense_tc.cuh"
namespace bnbp {
cudaError_t launch_dense_tc(const DenseTcArgs& a, dim3 grid, size_t smem, cudaStream_t st)
{
    dense_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}
cudaError_t set_dense_tc_smem(int bytes)
{
    return cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
} // namespace bnbp
