// One explicit instantiation of the sweep kernel family per translation unit:
//   nvcc -DBNBP_T=double -DBNBP_VEC=2 -DBNBP_RMAX=4 -DBNBP_KNET=4 -c bnbp_sweep_inst.cu
#include "bnbp_sweep.cuh"
#if !defined(BNBP_T) || !defined(BNBP_VEC) || !defined(BNBP_RMAX) || !defined(BNBP_KNET)
#error "define BNBP_T, BNBP_VEC, BNBP_RMAX and BNBP_KNET"
#endif
namespace bnbp {
template cudaError_t launch_sweep_vr<BNBP_T, BNBP_VEC, BNBP_RMAX, BNBP_KNET>(const SweepArgs<BNBP_T>&, dim3, size_t, bool, bool, cudaStream_t, bool);
template cudaError_t set_sweep_smem<BNBP_T, BNBP_VEC, BNBP_RMAX, BNBP_KNET>(int);
} // namespace bnbp
