// cuda_host_shim.h -- just enough of the CUDA C++ surface for g++ to compile the source the network compiler
// generates (bnbp_spec_source: trait structs + bnbp_spec.cuh) as ORDINARY HOST CODE.  Test infrastructure only
// (tests/test_netcompiler_emul.py): the streaming sweep kernel gives every case to one thread and no thread ever
// reads what another one wrote, so calling the kernel function once per (block, thread) on the host IS the kernel --
// addressing, the generated offset tables, the node arithmetic and the walk order are checked against the oracle
// without a GPU.  What it cannot show: anything about timing, and the warp-level parts (variants 6/7, the on-chip
// kernel), which are not compiled here.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __global__
#define __constant__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static

struct emul_dim3 { unsigned x = 0, y = 0, z = 0; };
static emul_dim3 threadIdx, blockIdx, gridDim;

template <typename U> static inline U __ldg(const U* p) { return *p; }
static inline double __dmul_rn(double a, double b) { return a * b; }      // built with -ffp-contract=off
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline unsigned __activemask() { return 1u << (threadIdx.x & 31); } // one lane at a time: the calling one is the only live lane
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline void __syncwarp() {}
// inline PTX exists only in code paths of the on-chip kernel (discarded `if constexpr` branches here)
#define asm(...) ((void)0)
