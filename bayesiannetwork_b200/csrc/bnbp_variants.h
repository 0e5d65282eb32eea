// bnbp_variants.h — the (VEC, RMAX, KNET) instantiations of the sweep kernel family.
//   VEC   cases per thread (vector width of the state accesses)
//   RMAX  unroll bound >= the largest cardinality of the network
//   KNET  unroll bound >= the largest in-degree of the network
// One translation unit per line and element type (see ../_build.py, which parses this list).
#pragma once
#define BNBP_SWEEP_VARIANTS(X, T) \
    X(T, 1, 2, 2) X(T, 1, 2, 4) X(T, 1, 2, 8) \
    X(T, 2, 2, 2) X(T, 2, 2, 4) X(T, 2, 2, 8) \
    X(T, 1, 4, 2) X(T, 1, 4, 4) X(T, 1, 4, 8) \
    X(T, 2, 4, 2) X(T, 2, 4, 4) X(T, 2, 4, 8) \
    X(T, 1, 8, 4) X(T, 1, 8, 8) X(T, 2, 8, 4) X(T, 2, 8, 8) \
    X(T, 1, 16, 8) X(T, 1, 32, 8) X(T, 1, 64, 8) X(T, 1, 128, 8)
