// bnbp_sweep.cuh — the sweep kernel (one launch = one iteration of the reference's while(true),
// belief_propagation.hpp:75-148, for every resident case) and its launchers.
//
// Thread = VEC cases, block = 128 threads = one tile of TBC = 128*VEC cases, grid.y = node chunks.
// Per node X the thread gathers everything X reads (pi_X, lambda_X, pi-messages from the parents,
// lambda-messages from the children -- all time t), and produces everything X owns at time t+1:
//   lambda_X (:220-238), pi-messages to the children (:202-218)            "child side"
//   pi_X (:174-200), lambda-messages to the parents (:240-266)             "parent side"
// Normalisation (:298-311) multiplies by one reciprocal of the plain sum per vector (the reference
// divides entry by entry; the difference is <= 1 ulp, far inside the 1e-9 parity bound, and fp64
// division was 30 % of the instruction stream in profiles/r01a).
#pragma once
#include "bnbp_kernels.cuh"

namespace bnbp {

constexpr int BLOCK = 128;

// ------------------------------------------------------------------------------------------------
// Parent side of one node: pi_X and all lambda-messages X->U_j from ONE pass over the CPT.
//   w(u)      = sum_x lambda_X(x) P(x|u)
//   pi_X(x)   = sum_u P(x|u) prod_j m_j(u_j)
//   lmsg_j(a) = sum_{u:u_j=a} w(u) prod_{i!=j} m_i(u_i)
// evaluated by a depth-K recursion over the parents (compile-time depth, runtime cardinalities):
// each level passes down the prefix product P and returns the message-weighted sum R of its
// subtree, so the leave-one-out product for level j is simply P*R -- (2k+2)|CPT| flops instead
// of the reference's (k^2+k+1)|CPT|.  The last parent and the states of X are unrolled to RMAX
// and live in registers; parents further out keep their messages / accumulators in a per-thread
// shared-memory scratch column (dynamic index, conflict-free [value][thread] layout).
// MAXP: the max-product semiring (opt-in extension, bnbp_run_params.semiring): every sum over parent configurations /
// child states of :174-200 and :240-266 becomes a maximum -- all factors are non-negative, so the prefix-product
// recursion below carries over with acc = max(acc, a * b) in place of acc += a * b.  fmax ignores NaN.
template <bool MAXP, typename T> __device__ __forceinline__ T acc_op(T a, T b, T acc)
{
    if constexpr (MAXP) return fmax(acc, a * b);
    else return fma(a, b, acc);
}

template <typename T, int VEC, int RMAX, bool MAXP = false> struct ParentCtx {
    const T* cpt;
    int r, rk;
    int rj[KMAX];
    int soff[KMAX];
    T* scr;            // this thread's scratch column: scr[(i*VEC+v)*BLOCK]
    int sacc_base;     // value index where the outer accumulators start
    T lam[RMAX][VEC], pacc[RMAX][VEC], mk[RMAX][VEC], lacck[RMAX][VEC];
};

// SRC selects where the table comes from:
//   SRC_CPT    the node's CPT in the device arena (warp-uniform broadcast loads)
//   SRC_T1     a PER-CASE table T1[(uB,x)] produced by the dense contraction (bnbp_dense.cuh, GEMM 1):
//              same recursion over the parents of group B, one coalesced load per entry
//   SRC_T2     the per-case table T2[uA] of GEMM 2: r = 1 and lambda = 1, only lambda-messages come out
enum { SRC_CPT = 0, SRC_T1 = 1, SRC_T2 = 2 };

template <int SRC, typename T, int VEC, int RMAX, bool MAXP>
__device__ __forceinline__ void parent_leaf(ParentCtx<T, VEC, RMAX, MAXP>& c, const T (&P)[VEC], int q, T (&ret)[VEC])
{
    constexpr size_t STR = SRC == SRC_CPT ? 1 : (size_t)BLOCK * VEC;   // per-case tables are batch-minor
    const int rr = SRC == SRC_T2 ? 1 : c.r;
    const T* blk = c.cpt + (size_t)q * (size_t)(c.rk * rr) * STR;
#pragma unroll
    for (int v = 0; v < VEC; ++v) ret[v] = T(0);
    auto body = [&](int b) {
        const T* row = blk + (size_t)(b * rr) * STR;
        T w[VEC], pm[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { w[v] = T(0); pm[v] = P[v] * c.mk[b][v]; }
        auto inner = [&](int x) {
            T p[VEC];
            if constexpr (SRC == SRC_CPT) {
                const T pu = __ldg(row + x);
#pragma unroll
                for (int v = 0; v < VEC; ++v) p[v] = pu;
            } else {
                const Pk<T, VEC> pk = ldp<T, VEC>(row + (size_t)x * STR);
#pragma unroll
                for (int v = 0; v < VEC; ++v) p[v] = pk.v[v];
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if constexpr (SRC == SRC_T2) {
                    w[v] = p[v];
                } else {
                    w[v] = acc_op<MAXP>(c.lam[x][v], p[v], w[v]);
                    c.pacc[x][v] = acc_op<MAXP>(p[v], pm[v], c.pacc[x][v]);
                }
            }
        };
        if constexpr (SRC == SRC_T2) {
            inner(0);
        } else if constexpr (RMAX <= 8) {
#pragma unroll
            for (int x = 0; x < RMAX; ++x)
                if (x < c.r) inner(x);
        } else {
            for (int x = 0; x < c.r; ++x) inner(x);     // large cardinalities: arrays live in local memory
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            c.lacck[b][v] = acc_op<MAXP>(P[v], w[v], c.lacck[b][v]);
            ret[v] = acc_op<MAXP>(c.mk[b][v], w[v], ret[v]);
        }
    };
    if constexpr (RMAX <= 8) {
#pragma unroll
        for (int b = 0; b < RMAX; ++b)
            if (b < c.rk) body(b);
    } else {
        for (int b = 0; b < c.rk; ++b) body(b);
    }
}

template <int LEVEL, int K, int SRC, typename T, int VEC, int RMAX, bool MAXP>
__device__ __forceinline__ void parent_rec(ParentCtx<T, VEC, RMAX, MAXP>& c, const T (&P)[VEC], int q, T (&ret)[VEC])
{
    if constexpr (LEVEL == K - 1) {
        parent_leaf<SRC, T, VEC, RMAX, MAXP>(c, P, q, ret);
    } else {
        const int rl = c.rj[LEVEL];
#pragma unroll
        for (int v = 0; v < VEC; ++v) ret[v] = T(0);
        for (int a = 0; a < rl; ++a) {
            T mv[VEC], P2[VEC], R[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                mv[v] = c.scr[((c.soff[LEVEL] + a) * VEC + v) * BLOCK];
                P2[v] = P[v] * mv[v];
            }
            parent_rec<LEVEL + 1, K, SRC, T, VEC, RMAX, MAXP>(c, P2, q * rl + a, R);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                T* acc = &c.scr[((c.sacc_base + c.soff[LEVEL] + a) * VEC + v) * BLOCK];
                *acc = acc_op<MAXP>(P[v], R[v], *acc);
                ret[v] = acc_op<MAXP>(mv[v], R[v], ret[v]);
            }
        }
    }
}

template <int K, int SRC, typename T, int VEC, int RMAX, bool MAXP>
__device__ __forceinline__ void parent_run(ParentCtx<T, VEC, RMAX, MAXP>& c)
{
    T one[VEC], ret[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) one[v] = T(1);
    parent_rec<0, K, SRC, T, VEC, RMAX, MAXP>(c, one, 0, ret);
}

// One group of consecutive parents [j0, j0+kk) of a node against one table: stages their
// pi-messages (outer ones in the scratch column, the last one in registers), runs the recursion and
// emits the lambda-messages to exactly these parents.  An ordinary node is one group over its CPT;
// a dense node is group A over T2 followed by group B over T1.
template <int SRC, int KNET, typename T, int VEC, int RMAX, bool MAXP, typename EmitMsg>
__device__ __forceinline__ void parent_group(ParentCtx<T, VEC, RMAX, MAXP>& pc, const T* table, const int32_t* ecard,
                                             const int32_t* louts, const T* cur, T* scr, int scr_half, int kk,
                                             int slot, EmitMsg& emit_msg)
{
    constexpr size_t TBC = (size_t)BLOCK * VEC;
    pc.cpt = table;
    pc.scr = scr;
#pragma unroll
    for (int x = 0; x < RMAX; ++x)
#pragma unroll
        for (int v = 0; v < VEC; ++v) { pc.lacck[x][v] = T(0); pc.mk[x][v] = T(0); }
    // stage the outer parents' messages in scratch, the last parent's in registers
    int so = 0;
    if (kk > 1) {
#pragma unroll
        for (int j = 0; j < KNET - 1; ++j) {
            if (j < kk - 1) {
                const int rjj = ecard[j];
                pc.rj[j] = rjj;
                pc.soff[j] = so;
                const T* const mp = cur + (size_t)slot * TBC;
                for (int u = 0; u < rjj; ++u) {
                    const Pk<T, VEC> mm = ldp<T, VEC>(mp + u * TBC);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        scr[((so + u) * VEC + v) * BLOCK] = mm.v[v];
                        scr[((scr_half + so + u) * VEC + v) * BLOCK] = T(0);
                    }
                }
                so += rjj;
                slot += rjj;
            }
        }
    }
    pc.sacc_base = scr_half;
    const int rk = ecard[kk - 1];
    pc.rk = rk;
    {
        const T* const mp = cur + (size_t)slot * TBC;
#pragma unroll
        for (int u = 0; u < RMAX; ++u)
            if (u < rk) {
                const Pk<T, VEC> mm = ldp<T, VEC>(mp + u * TBC);
#pragma unroll
                for (int v = 0; v < VEC; ++v) pc.mk[u][v] = mm.v[v];
            }
    }
    if (kk == 1) parent_run<1, SRC, T, VEC, RMAX, MAXP>(pc);
    else if (kk == 2) parent_run<2, SRC, T, VEC, RMAX, MAXP>(pc);
    else if constexpr (KNET > 2) {
        if (kk == 3) parent_run<3, SRC, T, VEC, RMAX, MAXP>(pc);
        else if (kk == 4) parent_run<4, SRC, T, VEC, RMAX, MAXP>(pc);
        else if constexpr (KNET > 4) {
            if (kk == 5) parent_run<5, SRC, T, VEC, RMAX, MAXP>(pc);
            else if (kk == 6) parent_run<6, SRC, T, VEC, RMAX, MAXP>(pc);
            else if (kk == 7) parent_run<7, SRC, T, VEC, RMAX, MAXP>(pc);
            else parent_run<8, SRC, T, VEC, RMAX, MAXP>(pc);
        }
    }
    // lambda-messages to the outer parents (accumulated in scratch) ...
    int so2 = 0;
    for (int j = 0; j < kk - 1; ++j) {
        const int rjj = ecard[j];
        T val[RMAX][VEC];
        if constexpr (RMAX <= 8) {
#pragma unroll
            for (int u = 0; u < RMAX; ++u)
                if (u < rjj) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) val[u][v] = scr[((scr_half + so2 + u) * VEC + v) * BLOCK];
                }
        } else {
            for (int u = 0; u < rjj; ++u)
#pragma unroll
                for (int v = 0; v < VEC; ++v) val[u][v] = scr[((scr_half + so2 + u) * VEC + v) * BLOCK];
        }
        emit_msg(louts[j], val, rjj);
        so2 += rjj;
    }
    // ... and to the last parent (accumulated in registers)
    emit_msg(louts[kk - 1], pc.lacck, rk);
}

template <typename T> __device__ __forceinline__ T recip(T s) { return T(1) / s; }

// ------------------------------------------------------------------------------------------------
#ifndef BNBP_GEN_MINB
#define BNBP_GEN_MINB 1        // tuning knob (an alternative build, selected with BNBP_LIB): resident blocks per SM asked of ptxas
#endif
template <typename T, int VEC, int RMAX, int KNET, bool FREEZE, bool CHECK, bool MAXP = false>
__global__ void __launch_bounds__(BLOCK, BNBP_GEN_MINB)
sweep_kernel(const SweepArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr size_t TBC = (size_t)BLOCK * VEC;        // cases per tile = slot stride
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int chunk = blockIdx.y;
    const int lane0 = tid * VEC;                       // first case (within the tile) of this thread
    const size_t case0 = (size_t)tile * TBC + lane0;   // index into per-case arrays

    bool act[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) act[v] = true;
    if constexpr (FREEZE) {
        // device-side loop termination: once a sweep found no active case, later (speculatively
        // enqueued) launches return at once -- no host round-trip per sweep.
        if (a.sweep_index > 0 && *reinterpret_cast<volatile int32_t*>(a.last_active) < a.sweep_index - 1) return;
        // a case that met delta < eps at the previous (tested) sweep is frozen from now on: the
        // reference breaks right after the commit (:135-147), so its state is final.
        bool any = false;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            bool frozen = a.status[case0 + v] != 0;
            if (!frozen && a.prev_tested && a.delta_prev[case0 + v] < a.eps) {
                frozen = true;
                if (chunk == 0) { a.status[case0 + v] = 1; a.sweeps[case0 + v] = a.sweep_index; }
            }
            act[v] = !frozen;
            any |= act[v];
        }
        if (chunk == 0) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) a.delta_next[case0 + v] = Lim<T>::floor_();
        }
        if (!any) return;
        if (chunk == 0) {
            // one lane per warp that still has an active case publishes the sweep index
            const unsigned live = __activemask();
            if ((tid & 31) == __ffs(live) - 1) *a.last_active = a.sweep_index;   // benign race: same value
        }
    }

    T* const pl = a.pl + ((size_t)tile * a.PL) * TBC + lane0;
    const T* const cur = a.msg_cur + ((size_t)tile * a.M) * TBC + lane0;
    T* const nxt = a.msg_nxt + ((size_t)tile * a.M) * TBC + lane0;
    const uint32_t* const evb = a.evbits + ((size_t)tile * a.W) * TBC + lane0;
    T* const scr = reinterpret_cast<T*>(smem_raw) + tid;

    T dmax[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) dmax[v] = Lim<T>::floor_();

    // normalise a message (RMAX-padded register vector), damp / delta it against the time-t value
    // when CHECK, store it to the time-(t+1) buffer
    auto emit_msg = [&](int out, const T (&val)[RMAX][VEC], int rr) {
        T s[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] = T(0);
#pragma unroll
        for (int x = 0; x < RMAX; ++x)
            if (x < rr) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) s[v] += val[x][v];
            }
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] = recip(s[v]);
        T* dst = nxt + (size_t)out * TBC;
        const T* oldp = cur + (size_t)out * TBC;
#pragma unroll
        for (int x = 0; x < RMAX; ++x)
            if (x < rr) {
                Pk<T, VEC> o;
#pragma unroll
                for (int v = 0; v < VEC; ++v) o.v[v] = val[x][v] * s[v];
                if constexpr (CHECK) {
                    const Pk<T, VEC> old = ldp<T, VEC>(oldp + x * TBC);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        if (a.damping != T(0)) o.v[v] = (T(1) - a.damping) * o.v[v] + a.damping * old.v[v];
                        dmax[v] = absdiff_max(dmax[v], o.v[v], old.v[v]);
                    }
                }
                stp<T, VEC>(dst + x * TBC, o);
            }
    };
    // normalise pi_X / lambda_X and store in place; evidence nodes and frozen cases keep the old row
    auto emit_node = [&](T* dst, const T (&val)[RMAX][VEC], const T (&oldv)[RMAX][VEC], const bool (&upd)[VEC], int rr) {
        T s[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] = T(0);
#pragma unroll
        for (int x = 0; x < RMAX; ++x)
            if (x < rr) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) s[v] += val[x][v];
            }
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] = recip(s[v]);
#pragma unroll
        for (int x = 0; x < RMAX; ++x)
            if (x < rr) {
                Pk<T, VEC> o;
#pragma unroll
                for (int v = 0; v < VEC; ++v) o.v[v] = upd[v] ? val[x][v] * s[v] : oldv[x][v];
                stp<T, VEC>(dst + x * TBC, o);
            }
    };

    const int n0 = a.chunk_off[chunk], n1 = a.chunk_off[chunk + 1];
    Pk<uint32_t, VEC> evw;
    int evw_idx = -1;

    const int n_total = a.chunk_off[a.n_chunks];
    const int pin_total = a.nodes[0].lin_off;          // the lambda inboxes follow the pi inboxes in a message buffer

    for (int X = n0; X < n1; ++X) {
        const NodeMeta nd = a.nodes[X];
        const int r = nd.card, k = nd.k, m = nd.m;
        if (a.prefetch && X + 1 < n1) {
            // The walk is a dependent chain per node -- load the inputs, compute, store -- and at 12-16 resident warps per
            // SM (128-168 registers) the loads of one node do not cover the HBM latency (r02a: 8.6-10.2 long-scoreboard
            // stall cycles per issue, DRAM at 25-54 % of peak on dag2000 / grid100).  Asking L2 for the NEXT node's rows
            // now (its pi/lambda and its two inboxes) doubles the requests in flight without a single extra register.
            const NodeMeta nn = a.nodes[X + 1];
            const int pin_len = (X + 2 < n_total ? a.nodes[X + 2].pin_off : pin_total) - nn.pin_off;
            const T* p0 = pl + (size_t)nn.pl_off * TBC;
            for (int i = 0; i < 2 * nn.card; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + (size_t)i * TBC));
            p0 = cur + (size_t)nn.pin_off * TBC;
            for (int i = 0; i < pin_len; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + (size_t)i * TBC));
            p0 = cur + (size_t)nn.lin_off * TBC;
            for (int i = 0; i < nn.m * nn.card; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + (size_t)i * TBC));
        }
        if ((X >> 5) != evw_idx) { evw_idx = X >> 5; evw = ldp<uint32_t, VEC>(evb + (size_t)evw_idx * TBC); }
        bool upd[VEC];                                  // may pi_X / lambda_X be rewritten?
#pragma unroll
        for (int v = 0; v < VEC; ++v) upd[v] = act[v] && !((evw.v[v] >> (X & 31)) & 1u);

        // ---- time-t pi_X and lambda_X ------------------------------------------------------------
        T* const pX = pl + (size_t)nd.pl_off * TBC;
        T* const lX = pX + (size_t)r * TBC;
        T pi[RMAX][VEC];
        ParentCtx<T, VEC, RMAX, MAXP> pc;
#pragma unroll
        for (int x = 0; x < RMAX; ++x) {
            if (x < r) {
                const Pk<T, VEC> p = ldp<T, VEC>(pX + x * TBC);
                const Pk<T, VEC> l = ldp<T, VEC>(lX + x * TBC);
#pragma unroll
                for (int v = 0; v < VEC; ++v) { pi[x][v] = p.v[v]; pc.lam[x][v] = l.v[v]; }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { pi[x][v] = T(0); pc.lam[x][v] = T(0); }
            }
        }

        // ---- child side: lambda_X (:220-238) and pi-messages X->children (:202-218) -------------
        {
            const T* const Lb = cur + (size_t)nd.lin_off * TBC;
            const int32_t* const outs = a.c_pi_out + nd.c0;
            T ln[RMAX][VEC];
            if (m == 0) {
#pragma unroll
                for (int x = 0; x < RMAX; ++x)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) ln[x][v] = T(1);
            } else if (m == 1) {
#pragma unroll
                for (int x = 0; x < RMAX; ++x) {
                    if (x < r) {
                        const Pk<T, VEC> L = ldp<T, VEC>(Lb + x * TBC);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) ln[x][v] = L.v[v];
                    }
                }
                emit_msg(outs[0], pi, r);                       // no other child: N(pi_X)
            } else if (m == 2) {
                T L1[RMAX][VEC], pv[RMAX][VEC];
                const T* const Lb1 = Lb + (size_t)r * TBC;
#pragma unroll
                for (int x = 0; x < RMAX; ++x) {
                    if (x < r) {
                        const Pk<T, VEC> A = ldp<T, VEC>(Lb + x * TBC);
                        const Pk<T, VEC> B = ldp<T, VEC>(Lb1 + x * TBC);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) { ln[x][v] = A.v[v]; L1[x][v] = B.v[v]; }
                    } else {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) { ln[x][v] = T(0); L1[x][v] = T(0); }
                    }
                }
#pragma unroll
                for (int x = 0; x < RMAX; ++x)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) pv[x][v] = pi[x][v] * L1[x][v];
                emit_msg(outs[0], pv, r);
#pragma unroll
                for (int x = 0; x < RMAX; ++x)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) { pv[x][v] = pi[x][v] * ln[x][v]; ln[x][v] *= L1[x][v]; }
                emit_msg(outs[1], pv, r);
            } else {
#pragma unroll
                for (int x = 0; x < RMAX; ++x)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) ln[x][v] = T(1);
                for (int c = 0; c < m; ++c) {
                    const T* const Lc = Lb + (size_t)(c * r) * TBC;
#pragma unroll
                    for (int x = 0; x < RMAX; ++x) {
                        if (x < r) {
                            const Pk<T, VEC> L = ldp<T, VEC>(Lc + x * TBC);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) ln[x][v] *= L.v[v];
                        }
                    }
                }
                for (int c = 0; c < m; ++c) {
                    T pv[RMAX][VEC];
#pragma unroll
                    for (int x = 0; x < RMAX; ++x)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) pv[x][v] = pi[x][v];
                    for (int c2 = 0; c2 < m; ++c2) {
                        if (c2 == c) continue;
                        const T* const Lc = Lb + (size_t)(c2 * r) * TBC;
#pragma unroll
                        for (int x = 0; x < RMAX; ++x) {
                            if (x < r) {
                                const Pk<T, VEC> L = ldp<T, VEC>(Lc + x * TBC);
#pragma unroll
                                for (int v = 0; v < VEC; ++v) pv[x][v] *= L.v[v];
                            }
                        }
                    }
                    emit_msg(outs[c], pv, r);
                }
            }
            emit_node(lX, ln, pc.lam, upd, r);
        }

        // ---- parent side: pi_X (:174-200) and lambda-messages X->parents (:240-266) ---------------
        pc.r = r;
#pragma unroll
        for (int x = 0; x < RMAX; ++x)
#pragma unroll
            for (int v = 0; v < VEC; ++v) pc.pacc[x][v] = T(0);
        if (k == 0) {
            // root: all_combination_pattern calls the body once with the empty condition (:280-283)
            const T* const prior = a.cpt + nd.cpt_off;
#pragma unroll
            for (int x = 0; x < RMAX; ++x)
                if (x < r) {
                    const T p = __ldg(prior + x);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) pc.pacc[x][v] = p;
                }
        } else if (nd.dense_s == 0) {
            parent_group<SRC_CPT, KNET>(pc, a.cpt + nd.cpt_off, a.e_card + nd.e0, a.e_lam_out + nd.e0, cur, scr,
                                        nd.scr_half, k, nd.pin_off, emit_msg);
        } else {
            // dense node: the CPT already met the batch in two matrix products (bnbp_dense.cuh).
            // Group A = parents [0, s) against T2[uA]; group B = parents [s, k) against T1[(uB,x)].
            const int s = nd.dense_s;
            const T* const tsc = a.tscr + ((size_t)tile * a.TS) * TBC + lane0;
            const int32_t* const ecard = a.e_card + nd.e0;
            parent_group<SRC_T2, KNET>(pc, tsc + (size_t)nd.t2_off * TBC, ecard, a.e_lam_out + nd.e0, cur, scr,
                                       nd.scr_half, s, nd.pin_off, emit_msg);
            const T* const t1 = tsc + (size_t)nd.t1_off * TBC;
            if (s == k) {
                // no parent left in group B: T1[x] is the unnormalised pi_X
#pragma unroll
                for (int x = 0; x < RMAX; ++x)
                    if (x < r) {
                        const Pk<T, VEC> pk = ldp<T, VEC>(t1 + (size_t)x * TBC);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) pc.pacc[x][v] = pk.v[v];
                    }
            } else {
                int slot = nd.pin_off;
                for (int j = 0; j < s; ++j) slot += ecard[j];
                parent_group<SRC_T1, KNET>(pc, t1, ecard + s, a.e_lam_out + nd.e0 + s, cur, scr, nd.scr_half, k - s,
                                           slot, emit_msg);
            }
        }
        // pi_X = normalize(acc) unless X is evidence (:177) or the case is frozen
        emit_node(pX, pc.pacc, pi, upd, r);
    }

    if constexpr (CHECK) {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            if (act[v]) {
                if (a.n_chunks == 1) a.delta_cur[case0 + v] = dmax[v];
                else atomic_max_nonneg(&a.delta_cur[case0 + v], dmax[v]);
            }
    }
}

template <typename T, int VEC, int RMAX, int KNET>
cudaError_t launch_sweep_vr(const SweepArgs<T>& a, dim3 grid, size_t smem, bool freeze, bool check, cudaStream_t st, bool maxp)
{
    // max-product runs one flavour only: freeze + check (with eps <= 0 no case ever freezes: a fixed sweep count)
    if (maxp) sweep_kernel<T, VEC, RMAX, KNET, true, true, true><<<grid, BLOCK, smem, st>>>(a);
    else if (check) sweep_kernel<T, VEC, RMAX, KNET, true, true><<<grid, BLOCK, smem, st>>>(a);
    else if (freeze) sweep_kernel<T, VEC, RMAX, KNET, true, false><<<grid, BLOCK, smem, st>>>(a);
    else sweep_kernel<T, VEC, RMAX, KNET, false, false><<<grid, BLOCK, smem, st>>>(a);
    return cudaGetLastError();
}

template <typename T, int VEC, int RMAX, int KNET> cudaError_t set_sweep_smem(int bytes)
{
    cudaError_t e;
    e = cudaFuncSetAttribute(sweep_kernel<T, VEC, RMAX, KNET, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(sweep_kernel<T, VEC, RMAX, KNET, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(sweep_kernel<T, VEC, RMAX, KNET, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(sweep_kernel<T, VEC, RMAX, KNET, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

} // namespace bnbp
