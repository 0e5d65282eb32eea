// bnbp_sweep.cuh — the sweep kernel (one launch = one iteration of the reference's while(true),
// belief_propagation.hpp:75-148, for every resident case) and its launchers.
#pragma once
#include "bnbp_kernels.cuh"

namespace bnbp {

// ------------------------------------------------------------------------------------------------
// Parent side of one node: pi_X and all lambda-messages X->U_j from ONE pass over the CPT.
//   w(u)      = sum_x lambda_X(x) P(x|u)
//   pi_X(x)   = sum_u P(x|u) prod_j m_j(u_j)
//   lmsg_j(a) = sum_{u:u_j=a} w(u) prod_{i!=j} m_i(u_i)
// evaluated by a depth-K recursion over the parents (compile-time depth, runtime cardinalities):
// each level passes down the prefix product P and returns the message-weighted sum of its
// subtree R, so the leave-one-out product for level j is simply P*R -- (2k+2)|CPT| flops instead
// of the reference's (k^2+k+1)|CPT|.  The last parent and the states of X are unrolled to RMAX
// and live in registers; the outer parents' messages / accumulators sit in a per-thread shared
// memory scratch (dynamic index, conflict-free [value][thread] layout).
template <typename T, int VEC, int RMAX> struct ParentCtx {
    const T* cpt;
    int r, rk;
    int rj[KMAX];
    int soff[KMAX];
    T* scr;            // this thread's scratch column: scr[(i*VEC+v)*BD]
    int BD;
    int sacc_base;     // value index where the outer accumulators start
    T lam[RMAX][VEC], pacc[RMAX][VEC], mk[RMAX][VEC], lacck[RMAX][VEC];
};

template <typename T, int VEC, int RMAX>
__device__ __forceinline__ void parent_leaf(ParentCtx<T, VEC, RMAX>& c, const T (&P)[VEC], int q, T (&ret)[VEC])
{
    const T* blk = c.cpt + (size_t)q * (size_t)(c.rk * c.r);
#pragma unroll
    for (int v = 0; v < VEC; ++v) ret[v] = T(0);
    auto body = [&](int b) {
        T w[VEC], pm[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { w[v] = T(0); pm[v] = P[v] * c.mk[b][v]; }
        auto inner = [&](int x) {
            const T p = __ldg(blk + b * c.r + x);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                w[v] = fma(c.lam[x][v], p, w[v]);
                c.pacc[x][v] = fma(p, pm[v], c.pacc[x][v]);
            }
        };
        if constexpr (RMAX <= 8) {
#pragma unroll
            for (int x = 0; x < RMAX; ++x)
                if (x < c.r) inner(x);
        } else {
            for (int x = 0; x < c.r; ++x) inner(x);     // large cardinalities: arrays live in local memory
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            c.lacck[b][v] = fma(P[v], w[v], c.lacck[b][v]);
            ret[v] = fma(c.mk[b][v], w[v], ret[v]);
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

template <int LEVEL, int K, typename T, int VEC, int RMAX>
__device__ __forceinline__ void parent_rec(ParentCtx<T, VEC, RMAX>& c, const T (&P)[VEC], int q, T (&ret)[VEC])
{
    if constexpr (LEVEL == K - 1) {
        parent_leaf<T, VEC, RMAX>(c, P, q, ret);
    } else {
        const int rl = c.rj[LEVEL];
#pragma unroll
        for (int v = 0; v < VEC; ++v) ret[v] = T(0);
        for (int a = 0; a < rl; ++a) {
            T mv[VEC], P2[VEC], R[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                mv[v] = c.scr[((c.soff[LEVEL] + a) * VEC + v) * c.BD];
                P2[v] = P[v] * mv[v];
            }
            parent_rec<LEVEL + 1, K, T, VEC, RMAX>(c, P2, q * rl + a, R);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                T* acc = &c.scr[((c.sacc_base + c.soff[LEVEL] + a) * VEC + v) * c.BD];
                *acc = fma(P[v], R[v], *acc);
                ret[v] = fma(mv[v], R[v], ret[v]);
            }
        }
    }
}

template <int K, typename T, int VEC, int RMAX>
__device__ __forceinline__ void parent_run(ParentCtx<T, VEC, RMAX>& c)
{
    T one[VEC], ret[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) one[v] = T(1);
    parent_rec<0, K, T, VEC, RMAX>(c, one, 0, ret);
}

// ------------------------------------------------------------------------------------------------
template <typename T, int VEC, int RMAX, bool FREEZE, bool CHECK>
__global__ void __launch_bounds__(128)
sweep_kernel(const SweepArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int BD = blockDim.x;
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int chunk = blockIdx.y;
    const int lane0 = tid * VEC;                       // first case (within the tile) of this thread
    const size_t case0 = (size_t)tile * a.TB + lane0;  // index into per-case arrays

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

    T* pl = a.pl + ((size_t)tile * a.PL) * a.TB + lane0;
    const T* cur = a.msg_cur + ((size_t)tile * a.M) * a.TB + lane0;
    T* nxt = a.msg_nxt + ((size_t)tile * a.M) * a.TB + lane0;
    const uint32_t* evb = a.evbits + ((size_t)tile * a.W) * a.TB + lane0;
    const size_t TB = (size_t)a.TB;
    T* scr = smem + tid;

    T dmax[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) dmax[v] = Lim<T>::floor_();

    const int n0 = a.chunk_off[chunk], n1 = a.chunk_off[chunk + 1];
    Pk<uint32_t, VEC> evw;
    int evw_idx = -1;

    for (int X = n0; X < n1; ++X) {
        const NodeMeta nd = a.nodes[X];
        const int r = nd.card, k = nd.k, m = nd.m;
        if ((X >> 5) != evw_idx) { evw_idx = X >> 5; evw = ldp<uint32_t, VEC>(evb + (size_t)evw_idx * TB); }
        bool upd[VEC];                                  // may pi_X / lambda_X be rewritten?
#pragma unroll
        for (int v = 0; v < VEC; ++v) upd[v] = act[v] && !((evw.v[v] >> (X & 31)) & 1u);

        // ---- time-t pi_X and lambda_X ------------------------------------------------------------
        T pi[RMAX][VEC];
        ParentCtx<T, VEC, RMAX> pc;
#pragma unroll
        for (int x = 0; x < RMAX; ++x) {
            if (x < r) {
                Pk<T, VEC> p = ldp<T, VEC>(pl + (size_t)(nd.pl_off + x) * TB);
                Pk<T, VEC> l = ldp<T, VEC>(pl + (size_t)(nd.pl_off + r + x) * TB);
#pragma unroll
                for (int v = 0; v < VEC; ++v) { pi[x][v] = p.v[v]; pc.lam[x][v] = l.v[v]; }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { pi[x][v] = T(0); pc.lam[x][v] = T(0); }
            }
        }

        // ---- child side: lambda_X (:220-238) and pi-messages X->children (:202-218) -------------
        {
            T ln[RMAX][VEC];
#pragma unroll
            for (int x = 0; x < RMAX; ++x)
#pragma unroll
                for (int v = 0; v < VEC; ++v) ln[x][v] = T(1);
            for (int c = 0; c < m; ++c) {
#pragma unroll
                for (int x = 0; x < RMAX; ++x) {
                    if (x < r) {
                        Pk<T, VEC> L = ldp<T, VEC>(cur + (size_t)(nd.lin_off + c * r + x) * TB);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) ln[x][v] *= L.v[v];
                    }
                }
            }
            T s[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) s[v] = T(0);
#pragma unroll
            for (int x = 0; x < RMAX; ++x)
                if (x < r) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) s[v] += ln[x][v];
                }
#pragma unroll
            for (int x = 0; x < RMAX; ++x) {
                if (x < r) {
                    T* dst = pl + (size_t)(nd.pl_off + r + x) * TB;
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        if (upd[v]) dst[v] = ln[x][v] / s[v];
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
#pragma unroll
                    for (int x = 0; x < RMAX; ++x) {
                        if (x < r) {
                            Pk<T, VEC> L = ldp<T, VEC>(cur + (size_t)(nd.lin_off + c2 * r + x) * TB);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) pv[x][v] *= L.v[v];
                        }
                    }
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) s[v] = T(0);
#pragma unroll
                for (int x = 0; x < RMAX; ++x)
                    if (x < r) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) s[v] += pv[x][v];
                    }
                const int out = a.c_pi_out[nd.c0 + c];
#pragma unroll
                for (int x = 0; x < RMAX; ++x) {
                    if (x < r) {
                        Pk<T, VEC> o;
#pragma unroll
                        for (int v = 0; v < VEC; ++v) o.v[v] = pv[x][v] / s[v];
                        if constexpr (CHECK) {
                            Pk<T, VEC> old = ldp<T, VEC>(cur + (size_t)(out + x) * TB);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) {
                                if (a.damping != T(0)) o.v[v] = (T(1) - a.damping) * o.v[v] + a.damping * old.v[v];
                                dmax[v] = absdiff_max(dmax[v], o.v[v], old.v[v]);
                            }
                        }
                        stp<T, VEC>(nxt + (size_t)(out + x) * TB, o);
                    }
                }
            }
        }

        // ---- parent side: pi_X (:174-200) and lambda-messages X->parents (:240-266) ---------------
        pc.cpt = a.cpt + nd.cpt_off;
        pc.r = r;
        pc.scr = scr;
        pc.BD = BD;
#pragma unroll
        for (int x = 0; x < RMAX; ++x)
#pragma unroll
            for (int v = 0; v < VEC; ++v) { pc.pacc[x][v] = T(0); pc.lacck[x][v] = T(0); pc.mk[x][v] = T(0); }
        if (k == 0) {
            // root: all_combination_pattern calls the body once with the empty condition (:280-283)
#pragma unroll
            for (int x = 0; x < RMAX; ++x)
                if (x < r) {
                    const T p = __ldg(pc.cpt + x);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) pc.pacc[x][v] = p;
                }
        } else {
            // stage the outer parents' messages in scratch, the last parent's in registers
            int so = 0, slot = nd.pin_off;
#pragma unroll
            for (int j = 0; j < KMAX; ++j) {
                if (j < k - 1) {
                    const int rjj = a.e_card[nd.e0 + j];
                    pc.rj[j] = rjj;
                    pc.soff[j] = so;
                    for (int u = 0; u < rjj; ++u) {
                        Pk<T, VEC> mm = ldp<T, VEC>(cur + (size_t)(slot + u) * TB);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) scr[((so + u) * VEC + v) * BD] = mm.v[v];
                    }
                    so += rjj;
                    slot += rjj;
                }
            }
            pc.sacc_base = so;
            for (int i = 0; i < so; ++i)
#pragma unroll
                for (int v = 0; v < VEC; ++v) scr[((so + i) * VEC + v) * BD] = T(0);
            const int rk = a.e_card[nd.e0 + k - 1];
            pc.rk = rk;
#pragma unroll
            for (int u = 0; u < RMAX; ++u)
                if (u < rk) {
                    Pk<T, VEC> mm = ldp<T, VEC>(cur + (size_t)(slot + u) * TB);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) pc.mk[u][v] = mm.v[v];
                }
            switch (k) {
            case 1: parent_run<1, T, VEC, RMAX>(pc); break;
            case 2: parent_run<2, T, VEC, RMAX>(pc); break;
            case 3: parent_run<3, T, VEC, RMAX>(pc); break;
            case 4: parent_run<4, T, VEC, RMAX>(pc); break;
            case 5: parent_run<5, T, VEC, RMAX>(pc); break;
            case 6: parent_run<6, T, VEC, RMAX>(pc); break;
            case 7: parent_run<7, T, VEC, RMAX>(pc); break;
            default: parent_run<8, T, VEC, RMAX>(pc); break;
            }
            // lambda-messages to the parents: normalise, (damp, delta), store
            int so2 = 0;
            for (int j = 0; j < k; ++j) {
                const int rjj = a.e_card[nd.e0 + j];
                const int out = a.e_lam_out[nd.e0 + j];
                T s[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) s[v] = T(0);
                if (j == k - 1) {
#pragma unroll
                    for (int u = 0; u < RMAX; ++u)
                        if (u < rk) {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) s[v] += pc.lacck[u][v];
                        }
                } else {
                    for (int u = 0; u < rjj; ++u)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) s[v] += scr[((pc.sacc_base + so2 + u) * VEC + v) * BD];
                }
                auto emit = [&](int u, const T (&val)[VEC]) {
                    Pk<T, VEC> o;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) o.v[v] = val[v] / s[v];
                    if constexpr (CHECK) {
                        Pk<T, VEC> old = ldp<T, VEC>(cur + (size_t)(out + u) * TB);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            if (a.damping != T(0)) o.v[v] = (T(1) - a.damping) * o.v[v] + a.damping * old.v[v];
                            dmax[v] = absdiff_max(dmax[v], o.v[v], old.v[v]);
                        }
                    }
                    stp<T, VEC>(nxt + (size_t)(out + u) * TB, o);
                };
                if (j == k - 1) {
#pragma unroll
                    for (int u = 0; u < RMAX; ++u)
                        if (u < rk) emit(u, pc.lacck[u]);
                } else {
                    for (int u = 0; u < rjj; ++u) {
                        T val[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) val[v] = scr[((pc.sacc_base + so2 + u) * VEC + v) * BD];
                        emit(u, val);
                    }
                    so2 += rjj;
                }
            }
        }
        // pi_X = normalize(acc) unless X is evidence (:177) or the case is frozen
        {
            T s[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) s[v] = T(0);
#pragma unroll
            for (int x = 0; x < RMAX; ++x)
                if (x < r) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) s[v] += pc.pacc[x][v];
                }
#pragma unroll
            for (int x = 0; x < RMAX; ++x)
                if (x < r) {
                    T* dst = pl + (size_t)(nd.pl_off + x) * TB;
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        if (upd[v]) dst[v] = pc.pacc[x][v] / s[v];
                }
        }
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

template <typename T, int VEC, int RMAX>
cudaError_t launch_sweep_vr(const SweepArgs<T>& a, dim3 grid, int block, size_t smem, bool freeze, bool check,
                            cudaStream_t st)
{
    if (check) sweep_kernel<T, VEC, RMAX, true, true><<<grid, block, smem, st>>>(a);
    else if (freeze) sweep_kernel<T, VEC, RMAX, true, false><<<grid, block, smem, st>>>(a);
    else sweep_kernel<T, VEC, RMAX, false, false><<<grid, block, smem, st>>>(a);
    return cudaGetLastError();
}

template <typename T, int VEC, int RMAX> cudaError_t set_sweep_smem(int bytes)
{
    cudaError_t e;
    e = cudaFuncSetAttribute(sweep_kernel<T, VEC, RMAX, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(sweep_kernel<T, VEC, RMAX, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(sweep_kernel<T, VEC, RMAX, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

} // namespace bnbp
