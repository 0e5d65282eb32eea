// bnbp_onchip.cuh — the ON-CHIP multi-sweep kernel (sm_100a), compiled at run time behind bnbp_spec.cuh.
//
// The streaming kernel of bnbp_spec.cuh moves the whole state of a case (S = PL + M values: alarm37 442) through
// HBM on every sweep and sits on the HBM roofline; the arithmetic units idle (fp64 pipe 18 %, issue 16 %,
// profiles/r01c).  A case's state is small (alarm37 fp64: 442 values = 3.5 KB), so here it never leaves the SM
// between the first and the last sweep of the case (belief_propagation.hpp:31-159 for one case = init :33-73,
// `sweeps` iterations of :75-148, beliefs :151-158, all inside ONE kernel):
//
//   * a GROUP of 32 cases lives in the shared memory of one CTA, slot-major `state[slot][32]` (a warp-wide
//     access to one slot is 32 consecutive values: conflict-free): pi/lambda and ONE message buffer;
//   * the group is walked by BNBP_ROLES warps: lane = case, warp = a fixed subset of the nodes (longest-
//     processing-time partition of the per-node cost, made by the network compiler).  Every warp runs the
//     fully unrolled straight-line code of ITS nodes (the node arithmetic of bnbp_spec.cuh, CPT entries as
//     constant-bank operands), so the schedulers of the SM work on different nodes of the same 32 cases;
//   * a sweep has two phases.  A: every warp reads the inbox of its nodes (the time-t messages addressed to them)
//     into registers; barrier; B: it computes its nodes and overwrites the inboxes of their neighbours with the
//     time-(t+1) messages; barrier.  All four updates read time-t state only (:75-101), so this IS the Jacobi
//     commit (:135-143) without a second message buffer -- which is what lets TWO groups share an SM in fp64
//     (113 KB each): the warps of the two groups that own the same nodes run the same instructions, and the kernel
//     is bound by instruction fetch (r02a: one warp per scheduler walking 29 KB of straight-line code of its own
//     stalls 6 cycles per issue on `no_instruction`, L1.5 I-cache = 32 KB);
//   * pi_X / lambda_X are kept unnormalised and the reciprocals skip the slow path (bnbp_spec.cuh: emit_node,
//     rcp_norm): 40 % fewer instructions per sweep;
//   * the convergence test (:105-131,:147) is free: each warp reduces |new - old| over the messages it emits in
//     registers (the old value is the word it is about to overwrite), the per-case maximum over the warps goes
//     through ROLES x 32 shared-memory words at the sweep barrier;
//   * a lane whose case has stopped (delta < epsilon, or the sweep cap) is frozen; once 8 lanes of the group wait
//     (or nothing else runs) they write their marginals, take the next cases from a global ticket counter, load
//     their evidence and start over at sweep 0 while the other lanes carry on: in epsilon mode no lane waits for
//     the slowest case of its group (the kernel is persistent: grid = resident CTAs, not cases / 32);
//   * HBM sees the evidence of a case once (CSR entries) and its marginals once: ~0.9 KB per case instead of
//     2 * S * sizeof(T) * sweeps = 141 KB for 20 fp64 sweeps of alarm37.
//
// Generator macros in addition to those of bnbp_spec.cuh: BNBP_ROLES, and per role r
//   BNBP_WALK_r   BNBP_DECL(all) BNBP_LOADM(all) BNBP_SYNC BNBP_COMP(all): the two phases of the role's nodes
//   BNBP_NODES_r(OP)   OP(Ni) for every node of the role
// plus the constant tables bnbp_owner[N] (role that owns a node), bnbp_ploff[N], bnbp_cardn[N].
#if BNBP_VARIANT >= 8

namespace bnbp_spec {

struct OcArgs {                         // mirrors OnchipArgs in bnbp_jit.h
    const long long* ev_off;            // entries of case c: [ev_off[c] - ev_base, ev_off[c+1] - ev_base)
    long long ev_base;
    const int* ev_node;
    const int* ev_state;
    long long n_cases;
    OUT* out;                           // [n_cases][out_stride] case-major marginals
    const int* bel_col;                 // [N] column of node x's marginal in an output row, -1: not asked for
    long long out_stride;               // values per output row (sum r_X, or of the queried nodes only)
    int* out_sweeps;                    // [n_cases] or null
    unsigned char* out_conv;            // [n_cases] or null
    unsigned long long* ticket;         // next case to hand out (zero at launch)
    int* error_flag;
    T eps;                              // <= 0: fixed sweep count
    T damping;
    int max_sweeps;
    int interval;                       // convergence tested every interval-th sweep (1 = reference)
    int first_reserved_sm;              // CTAs that land on an SM with %smid >= this leave at once (>= #SMs: none)
};

constexpr int ROLES = BNBP_ROLES;
constexpr int OC_THREADS = 32 * ROLES;
constexpr int RETIRE_BATCH = 8;         // frozen lanes a group collects before it writes marginals and refills
// shared memory: [PL][32] pi/lambda | [M][32] messages | partial deltas [ROLES][32] | tickets [32]

__device__ __forceinline__ void oc_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(OC_THREADS) : "memory"); }

// time-0 state of one node (:33-65): pi = lambda = 1, a root's pi = its raw prior row (:58-64); every message the
// node is about to READ (its inbox in the current buffer) = 1 (:44-55)
template <class N> __device__ __forceinline__ void oc_init_node(T* pl, T* cur, const bool need)
{
    if (!need) return;
#pragma unroll
    for (int x = 0; x < N::R; ++x) {
        T prior = T(1);
        if constexpr (N::K == 0) prior = bnbp_cpt[N::CPT + x];
        pl[(N::PL + x) * 32] = prior;
        pl[(N::PL + N::R + x) * 32] = T(1);
    }
    constexpr int n_pin = pin_row<N, N::K>();
#pragma unroll
    for (int i = 0; i < n_pin; ++i) cur[(N::PIN + i) * 32] = T(1);
#pragma unroll
    for (int i = 0; i < N::M * N::R; ++i) cur[(N::LIN + i) * 32] = T(1);
}

// BEL = normalize(pi .* lambda) (:151-158, matrix.hpp:73-93), same explicit fma / rounded product as K4
template <class N> __device__ __forceinline__ void oc_belief_node(const T* pl, OUT* row_all, const int* __restrict__ bel_col, bool emit)
{
    const int col = bel_col[N::X];
    emit = emit && col >= 0;
    OUT* const row = row_all + (col >= 0 ? col : 0) - N::BEL;
    T p[N::R], l[N::R];
#pragma unroll
    for (int x = 0; x < N::R; ++x) { p[x] = pl[(N::PL + x) * 32]; l[x] = pl[(N::PL + N::R + x) * 32]; }
    T s = T(0);
#pragma unroll
    for (int x = 0; x < N::R; ++x) s = fma(p[x], l[x], s);
    const T inv = rcp_norm(s);
    if (emit) {
#pragma unroll
        for (int x = 0; x < N::R; ++x) row[N::BEL + x] = (OUT)(mul_rn(p[x], l[x]) * inv);
    }
}

} // namespace bnbp_spec

extern "C" __global__ void __launch_bounds__(bnbp_spec::OC_THREADS, BNBP_MINB)
bnbp_onchip_run(const bnbp_spec::OcArgs a)
{
    using namespace bnbp_spec;
    extern __shared__ __align__(16) unsigned char oc_smem[];
    T* const s_pl = reinterpret_cast<T*>(oc_smem);
    T* const s_msg = s_pl + (long long)BNBP_PL * 32;
    T* const s_delta = s_msg + (long long)BNBP_M * 32;                        // [ROLES][32]
    long long* const s_case = reinterpret_cast<long long*>(s_delta + ROLES * 32);       // [32]

    // Compute / communication split of the SMs (gather runs).  This grid is persistent and fills every SM (two
    // groups = 226 KB of shared memory, 64 K registers), so a concurrent NCCL kernel -- the exchange of the previous chunk's
    // marginals -- finds no SM to run on and the "overlap" degenerates into a queue (r02f: +1.27 ms per step at N = 2, the
    // full serial cost).  The CTAs that land on the last few SMs therefore leave at once; the ticket counter hands their
    // share of the cases to the others, and the collective has those SMs to itself.
    {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        if ((int)smid >= a.first_reserved_sm) return;
    }
    const int lane = threadIdx.x & 31;
    const int role = threadIdx.x >> 5;
    const bool eps_mode = a.eps > T(0);

    unsigned evl[BNBP_W];               // observed-node bits of this lane's case (kept apart from Ctx: the evidence loop
                                        // indexes it dynamically, and a Ctx on the stack would lose the address space of
                                        // its shared-memory pointers: generic LD / ST instead of LDS / STS)
    long long cid = -1;                 // the case this lane works on (-1: needs one, -2: none left)
    int sw = 0;                         // sweeps its case has run
    bool fin = false, fconv = false;    // the case has stopped (frozen until the group retires a batch) / it met delta < eps
#pragma unroll
    for (int w = 0; w < BNBP_W; ++w) evl[w] = 0u;

    for (;;) {
        // ---- hand out cases to the lanes that have none (every warp takes the same decisions) --------------
        const unsigned need_mask = __ballot_sync(0xffffffffu, cid == -1);
        if (need_mask) {
            if (role == 0) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.ticket, (unsigned long long)__popc(need_mask));
                base = __shfl_sync(0xffffffffu, base, 0);
                const long long mine = (long long)base + __popc(need_mask & ((1u << lane) - 1u));
                if (cid == -1) s_case[lane] = mine < a.n_cases ? mine : -2;
            }
            oc_barrier();
            const bool need = cid == -1;
            if (need) { cid = s_case[lane]; sw = 0; fin = false; fconv = false; }
            const bool fresh = need && cid >= 0;
            if (fresh) {
#pragma unroll
                for (int w = 0; w < BNBP_W; ++w) evl[w] = 0u;
            }
            switch (role) {
#define BNBP_OC_INIT(NN) oc_init_node<NN>(s_pl + lane, s_msg + lane, fresh);
#define BNBP_OC_CASE(R) case R: { BNBP_NODES_##R(BNBP_OC_INIT) } break;
                BNBP_ROLE_LIST(BNBP_OC_CASE)
#undef BNBP_OC_CASE
#undef BNBP_OC_INIT
            }
            if (fresh) {
                // evidence (:68-73): the row becomes pi AND lambda of the node and is never updated again (:177,:223).
                // Every warp records the observed set; the owner of a node writes its rows.
                const long long e0 = a.ev_off[cid] - a.ev_base, e1 = a.ev_off[cid + 1] - a.ev_base;
                for (long long e = e0; e < e1; ++e) {
                    const int node = a.ev_node[e];
                    if (node < 0 || node >= BNBP_N) { *a.error_flag = 1; continue; }
                    const int st = a.ev_state[e];
                    const int r = bnbp_cardn[node];
                    if (st < 0 || st >= r) { *a.error_flag = 3; continue; }
                    evl[node >> 5] |= 1u << (node & 31);
                    if (bnbp_owner[node] == role) {
                        T* const row = s_pl + lane + (long long)bnbp_ploff[node] * 32;
                        for (int x = 0; x < r; ++x) {
                            const T hot = x == st ? T(1) : T(0);
                            row[x * 32] = hot;
                            row[(r + x) * 32] = hot;
                        }
                    }
                }
            }
            if (__ballot_sync(0xffffffffu, cid >= 0) == 0u) break;      // no case left in this group
            // (the inbox words written above are read by this warp only; no barrier needed before phase A)
        }

        // ---- one sweep (:75-101) for the 32 cases of the group, this warp's nodes --------------------------
        Ctx c;
        c.damping = a.damping;
        c.evst = nullptr;
        c.tile = nullptr;
        c.pl = s_pl + lane;
        c.cur = s_msg + lane;           // one message buffer: inboxes are read before the sweep barrier, written after it
        c.nxt = s_msg + lane;
#pragma unroll
        for (int w = 0; w < BNBP_W; ++w) c.evw[w][0] = evl[w];
        c.act[0] = cid >= 0 && !fin;
        c.dmax[0] = Floor<T>::v();
        switch (role) {
#define BNBP_DECL(NN) In<NN> in_##NN;
#define BNBP_LOADM(NN) load_msgs<NN>(c, in_##NN);
#define BNBP_SYNC oc_barrier();
#define BNBP_COMP(NN) load_pl<NN>(c, in_##NN); compute_node<NN>(c, in_##NN);
#define BNBP_OC_CASE(R) case R: { BNBP_WALK_##R } break;
            BNBP_ROLE_LIST(BNBP_OC_CASE)
#undef BNBP_OC_CASE
#undef BNBP_DECL
#undef BNBP_LOADM
#undef BNBP_SYNC
#undef BNBP_COMP
        }
        // partial deltas: written after this sweep's phase-A barrier, read by every warp right after the commit barrier and
        // before ITS next phase-A barrier -- so the next sweep's writes cannot overtake a slow reader
        T* const dpart = s_delta;
        if constexpr (CHECK) dpart[role * 32 + lane] = c.dmax[0];
        oc_barrier();                                      // commit (:135-143): every time-(t+1) value is written
        const bool active = c.act[0];
        if (active) ++sw;

        // ---- stopping rule (:105-131,:147) per case -----------------------------------------------------------
        if constexpr (CHECK) {
            if (eps_mode) {
                const bool tested = (sw % a.interval) == 0 || sw >= a.max_sweeps;
                T d = Floor<T>::v();
#pragma unroll
                for (int r = 0; r < ROLES; ++r) d = fmax(d, dpart[r * 32 + lane]);
                if (active && tested && d < a.eps) { fin = true; fconv = true; }
            }
        }
        if (active && sw >= a.max_sweeps) fin = true;
        // a stopped case is frozen (the reference returns right after this commit, :147-158); its marginals are written
        // when the group has a batch of them, or nothing else to do
        const unsigned fin_mask = __ballot_sync(0xffffffffu, fin && cid >= 0);
        const unsigned run_mask = __ballot_sync(0xffffffffu, cid >= 0 && !fin);
        if (fin_mask && (__popc(fin_mask) >= RETIRE_BATCH || run_mask == 0u)) {
            const bool done = fin && cid >= 0;
            OUT* const row = a.out + (done ? cid : 0) * a.out_stride;
            switch (role) {
#define BNBP_OC_BEL(NN) oc_belief_node<NN>(s_pl + lane, row, a.bel_col, done);
#define BNBP_OC_CASE(R) case R: { BNBP_NODES_##R(BNBP_OC_BEL) } break;
                BNBP_ROLE_LIST(BNBP_OC_CASE)
#undef BNBP_OC_CASE
#undef BNBP_OC_BEL
            }
            if (role == 0 && done) {
                if (a.out_sweeps) a.out_sweeps[cid] = sw;
                if (a.out_conv) a.out_conv[cid] = fconv ? 1 : 0;
            }
            if (done) { cid = -1; fin = false; }
        }
    }
}

#endif // BNBP_VARIANT >= 8
