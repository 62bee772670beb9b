// tick_kernels.cu -- the tick pass of the split engine: reservoirs AND routing of the whole domain in
// one persistent kernel that walks the time axis in windows of 8 steps.
//
// Why: the reservoirs (md_forward_structure.f90:106-144) want [t][cell] (coalesced over cells), the routing
// (md_routing_operator.f90:17-79) needs every cell's inflow series.  Writing whole series as rows for a later routing pass
// costs 2-3x the algorithmic DRAM traffic and leaves the walk down the main rivers as a serial tail.  Here every cell advances
// 8 steps at a time and hands its 8 discharge values to its consumer as one 32-byte block XR[window][cell][8].
//
// Units (host: build_tick_topo in route_graph.cpp)
//   tile   32 consecutive cells in `path` order = 32 consecutive columns of the forcing: forcing box [8 steps][32 cells] by
//          2-D TMA, reservoirs of all 32 cells, then the routing of its shallow cells (classes S and R) in the reference's
//          summation order, strictly sequential in time
//   reach  up to 32 consecutive cells of a heavy-path chain of the deep cells (class D: flow accumulation above
//          `shallow_acc`), lane = cell.  At one time step the discharge of lane i is affine in the discharge of lane i - 1
//          (slope 1 - E_i, constant in time), so the 32 cells are routed by one warp scan per time step, and every lane then
//          re-evaluates the reference's statements with the upstream value it received.
// Ticket = (unit, window).  Every unit has a stage sigma (1 + the largest stage among its producers); tickets are processed
// in the order of their key sigma + window, so a river is a pipeline: while the tiles work on window w, the reach k stages
// further down works on window w - k.  A ticket only ever reads what tickets with a smaller key wrote.
// Every warp of the fully resident grid OWNS a fixed set of units (dealt round-robin by stage) and walks its tickets in key
// order: the warp that holds the smallest unfinished key never waits, hence no deadlock; a unit's carried states are read
// back by the thread that wrote them.  Hand-over: every block a ticket writes for another unit is counted at
// cnt[window][reading unit] once the warp has left the tick (one release fence per tick, then one reduction per block); a
// ticket starts when its counter shows every block it reads (one acquire load).  Stages that hold many units run two ticks
// after their producers, so in the bulk of the run nobody waits for a ticket that is still being worked on.
//
// Pit pairs (class P) only get their runoff here, as rows; route_pairs_kernel (split_kernels.cu) runs them afterwards.
//
// Reference statements are cited as file:line under /root/reference/smash/solver/.
#include "split_kernels.cuh"

#include <algorithm>

#include "cell_math.cuh"

namespace smash {

namespace {

constexpr unsigned FULLM = 0xffffffffu;
constexpr int TK_WARPS = 4;    // warps per CTA; every warp runs its own pipeline, no CTA barrier
constexpr int TK_SLOTS = 2;    // forcing boxes in flight per warp = tile tickets requested ahead
constexpr int TK_MAXU = 24;    // units a warp can own

typedef float TkSlot[2][TK_W][32];   // prcp, pet: [step][lane]

__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// 2-D TMA tile load global -> shared with an L2 eviction hint (the forcing is read exactly once)
__device__ __forceinline__ void tma_load_2d_hint(void *dst, const CUtensorMap *tm, int x, int y, uint64_t *bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
// ---- L2 residency classes.  The per-cell records, carried states and arrival counters are re-read every window and must
// stay in L2 (evict_last); the forcing and the domain series stream through once (evict_first); an exchange block is kept
// with normal priority until its consumer has read it (the read demotes it), unless the consumer runs many ticks later.
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldg_keep_f4(const float4 *p, uint64_t pol) {
    float4 v;
    asm("ld.global.cg.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ldg_keep_f(const float *p, uint64_t pol) {            // carried states: ordered with their stores
    float v;
    asm volatile("ld.global.cg.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ float ldg_keep_cf(const float *p, uint64_t pol) {           // read-only
    float v;
    asm("ld.global.cg.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int ldg_keep_i(const int *p, uint64_t pol) {
    int v;
    asm("ld.global.cg.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));   // read-only data: the compiler may move it
    return v;
}
__device__ __forceinline__ int ldg_keep_vi(const int *p, uint64_t pol) {               // stays where it is written
    int v;
    asm volatile("ld.global.cg.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg_keep_f(float *p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
// exchange blocks: 32 bytes = one sector
__device__ __forceinline__ void ldx(const float *p, float *v) {            // read once: L1 no-allocate, L2 evict_first
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void stx_near(float *p, const float *v) {        // the consumer reads it within a few ticks
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void stx_far(float *p, const float *v) {         // the consumer runs many ticks later: leave L2 early
    asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2_block(const float *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void red_add(int *p, int v) { asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_release() { asm volatile("fence.release.gpu;" ::: "memory"); }
// lane 0 waits until `need` blocks of (unit, window) have arrived; a spin that does not end sets the error word instead of
// hanging the device
__device__ __forceinline__ void wait_arrivals(const int *cnt, int need, int unit, int *err, int lane) {
    if (lane == 0) {
        int spins = 0;
        while (ld_acquire(cnt) < need) {
            __nanosleep(64);
            if (++spins > (1 << 21)) { atomicExch(err, 1 + unit); break; }
        }
    }
    __syncwarp();
}

// gr-a cell-step for a window without forcing gap, tanh arguments below 0.25 and hp_imd <= 15 (1 + (hp_imd/1000)^4 == 1
// in float32): the statements of vertical_step_nogap (cell_math.cuh) with those warp-uniform branches resolved once per
// window.  Same code as vertical_step_lean of split_kernels.cu: bit-identical results.
template <bool EXC>
__device__ __forceinline__ float tick_step_lean(const CellConst &k, float prcp, float pet, float &hp, float &hft) {
    const float ei = fminf(pet, prcp);                                   // md_forward_structure.f90:112
    const float pn = fmaxf(0.0f, prcp - ei);                             // :114
    const float en = pet - ei;                                           // :116
    const bool wet = pn > 0.0f;
    const float x = (wet ? pn : en) * k.inv_cp;
    const float x2 = x * x;
    float p = fmaf(x2, 0.021869488f, -0.053968254f);
    p = fmaf(x2, p, 0.13333334f);
    p = fmaf(x2, p, -0.33333334f);
    const float th = fmaf(x * x2, p, x);
    const float num = (wet ? k.cp * (1.0f - hp * hp) : (hp * k.cp) * (2.0f - hp)) * th;     // md_gr_operator.f90:52,55
    const float den = fmaf(wet ? hp : 1.0f - hp, th, 1.0f);
    const float r = num * mufu_rcp(den);
    const float hp_imd = hp + (wet ? r : -r) * k.inv_cp;                 // :58
    const float pr = wet ? pn - (hp_imd - hp) * k.cp : 0.0f;             // :60-62
    hp = hp_imd;                                                         // perc == 0 (:66-68)
    const float l = EXC ? k.exc * ((hft * hft) * hft * fsqrt_fast(hft)) : 0.0f;             // :77
    const float prr = fmaf(0.9f, pr, l);                                 // md_forward_structure.f90:137
    const float prd = 0.1f * pr;                                         // :138
    const float u = fmaxf(1.e-6f, fmaf(prr, k.inv_cft, hft));            // md_gr_operator.f90:102
    const float z = pow4(u);
    const float s2 = fsqrt_fast(1.0f + z), s1 = fsqrt_fast(s2);
    const float g = z * mufu_rcp(s1 * (s1 + 1.0f) * (s2 + 1.0f));        // 1 - (1+u^4)^(-1/4), cancellation-free (:104)
    const float rel = u * g;
    hft = u - rel;
    return fmaf(rel, k.cft, fmaxf(0.0f, prd + l));                       // qt = qr + qd (:106, md_forward_structure.f90:142-144)
}

__global__ void __launch_bounds__(256) tick_prep_kernel(const TkArgs a, const float *fields) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int npad = a.tp.npad;
    if (j >= npad) return;
    float cp = 200.0f, cft = 500.0f, exc = 0.0f, lr = 5.0f, hp = 0.01f, hft = 0.01f, hlr = 0.0f;
    if (j < a.tp.n) {
        cp = fields[(size_t)F_CP * npad + j]; cft = fields[(size_t)F_CFT * npad + j]; exc = fields[(size_t)F_EXC * npad + j];
        lr = fields[(size_t)F_LR * npad + j];
        hp = fields[(size_t)F_HP * npad + j]; hft = fields[(size_t)F_HFT * npad + j]; hlr = fields[(size_t)F_HLR * npad + j];
    }
    a.cc[j] = make_float4(cp, cft, exc, expf(-a.dt / (lr * 60.0f)));     // md_routing_operator.f90:75
    a.fstates[j] = hp; a.fstates[(size_t)npad + j] = hft;
    // the routing state of the pit cells is carried by route_pairs_kernel (hcar)
    if ((a.tp.meta[j] & 3) != 2) a.fstates[(size_t)2 * npad + j] = hlr;
}

template <int MINB>   // CTAs per SM the register allocation aims at (8: 64 registers, 6: 80, 4: no cap)
__global__ void __launch_bounds__(TK_WARPS * 32, MINB) tick_forward_kernel(const __grid_constant__ CUtensorMap tm_prcp,
                                                                           const __grid_constant__ CUtensorMap tm_pet, const TkArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TkSlot *slots = reinterpret_cast<TkSlot *>(smem_raw) + warp * TK_SLOTS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + sizeof(TkSlot) * TK_WARPS * TK_SLOTS) + warp * TK_SLOTS;
    int4 *mine = reinterpret_cast<int4 *>(smem_raw + (sizeof(TkSlot) + sizeof(uint64_t)) * TK_WARPS * TK_SLOTS) + warp * TK_MAXU;
    const TkTopo &tp = a.tp;
    const int ntile = tp.ntile, n = tp.n, npad = tp.npad, T = a.T, ng = tp.ng;
    const int nunit = ntile + tp.nreach;
    const int gw = (int)blockIdx.x * TK_WARPS + warp;
    const uint64_t pol = policy_evict_first(), keep = policy_evict_last();
    const float c0 = a.dx * a.dx * 0.001f / a.dt;                         // md_forward_structure.f90:155
    const float d0 = 0.001f * a.dx * a.dx;                                // md_routing_operator.f90:56, times (flwacc - 1) per cell
    constexpr uint32_t SLOT_BYTES = sizeof(TkSlot);
    const int nwin = a.nwin;

    // ---- the units this warp owns, sorted by stage: (unit, key of its stage, blocks it waits for per window)
    int nu = 0;
    {
        const int2 *src = reinterpret_cast<const int2 *>(a.wunits) + (size_t)gw * a.maxu;
        int2 v = make_int2(-1, 0);
        if (lane < a.maxu) v = src[lane];
        if (lane < TK_MAXU) mine[lane] = make_int4(v.x, v.y, v.x >= 0 ? tp.need[v.x] : 0, 0);
        nu = __popc(__ballot_sync(FULLM, v.x >= 0));
    }
    __syncwarp();
    if (nu == 0) return;
    const int kfirst = mine[0].y;

    // ---- the forcing box requested next, in the order the boxes are used: (tick, unit index, window of the visit)
    const int nb = a.nb, nvis = (nwin + nb - 1) / nb;                     // windows per visit, visits per unit
    const int kmax = mine[nu - 1].y + nvis - 1;
    auto next_box = [&](int &kk, int &ii, int &bb) {                      // kk > kmax: none left
        for (;;) {
            const int vv = kk - (ii >= 0 ? mine[ii].y : 0);
            if (ii >= 0 && kk <= kmax && vv >= 0 && vv < nvis && mine[ii].x < ntile && bb + 1 < nb && vv * nb + bb + 1 < nwin) { bb++; return; }
            bb = -1;
            if (++ii == nu) { ii = 0; kk++; }
            if (kk > kmax) return;
        }
    };
    auto request = [&](int s, int kk, int ii, int bb) {
        if (lane == 0) {
            const int ww = (kk - mine[ii].y) * nb + bb;
            mbar_expect_tx(&bars[s], SLOT_BYTES);
            tma_load_2d_hint(&slots[s][0][0][0], &tm_prcp, mine[ii].x * 32, ww * TK_W, &bars[s], pol);
            tma_load_2d_hint(&slots[s][1][0][0], &tm_pet, mine[ii].x * 32, ww * TK_W, &bars[s], pol);
        }
    };
    int kq = kfirst, iq = -1, bq = -1;
    next_box(kq, iq, bq);
    if (lane == 0) {
        for (int s = 0; s < TK_SLOTS; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    for (int s = 0; s < TK_SLOTS; s++) {
        if (kq <= kmax) {
            request(s, kq, iq, bq);
            next_box(kq, iq, bq);
        }
    }

    float *fs = a.fstates;
    float *qsim = a.qsim;
    const size_t qpitch = (size_t)a.qpitch;
    const size_t xstride = (size_t)npad * TK_W;                           // one window of exchange blocks
    uint32_t parity = 0;
    int slot = 0;
#pragma unroll 1
    for (int k = kfirst; k <= kmax; k++) {
        bool any = false;
#pragma unroll 1
        for (int i = 0; i < nu; i++) {
            const int4 me = mine[i];
            const int unit = me.x, v = k - me.y;                           // visit v covers windows [v nb, v nb + nbe)
            if (v < 0 || v >= nvis) continue;
            any = true;
            const int w0 = v * nb, nbe = min(nb, nwin - w0);
            if (unit < ntile) {
                // =============================================================== tile visit
                const int tile = unit;
                const int j = tile * 32 + lane;
                const bool valid = j < n;
                const float4 c4 = ldg_keep_f4(a.cc + j, keep);
                float hp = ldg_keep_f(fs + j, keep), hft = ldg_keep_f(fs + (size_t)npad + j, keep);
                const int meta = ldg_keep_i(tp.meta + j, keep);
                const int cls = meta & 3;
                const int cons = ldg_keep_i(tp.cons1 + j, keep);
                float hlr = (cls <= 1) ? ldg_keep_f(fs + (size_t)2 * npad + j, keep) : 0.0f;
                const int nup = (cls == 1) ? (meta >> 8 & 15) : 0;
                const int uo = (nup > 0) ? ldg_keep_i(tp.upoff + j, keep) : 0;
                if (me.z > 0) {
                    // blocks of other units: all of them have arrived when the counter of this visit is full; pull them towards L2
                    wait_arrivals(a.cnt + (size_t)v * nunit + unit, me.z, unit, a.err, lane);
                    for (int e = 0; e < nup; e++) {
                        const int src = ldg_keep_i(tp.ups + uo + e, keep) & ~(1 << 30);
                        for (int b = 0; b < nbe; b++) prefetch_l2_block(a.X + (size_t)(w0 + b) * xstride + (size_t)src * TK_W);
                    }
                }
                CellConst kc;
                kc.cp = c4.x; kc.inv_cp = __frcp_rn(c4.x);                 // md_gr_operator.f90:47
                kc.cft = c4.y; kc.inv_cft = __frcp_rn(c4.y); kc.cft_m4 = __frcp_rn(pow4(c4.y));
                kc.exc = c4.z; kc.lr = 0.0f; kc.E = c4.w; kc.kr = 0.9f; kc.kd = 0.1f;
                kc.fa1 = (float)((unsigned)meta >> 13); kc.den = 0.0f; kc.c0 = c0;
                kc.s_q = (cls == 1) ? __fdiv_rn(a.dt, d0 * kc.fa1) : 0.0f; // dt / (0.001 dx^2 (flwacc - 1)), md_routing_operator.f90:56
                const bool far = cons >= 0 && (cons & (1 << 30)) != 0;
                const int nrounds = tp.tile_rounds[tile];
                const int myround = cls == 0 ? -1 : cls == 1 ? (meta >> 5 & 7) : 99;
                const bool exc_on = __any_sync(FULLM, kc.exc != 0.0f);
#pragma unroll 1
                for (int b = 0; b < nbe; b++) {
                    const int w = w0 + b, t0 = w * TK_W;
                    const int nst = min(TK_W, T - t0);
                    float *Xw = a.X + (size_t)w * xstride;
                    // ---- forcing of this window (shared memory): range check first, then the 8 steps straight from the box
                    mbar_wait(&bars[slot], parity);
                    float qv[TK_W];
                    float mn = 0.0f, mx = 0.0f;
#pragma unroll
                    for (int s = 0; s < TK_W; s++) {
                        const float p_ = slots[slot][0][s][lane], e_ = slots[slot][1][s][lane];
                        mn = fminf(mn, fminf(p_, e_));
                        mx = fmaxf(mx, fmaxf(p_, e_));
                    }
                    // ---- reservoirs, 8 steps (md_forward_structure.f90:106-144)
                    const bool full = (tile * 32 + 32 <= n) && nst == TK_W;
                    const float xm = mx * kc.inv_cp;
                    const bool lean = full && __all_sync(FULLM, mn >= 0.0f && xm < 0.25f && fmaf(8.0f, xm, hp) < 15.0f);
                    if (lean) {
                        if (exc_on) {
#pragma unroll
                            for (int s = 0; s < TK_W; s++) qv[s] = tick_step_lean<true>(kc, slots[slot][0][s][lane], slots[slot][1][s][lane], hp, hft);
                        } else {
#pragma unroll
                            for (int s = 0; s < TK_W; s++) qv[s] = tick_step_lean<false>(kc, slots[slot][0][s][lane], slots[slot][1][s][lane], hp, hft);
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < TK_W; s++) {
                            const float p_ = slots[slot][0][s][lane], e_ = slots[slot][1][s][lane];
                            const bool act = valid && s < nst;
                            float hp_n = hp, hft_n = hft, qt;
                            const bool gapless = (p_ >= 0.0f) && (e_ >= 0.0f);
                            if (__all_sync(FULLM, gapless)) qt = vertical_step_nogap(kc, p_, e_, hp_n, hft_n);
                            else qt = vertical_step<1>(kc, p_, e_, hp_n, hft_n).qt;
                            if (act) { hp = hp_n; hft = hft_n; }
                            qv[s] = qt;
                        }
                    }
                    __syncwarp();                                          // every lane is done with the box: ask for the one after the next
                    if (kq <= kmax) {
                        request(slot, kq, iq, bq);
                        next_box(kq, iq, bq);
                    }
                    if (++slot == TK_SLOTS) { slot = 0; parity ^= 1u; }
                    if (a.save_netp && valid) {
                        float *np_ = a.netp + j;
#pragma unroll
                        for (int s = 0; s < TK_W; s++)
                            if (s < nst) __stcs(np_ + (size_t)(t0 + s) * qpitch, qv[s]);
                    }
                    if (valid && cls == 3) {
                        if (far) stx_far(Xw + (size_t)j * TK_W, qv);       // runoff block of a deep cell: its reach routes it
                        else stx_near(Xw + (size_t)j * TK_W, qv);
                    } else if (valid && cls == 2) {
                        stx_far(a.rows + (size_t)j * a.Tp + t0, qv);       // runoff row of a pit cell: route_pairs_kernel
                    }
                    // ---- discharge: the source cells first (round -1), then the shallow routed cells round by round; the cells
                    // of a round only gather cells of earlier rounds of this tile and blocks of other units
#pragma unroll 1
                    for (int r = -1; r < nrounds; r++) {
                        if (valid && myround == r) {
                            if (r < 0) {
#pragma unroll
                                for (int s = 0; s < TK_W; s++) {
                                    qv[s] = qv[s] * c0;                    // md_forward_structure.f90:155 with flwacc - 1 = 0
                                    if (s < nst) hlr = (hlr + 0.0f) * kc.E;   // linear_routing with qup = 0, md_routing_operator.f90:73-77
                                }
                            } else {
                                float qup[TK_W];
#pragma unroll
                                for (int s = 0; s < TK_W; s++) qup[s] = 0.0f;
                                for (int e = 0; e < nup; e++) {            // md_routing_operator.f90:37-53, same order
                                    const int s0 = ldg_keep_i(tp.ups + uo + e, keep) & ~(1 << 30);
                                    float v0[TK_W];
                                    ldx(Xw + (size_t)s0 * TK_W, v0);
#pragma unroll
                                    for (int s = 0; s < TK_W; s++) qup[s] = qup[s] + v0[s];
                                }
#pragma unroll
                                for (int s = 0; s < TK_W; s++) {
                                    const float hr = hlr + qup[s] * kc.s_q;   // :55-56, :73
                                    const float hn = hr * kc.E;            // :75
                                    qv[s] = fmaf(hr - hn, kc.fa1, qv[s]) * c0;   // :77, md_forward_structure.f90:155
                                    if (s < nst) hlr = hn;
                                }
                            }
                            // the cell's 8 discharge values to wherever they are needed
                            if (a.save_q) {
                                float *qd = a.qdom + j;
#pragma unroll
                                for (int s = 0; s < TK_W; s++)
                                    if (s < nst) __stcs(qd + (size_t)(t0 + s) * qpitch, qv[s]);
                            }
                            if (meta & 4) {
                                if (far) stx_far(Xw + (size_t)j * TK_W, qv);
                                else stx_near(Xw + (size_t)j * TK_W, qv);
                            }
                            if (meta & 8) stx_far(a.rows + (size_t)j * a.Tp + t0, qv);
                            if (meta & 16)
                                for (int g = tp.gauge_first[j]; g >= 0; g = tp.gauge_next[g])
#pragma unroll
                                    for (int s = 0; s < TK_W; s++)
                                        if (s < nst) qsim[(size_t)(t0 + s) * ng + g] = qv[s];   // md_forward_structure.f90:206-210
                        }
                        if (r + 1 < nrounds) __syncwarp();                 // blocks of this round before the next round reads them
                    }
                }
                if (valid) {
                    stg_keep_f(fs + j, hp, keep); stg_keep_f(fs + (size_t)npad + j, hft, keep);
                    if (cls <= 1) stg_keep_f(fs + (size_t)2 * npad + j, hlr, keep);
                }
            } else {
                // =============================================================== reach visit
                const int c = tp.reach_cells[(size_t)(unit - ntile) * 32 + lane];
                const bool on = c >= 0;
                const int cj = on ? c : 0;
                const int meta = on ? ldg_keep_i(tp.meta + cj, keep) : 0;
                const float E = ldg_keep_cf(&a.cc[cj].w, keep);
                const float fa1 = (float)((unsigned)meta >> 13);
                const float s_q = on ? __fdiv_rn(a.dt, d0 * fa1) : 0.0f;    // md_routing_operator.f90:56
                float hlr = on ? ldg_keep_f(fs + (size_t)2 * npad + cj, keep) : 0.0f;
                const int nup = meta >> 8 & 15;
                const int uo = ldg_keep_i(tp.upoff + cj, keep);
                const bool chained = (meta & 4096) != 0;
                wait_arrivals(a.cnt + (size_t)v * nunit + unit, me.z, unit, a.err, lane);
                if (on)
                    for (int b = 0; b < nbe; b++) {
                        prefetch_l2_block(a.X + (size_t)(w0 + b) * xstride + (size_t)cj * TK_W);
                        for (int e = 0; e < nup; e++)
                            prefetch_l2_block(a.X + (size_t)(w0 + b) * xstride + (size_t)(ldg_keep_i(tp.ups + uo + e, keep) & ~(1 << 30)) * TK_W);
                    }
                // slope of q_i in q_(i-1): c0 fa1 (1 - E) s_q, composed over 1, 2, 4, 8, 16 lanes (0 where the chain starts)
                float b1 = chained ? c0 * fa1 * (1.0f - E) * s_q : 0.0f, b2, b4, b8, b16;
                {
                    float y = __shfl_up_sync(FULLM, b1, 1); b2 = lane >= 1 ? b1 * y : 0.0f;
                    y = __shfl_up_sync(FULLM, b2, 2); b4 = lane >= 2 ? b2 * y : 0.0f;
                    y = __shfl_up_sync(FULLM, b4, 4); b8 = lane >= 4 ? b4 * y : 0.0f;
                    y = __shfl_up_sync(FULLM, b8, 8); b16 = lane >= 8 ? b8 * y : 0.0f;
                    if (lane < 1) b1 = 0.0f;
                    if (lane < 2) b2 = 0.0f;
                    if (lane < 4) b4 = 0.0f;
                    if (lane < 8) b8 = 0.0f;
                    if (lane < 16) b16 = 0.0f;
                }
#pragma unroll 1
                for (int b = 0; b < nbe; b++) {
                    const int w = w0 + b, t0 = w * TK_W;
                    const int nst = min(TK_W, T - t0);
                    float *Xw = a.X + (size_t)w * xstride;
                    float qt[TK_W], lat[TK_W];
#pragma unroll
                    for (int s = 0; s < TK_W; s++) { qt[s] = 0.0f; lat[s] = 0.0f; }
                    if (on) {
                        ldx(Xw + (size_t)cj * TK_W, qt);
                        for (int e = 0; e < nup; e++) {                    // md_routing_operator.f90:37-53 (the heavy inflow joins last)
                            const int s0 = ldg_keep_i(tp.ups + uo + e, keep) & ~(1 << 30);
                            float v0[TK_W];
                            ldx(Xw + (size_t)s0 * TK_W, v0);
#pragma unroll
                            for (int s = 0; s < TK_W; s++) lat[s] = lat[s] + v0[s];
                        }
                    }
#pragma unroll
                    for (int s = 0; s < TK_W; s++) {
                        // without the heavy inflow ...
                        const float h0 = hlr + lat[s] * s_q;
                        float x = fmaf(h0 - h0 * E, fa1, qt[s]) * c0;
                        // ... then the affine scan down the reach
                        x = fmaf(b1, __shfl_up_sync(FULLM, x, 1), x);
                        x = fmaf(b2, __shfl_up_sync(FULLM, x, 2), x);
                        x = fmaf(b4, __shfl_up_sync(FULLM, x, 4), x);
                        x = fmaf(b8, __shfl_up_sync(FULLM, x, 8), x);
                        x = fmaf(b16, __shfl_up_sync(FULLM, x, 16), x);
                        // the reference's statements with the upstream discharge this lane received
                        const float qprev = __shfl_up_sync(FULLM, x, 1);
                        const float qup = chained ? lat[s] + qprev : lat[s];
                        const float hr = hlr + qup * s_q;                  // md_routing_operator.f90:55-56, :73
                        const float hn = hr * E;                           // :75
                        qt[s] = fmaf(hr - hn, fa1, qt[s]) * c0;            // :77, md_forward_structure.f90:155
                        if (s < nst) hlr = hn;
                    }
                    if (on) {
                        if (a.save_q) {
                            float *qd = a.qdom + cj;
#pragma unroll
                            for (int s = 0; s < TK_W; s++)
                                if (s < nst) qd[(size_t)(t0 + s) * qpitch] = qt[s];
                        }
                        if (meta & 4) stx_near(Xw + (size_t)cj * TK_W, qt);
                        if (meta & 8) stx_far(a.rows + (size_t)cj * a.Tp + t0, qt);
                        if (meta & 16)
                            for (int g = tp.gauge_first[cj]; g >= 0; g = tp.gauge_next[g])
#pragma unroll
                                for (int s = 0; s < TK_W; s++)
                                    if (s < nst) qsim[(size_t)(t0 + s) * ng + g] = qt[s];
                    }
                }
                if (on) stg_keep_f(fs + (size_t)2 * npad + cj, hlr, keep);
            }
        }
        if (a.dbg && lane == 0) {                                          // diagnostics: when the last warp left tick k
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (k < 1024) atomicMax(a.dbg + k, t);
            if (k == kfirst) atomicMin(a.dbg + 1024, t);
        }
        if (!any) continue;
        // ---- publish the visits of this tick: one fence, then one arrival per cell at the counter of the unit that reads it
        fence_release();
#pragma unroll 1
        for (int i = 0; i < nu; i++) {
            const int4 me = mine[i];
            const int v = k - me.y;
            if (v < 0 || v >= nvis) continue;
            int cons = -1;
            if (me.x < ntile) {
                cons = ldg_keep_i(tp.cons1 + me.x * 32 + lane, keep);
            } else {
                const int c = tp.reach_cells[(size_t)(me.x - ntile) * 32 + lane];
                if (c >= 0) cons = ldg_keep_i(tp.cons2 + c, keep);
            }
            if (cons >= 0) red_add(a.cnt + (size_t)v * nunit + (cons & ~(1 << 30)), 1);
        }
    }
}

}  // namespace

static int g_tick_sms = 0, g_tick_per_sm[3] = {0, 0, 0};

template <int MINB> static cudaError_t tick_occupancy(int *blocks) {
    constexpr int vi = MINB == 8 ? 0 : MINB == 6 ? 1 : 2;
    const size_t smem = tick_smem_bytes();
    cudaError_t e;
    if (!g_tick_sms) {
        int dev = 0;
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&g_tick_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    if (!g_tick_per_sm[vi]) {
        e = cudaFuncSetAttribute(tick_forward_kernel<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, tick_forward_kernel<MINB>, TK_WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        g_tick_per_sm[vi] = per < 1 ? 1 : per;
    }
    *blocks = g_tick_sms * g_tick_per_sm[vi];
    return cudaSuccess;
}

size_t tick_smem_bytes() { return (sizeof(TkSlot) + sizeof(uint64_t)) * TK_WARPS * TK_SLOTS + sizeof(int4) * TK_WARPS * TK_MAXU; }

cudaError_t tick_grid_warps(int variant, int ctas_per_sm, int *nwarp) {
    int blocks = 0;
    cudaError_t e = variant == 4 ? tick_occupancy<4>(&blocks) : variant == 6 ? tick_occupancy<6>(&blocks) : tick_occupancy<8>(&blocks);
    if (e != cudaSuccess) return e;
    if (ctas_per_sm > 0) blocks = std::min(blocks, g_tick_sms * ctas_per_sm);
    *nwarp = blocks * TK_WARPS;
    return cudaSuccess;
}

int tick_max_units() { return TK_MAXU; }

cudaError_t launch_tick_forward(const TkArgs &a, const float *fields, const CUtensorMap &prcp, const CUtensorMap &pet,
                                cudaStream_t s, int variant) {
    if (a.nb < 1) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(a.cnt, 0, sizeof(int) * (size_t)(a.tp.ntile + a.tp.nreach) * ((a.nwin + a.nb - 1) / a.nb), s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(a.err, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    tick_prep_kernel<<<(a.tp.npad + 255) / 256, 256, 0, s>>>(a, fields);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)(a.nwarp / TK_WARPS);
    const size_t smem = tick_smem_bytes();
    if (variant == 4) tick_forward_kernel<4><<<blocks, TK_WARPS * 32, smem, s>>>(prcp, pet, a);
    else if (variant == 6) tick_forward_kernel<6><<<blocks, TK_WARPS * 32, smem, s>>>(prcp, pet, a);
    else tick_forward_kernel<8><<<blocks, TK_WARPS * 32, smem, s>>>(prcp, pet, a);
    return cudaGetLastError();
}

}  // namespace smash
