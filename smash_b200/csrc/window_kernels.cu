// window_kernels.cu -- the window pass of the split engine: reservoirs AND the routing of the
// shallow part of the drainage forest in one persistent kernel that walks the time axis in windows of 8 steps.
//
// Why: the reservoir pass (md_forward_structure.f90:106-144) wants [t][cell] (coalesced over cells), the routing
// (md_routing_operator.f90:17-79) needs every cell's inflow series.  Writing whole series as rows for a later routing
// pass costs 2-3x the algorithmic DRAM traffic.  Here every cell advances 8 steps at a time, in `path` order, and hands
// its 8 discharge values to its consumer through a small exchange buffer X[parity][cell][8] that is overwritten every
// second window: it stays in the 126 MB L2 and never reaches DRAM.  90 % of a large domain (cells whose flow
// accumulation is at most `shallow_acc`) is routed this way, strictly sequentially in time and in the reference's
// summation order.  The remaining deep cells (main rivers: long serial chains) only get their reservoir series here, as
// rows in DRAM, and are routed afterwards by the chain scans of split_kernels.cu.
//
// Cell classes (host: build_window_topo in route_graph.cpp):
//   S  source, flwacc == 1: q = qt * dx^2 * 1e-3 / dt                       (md_forward_structure.f90:155)
//   R  shallow routed cell: all inflows are S or R cells earlier in path
//   D  deep cell (flwacc > shallow_acc, pit pairs, anything downstream of a D cell)
// Work unit ("ticket") = one tile of 32 consecutive cells x one window; tickets are numbered window-major and dealt to
// the warps of a fully resident grid round-robin, so the warp that owns the lowest unfinished ticket is always working
// on it and only ever waits for lower tickets: deadlock-free.  prog[tile] = windows finished by the tile.
//
// Reference statements are cited as file:line under /root/reference/smash/solver/.
#include "split_kernels.cuh"

#include <algorithm>

#include "cell_math.cuh"

namespace smash {

namespace {

constexpr unsigned FULLM = 0xffffffffu;
constexpr int WF_WARPS = 4;    // warps per CTA; every warp runs its own pipeline, no CTA barrier
constexpr int WF_SLOTS = 2;    // forcing boxes in flight per warp = tickets requested ahead
constexpr int WF_MAXUP = 8;    // inflows of a D8 cell

typedef float WfSlot[2][WF_W][32];   // prcp, pet: [step][lane]

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
__device__ __forceinline__ void ld8cg(const float *p, float *v) {
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st8wb(float *p, const float *v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// gr-a cell-step for a window without forcing gap, tanh arguments below 0.25 and hp_imd <= 15 (1 + (hp_imd/1000)^4 == 1
// in float32): the statements of vertical_step_nogap (cell_math.cuh) with those warp-uniform branches resolved once per
// window.  Same code as vertical_step_lean of split_kernels.cu: bit-identical results.
template <bool EXC>
__device__ __forceinline__ float window_step_lean(const CellConst &k, float prcp, float pet, float &hp, float &hft) {
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

__global__ void __launch_bounds__(256) window_prep_kernel(const WfArgs a, const float *fields) {
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
    // the routing state of the deep cells is carried by the chain scans (hcar); theirs is set at the end of the run
    if ((a.tp.meta[j] & 3) != 2) a.fstates[(size_t)2 * npad + j] = hlr;
}

template <int MINB>   // CTAs per SM the register allocation aims at (8: 64 registers, 6: 80, 4: no cap)
__global__ void __launch_bounds__(WF_WARPS * 32, MINB) window_forward_kernel(const __grid_constant__ CUtensorMap tm_prcp,
                                                                             const __grid_constant__ CUtensorMap tm_pet, const WfArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WfSlot *slots = reinterpret_cast<WfSlot *>(smem_raw) + warp * WF_SLOTS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + sizeof(WfSlot) * WF_WARPS * WF_SLOTS) + warp * WF_SLOTS;
    const WfTopo &tp = a.tp;
    const int ntile = tp.ntile, n = tp.n, npad = tp.npad, T = a.T, ng = tp.ng;
    const int G = (int)gridDim.x * WF_WARPS;                              // warps of the grid: all resident
    const int gw = (int)blockIdx.x * WF_WARPS + warp;
    const uint64_t pol = policy_evict_first();
    const float c0 = a.dx * a.dx * 0.001f / a.dt;                         // md_forward_structure.f90:155
    constexpr uint32_t SLOT_BYTES = sizeof(WfSlot);

    // ticket cursors (window, tile): the one being worked on and the one whose forcing is requested next
    int w = a.w_begin, tile = gw;
    while (tile >= ntile) { tile -= ntile; w++; }
    int wq = w, tq = tile;
    auto advance = [&](int &ww, int &tt) {
        tt += G;
        while (tt >= ntile) { tt -= ntile; ww++; }
    };
    if (lane == 0) {
        for (int s = 0; s < WF_SLOTS; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < WF_SLOTS; s++) {
            if (wq < a.w_end) {
                mbar_expect_tx(&bars[s], SLOT_BYTES);
                tma_load_2d_hint(&slots[s][0][0][0], &tm_prcp, tq * 32, wq * WF_W, &bars[s], pol);
                tma_load_2d_hint(&slots[s][1][0][0], &tm_pet, tq * 32, wq * WF_W, &bars[s], pol);
            }
            advance(wq, tq);
        }
    } else {
        for (int s = 0; s < WF_SLOTS; s++) advance(wq, tq);
    }
    __syncwarp();

    float *fs = a.fstates;
    float *qsim = a.qsim;
    const size_t qpitch = (size_t)a.qpitch;
    uint32_t parity = 0;
    int slot = 0;
#pragma unroll 1
    for (; w < a.w_end; advance(w, tile)) {
        const int j = tile * 32 + lane;
        const bool valid = j < n;
        // ---- per-cell records of this ticket (L2-resident planes)
        const int meta = tp.meta[j];
        const float4 c4 = a.cc[j];
        // the states were left by whichever warp ran this tile's previous window: wait for it (a lower ticket), and read
        // them from L2 -- this SM's L1 may still hold the line from an earlier window
        if (w > a.w_begin || w > 0) {
            if (lane == 0)
                while (ld_acquire(a.prog + tile) < w) __nanosleep(64);
            __syncwarp();
        }
        float hp = __ldcg(fs + j), hft = __ldcg(fs + (size_t)npad + j);
        const int cls = meta & 3;
        float hlr = (cls != 2) ? __ldcg(fs + (size_t)2 * npad + j) : 0.0f;
        const int nup = (cls == 1) ? (meta >> 8 & 15) : 0;
        int us[WF_MAXUP];
        if (nup > 0) {
            const int uo = tp.upoff[j];
#pragma unroll
            for (int e = 0; e < WF_MAXUP; e++) us[e] = (e < nup) ? tp.ups[uo + e] : -1;
        } else {
#pragma unroll
            for (int e = 0; e < WF_MAXUP; e++) us[e] = -1;
        }
        CellConst k;
        k.cp = c4.x; k.inv_cp = __frcp_rn(c4.x);                          // md_gr_operator.f90:47
        k.cft = c4.y; k.inv_cft = __frcp_rn(c4.y); k.cft_m4 = __frcp_rn(pow4(c4.y));
        k.exc = c4.z; k.lr = 0.0f; k.E = c4.w;
        k.fa1 = (float)(meta >> 12);
        k.den = 0.001f * a.dx * a.dx * k.fa1;                             // md_routing_operator.f90:56
        k.s_q = (cls == 1) ? __fdiv_rn(a.dt, k.den) : 0.0f;
        k.c0 = c0;

        // ---- forcing of this ticket: shared memory -> registers, then ask for the ticket after the next one
        mbar_wait(&bars[slot], parity);
        float pv[WF_W], ev[WF_W], qv[WF_W];
        float mn = 0.0f, mx = 0.0f;
#pragma unroll
        for (int i = 0; i < WF_W; i++) {
            pv[i] = slots[slot][0][i][lane];
            ev[i] = slots[slot][1][i][lane];
            mn = fminf(mn, fminf(pv[i], ev[i]));
            mx = fmaxf(mx, fmaxf(pv[i], ev[i]));
        }
        __syncwarp();
        if (lane == 0 && wq < a.w_end) {
            mbar_expect_tx(&bars[slot], SLOT_BYTES);
            tma_load_2d_hint(&slots[slot][0][0][0], &tm_prcp, tq * 32, wq * WF_W, &bars[slot], pol);
            tma_load_2d_hint(&slots[slot][1][0][0], &tm_pet, tq * 32, wq * WF_W, &bars[slot], pol);
        }
        advance(wq, tq);

        // ---- reservoirs, 8 steps (md_forward_structure.f90:106-144)
        const int t0 = w * WF_W;
        const int nst = min(WF_W, T - t0);
        const bool full = (tile * 32 + 32 <= n) && nst == WF_W;
        const float xm = mx * k.inv_cp;
        const bool exc_on = __any_sync(FULLM, k.exc != 0.0f);
        const bool lean = full && __all_sync(FULLM, mn >= 0.0f && xm < 0.25f && fmaf(8.0f, xm, hp) < 15.0f);
        if (lean) {
            if (exc_on) {
#pragma unroll
                for (int i = 0; i < WF_W; i++) qv[i] = window_step_lean<true>(k, pv[i], ev[i], hp, hft);
            } else {
#pragma unroll
                for (int i = 0; i < WF_W; i++) qv[i] = window_step_lean<false>(k, pv[i], ev[i], hp, hft);
            }
        } else {
#pragma unroll
            for (int i = 0; i < WF_W; i++) {
                const bool act = valid && i < nst;
                float hp_n = hp, hft_n = hft, qt;
                const bool gapless = (pv[i] >= 0.0f) && (ev[i] >= 0.0f);
                if (__all_sync(FULLM, gapless)) qt = vertical_step_nogap(k, pv[i], ev[i], hp_n, hft_n);
                else qt = vertical_step<1>(k, pv[i], ev[i], hp_n, hft_n).qt;
                if (act) { hp = hp_n; hft = hft_n; }
                qv[i] = qt;
            }
        }
        if (a.save_netp && valid) {
            float *np_ = a.netp + j;
#pragma unroll
            for (int i = 0; i < WF_W; i++)
                if (i < nst) __stcs(np_ + (size_t)(t0 + i) * qpitch, qv[i]);
        }

        // ---- exchange buffer of this window; the consumer must have left the slot (it read it NX windows ago)
        const int nx = a.nx;
        float *Xw = a.X + (size_t)(w % nx) * npad * WF_W;
        const bool want_x = (meta & 4) != 0, want_row = (meta & 8) != 0, gauge = (meta & 16) != 0;
        if (want_x && w >= nx) {
            const int ct = tp.down[j] >> 5;
            while (ld_acquire(a.prog + ct) < w - nx + 1) __nanosleep(64);
        }
        auto emit = [&](const float *q) {                                   // one cell's 8 discharge values to wherever they are needed
            if (a.save_q) {
                float *qd = a.qdom + j;
#pragma unroll
                for (int i = 0; i < WF_W; i++)
                    if (i < nst) __stcs(qd + (size_t)(t0 + i) * qpitch, q[i]);
            }
            if (want_x) st8wb(Xw + (size_t)j * WF_W, q);
            if (want_row) st8wb(a.rows + (size_t)j * a.Tp + t0, q);
            if (gauge)
                for (int g = tp.gauge_first[j]; g >= 0; g = tp.gauge_next[g])
#pragma unroll
                    for (int i = 0; i < WF_W; i++)
                        if (i < nst) qsim[(size_t)(t0 + i) * ng + g] = q[i];      // md_forward_structure.f90:206-210
        };
        if (valid && cls == 0) {
#pragma unroll
            for (int i = 0; i < WF_W; i++) {
                qv[i] = qv[i] * c0;                                        // :155 with flwacc - 1 = 0
                if (i < nst) hlr = (hlr + 0.0f) * k.E;                     // linear_routing with qup = 0, md_routing_operator.f90:73-77
            }
            emit(qv);
        } else if (valid && cls == 2) {
            st8wb(a.rows + (size_t)j * a.Tp + t0, qv);                    // qt row of a deep cell: routed by the chain scans
        }
        const int nrounds = tp.tile_rounds[tile];
        if (nrounds > 0) {
            // inflows produced by other tiles: published when the producing tile finished this window
            if (nup > 0) {
#pragma unroll
                for (int e = 0; e < WF_MAXUP; e++)
                    if (e < nup && (us[e] >> 5) != tile)
                        while (ld_acquire(a.prog + (us[e] >> 5)) <= w) __nanosleep(40);
            }
            __syncwarp();                                                  // the source cells of this tile have written their X blocks
            const int myround = meta >> 5 & 7;
#pragma unroll 1
            for (int r = 0; r < nrounds; r++) {
                if (valid && cls == 1 && myround == r) {
                    float qup[WF_W];
#pragma unroll
                    for (int i = 0; i < WF_W; i++) qup[i] = 0.0f;
#pragma unroll
                    for (int e = 0; e < WF_MAXUP; e++)                     // md_routing_operator.f90:37-53, same order
                        if (e < nup) {
                            float v[WF_W];
                            ld8cg(Xw + (size_t)us[e] * WF_W, v);
#pragma unroll
                            for (int i = 0; i < WF_W; i++) qup[i] = qup[i] + v[i];
                        }
#pragma unroll
                    for (int i = 0; i < WF_W; i++) {
                        const float hr = hlr + qup[i] * k.s_q;             // :55-56, :73
                        const float hn = hr * k.E;                         // :75
                        qv[i] = fmaf(hr - hn, k.fa1, qv[i]) * c0;          // :77, md_forward_structure.f90:155
                        if (i < nst) hlr = hn;
                    }
                    emit(qv);
                }
                __syncwarp();                                              // X blocks of this round before the next round reads them
            }
        }
        if (valid) {
            fs[j] = hp; fs[(size_t)npad + j] = hft;
            if (cls != 2) fs[(size_t)2 * npad + j] = hlr;
        }
        // every lane's stores happen before the barrier, the release store after it (cumulativity)
        __syncwarp();
        if (lane == 0) st_release(a.prog + tile, w + 1);
        if (++slot == WF_SLOTS) { slot = 0; parity ^= 1u; }
    }
}

}  // namespace

template <int MINB>
static cudaError_t window_launch(const WfArgs &a, const CUtensorMap &prcp, const CUtensorMap &pet, cudaStream_t s, int ctas_per_sm) {
    static int sms = 0, per_sm = 0;
    const size_t smem = sizeof(WfSlot) * WF_WARPS * WF_SLOTS + sizeof(uint64_t) * WF_WARPS * WF_SLOTS;
    cudaError_t e;
    if (!sms) {
        int dev = 0;
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(window_forward_kernel<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, window_forward_kernel<MINB>, WF_WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
    }
    const long long tickets = (long long)(a.w_end - a.w_begin) * a.tp.ntile;
    int per = per_sm;
    if (ctas_per_sm > 0) per = std::min(per, ctas_per_sm);
    long long blocks = (long long)sms * per;                              // every CTA resident: the round-robin deal needs it
    blocks = std::min(blocks, (tickets + WF_WARPS - 1) / WF_WARPS);
    if (blocks < 1) return cudaSuccess;
    window_forward_kernel<MINB><<<(unsigned)blocks, WF_WARPS * 32, smem, s>>>(prcp, pet, a);
    return cudaGetLastError();
}

cudaError_t launch_window_forward(const WfArgs &a, const float *fields, const CUtensorMap &prcp, const CUtensorMap &pet, cudaStream_t s,
                                  int ctas_per_sm, int variant) {
    if (a.w_begin == 0) {
        cudaError_t e = cudaMemsetAsync(a.prog, 0, sizeof(int) * (size_t)a.tp.ntile, s);
        if (e != cudaSuccess) return e;
        window_prep_kernel<<<(a.tp.npad + 255) / 256, 256, 0, s>>>(a, fields);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (variant == 4) return window_launch<4>(a, prcp, pet, s, ctas_per_sm);
    if (variant == 6) return window_launch<6>(a, prcp, pet, s, ctas_per_sm);
    return window_launch<8>(a, prcp, pet, s, ctas_per_sm);
}

}  // namespace smash
