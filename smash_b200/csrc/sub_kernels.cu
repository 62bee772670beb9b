// sub_kernels.cu -- the subtree engine: reservoirs AND routing of the whole domain in one pass over
// the forcing, with the engine owning the cell order.
//
// Host side (build_sub_topo, route_graph.cpp): the drainage forest is cut into connected subtrees of at most 32 cells and
// dmax + 1 cells of depth; subtrees of the same level are packed into tiles of 32 lanes; engine column j' = tile * 32 + lane.
// The forcing is packed once per plan into that order, so a tile's box [8 steps][32 cells] is one 2-D TMA load.
//
// Device side: one warp per tile for the whole run, states in registers.  The warp advances one micro-tick at a time; lane l
// works on time step t = m - delay(l), where delay(root) = depth of its subtree - 1 and delay(child) = delay(parent) - 1.
// Every in-tile edge then spans exactly one micro-tick: the discharge a lane computed at micro-tick m - 1 is what its parent
// gathers at micro-tick m, by a shuffle -- the routing recurrence (md_routing_operator.f90:17-79) is evaluated strictly
// sequentially in time, statement by statement, inside the loop of the reservoirs (md_forward_structure.f90:106-156).
// Only the root of a subtree hands its series to another tile: 8 values per window in an exchange slot X[slot][window][8]
// that starts as NaN; the reader polls the block until all 8 values are numbers (every 4-byte store is atomic, nothing else
// hangs on the block, so neither flags nor fences are needed).  Tiles are numbered by level and a tile only reads lower
// tiles, which the hardware dispatches first: a waiting warp always waits for a warp that is resident or done.
// The discharge of a time step leaves the warp when its slowest lane has produced it (a per-lane delay line in shared
// memory), as one coalesced 128-byte row of qsim_domain in engine order.
//
// Pit pairs only get their runoff here (rows); route_pairs_kernel (split_kernels.cu) runs them afterwards.
//
// Reference statements are cited as file:line under /root/reference/smash/solver/.
#include "split_kernels.cuh"

#include <algorithm>
#include <type_traits>

#include "cell_math.cuh"

namespace smash {

namespace {

constexpr unsigned FULLM = 0xffffffffu;
constexpr int SB_WARPS = 4;     // tiles per CTA; every warp runs on its own, no CTA barrier
constexpr int SB_NST = 2;       // forcing boxes in flight per warp (a box is copied to registers as soon as it has arrived)
constexpr int SB_DL = 16;       // rows of the runoff / discharge ring (8 steps ahead, dmax <= 8 behind)

typedef float SbBox[SB_W][32];         // one forcing box: [step][lane]
constexpr int SB_ROWS = SB_NST * SB_W;  // rows of a forcing ring

__device__ __forceinline__ uint64_t sb_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void sb_tma_load_2d(void *dst, const CUtensorMap *tm, int x, int y, uint64_t *bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void sb_ld8(const float *p, float *v) {       // straight from L2: the block may have been written a moment ago
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p)
                 : "memory");
}

// shared-memory accesses through addresses kept in registers: left to itself the compiler rebuilds every address from the
// thread index (about ten instructions per access at this register budget)
__device__ __forceinline__ uint32_t sb_opaque(uint32_t x) { uint32_t y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float sb_lds(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sb_sts(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// gr-a cell-step without forcing gap, tanh argument below 0.25 and hp_imd <= 15: the statements of vertical_step_nogap
// (cell_math.cuh) with those branches resolved by the caller.  Same code as vertical_step_lean of split_kernels.cu.
template <bool EXC>
__device__ __forceinline__ float sub_step_lean(const CellConst &k, float prcp, float pet, float &hp, float &hft) {
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

template <int DUMMY>
__global__ void __launch_bounds__(SB_WARPS * 32, 7) sub_forward_kernel(const __grid_constant__ CUtensorMap tm_prcp,
                                                                       const __grid_constant__ CUtensorMap tm_pet, const SbArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SbTopo &tp = a.tp;
    const int tile = (int)blockIdx.x * SB_WARPS + warp;
    if (tile >= tp.ntile) return;
    SbBox *ring_p = reinterpret_cast<SbBox *>(smem_raw) + warp * SB_NST;                   // prcp boxes
    SbBox *ring_e = reinterpret_cast<SbBox *>(smem_raw) + (SB_WARPS + warp) * SB_NST;      // pet boxes
    unsigned char *after = smem_raw + sizeof(SbBox) * 2 * SB_WARPS * SB_NST;
    float(*qr)[32] = reinterpret_cast<float(*)[32]>(after) + warp * SB_DL;                 // [16][32]: row t & 15 holds the runoff of time
    after += sizeof(float) * 32 * SB_DL * SB_WARPS;                                        // step t, then (lane by lane) its discharge
    float(*extv)[32] = reinterpret_cast<float(*)[32]>(after) + warp * SB_W;                // [8][32] inflow from other tiles, this window
    after += sizeof(float) * 32 * SB_W * SB_WARPS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(after) + warp * SB_NST;

    const int T = a.T, npad = a.npad, ng = tp.ng, dmax = tp.dmax, nwin = a.nwin;
    const int jp = tile * 32 + lane;
    const int rec = tp.rec[jp];
    const bool valid = (rec & 1) != 0, pit = (rec & 2) != 0, root_out = (rec & 4) != 0, want_row = (rec & 8) != 0, gauge = (rec & 16) != 0;
    const int delay = rec >> 8 & 15, nch = rec >> 12 & 15, next = rec >> 16 & 15, segpos = rec >> 20 & 7, lastc = rec >> 23 & 31;
    const int j = valid ? tp.cell[jp] : 0;
    const bool tile_ext = tp.tile_ext[tile] != 0;
    const uint64_t pol = sb_policy_evict_first();
    constexpr uint32_t STAGE_BYTES = 2 * sizeof(SbBox);
    const int nblk = (T + SB_W - 1) / SB_W;

    if (lane == 0) {
        for (int s = 0; s < SB_NST; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int b = 0; b < SB_NST && b < nblk; b++) {
            mbar_expect_tx(&bars[b], STAGE_BYTES);
            sb_tma_load_2d(&ring_p[b][0][0], &tm_prcp, tile * 32, b * SB_W, &bars[b], pol);
            sb_tma_load_2d(&ring_e[b][0][0], &tm_pet, tile * 32, b * SB_W, &bars[b], pol);
        }
    }
    for (int r = 0; r < SB_DL; r++) qr[r][lane] = 0.0f;
    for (int r = 0; r < SB_W; r++) extv[r][lane] = 0.0f;
    __syncwarp();

    // ---- the cell of this lane
    float hp = 0.01f, hft = 0.01f, hlr = 0.0f;
    CellConst k = make_const(200.0f, 500.0f, 0.0f, 5.0f, 1, a.dt, a.dx);
    if (valid) {
        const float *fld = a.fields + j;
        k = make_const(fld[(size_t)F_CP * npad], fld[(size_t)F_CFT * npad], fld[(size_t)F_EXC * npad], fld[(size_t)F_LR * npad],
                       pit ? 1 : a.flwacc[j], a.dt, a.dx);
        hp = fld[(size_t)F_HP * npad]; hft = fld[(size_t)F_HFT * npad]; hlr = fld[(size_t)F_HLR * npad];
    }
    const float c0 = pit ? 1.0f : k.c0, E = k.E, fa1 = k.fa1, s_q = k.s_q;    // a pit cell hands on its runoff: route_pairs_kernel routes it
    const size_t pitch = (size_t)a.qpitch;
    float *qd = a.qdom + jp;                                              // row of the next complete time step
    float *np_ = a.netp + jp;
    float *xo = root_out ? a.X + (size_t)tp.xout[jp] * nwin * SB_W : nullptr;     // the slot is contiguous in time: xo[t]
    float *rowp = (want_row || pit) ? a.rows + (size_t)j * a.Tp : nullptr;
    const int eoff = tp.extoff[jp];
    float qprev = 0.0f;                                                   // the discharge of the previous micro-tick: what the parent gathers
    const bool exc_on = __any_sync(FULLM, k.exc != 0.0f);
    const bool save_netp = a.save_netp != 0 && valid;
    const bool save_q = a.save_q != 0 && valid && !pit;
    const bool rare = root_out || want_row || pit || gauge;
    const float inv_cp = k.inv_cp;
    // lane flags in one register the compiler cannot rebuild from the kernel arguments every tick
    const uint32_t fl = sb_opaque((save_q ? 1u : 0u) | ((want_row || pit || gauge) ? 2u : 0u) | ((tile_ext && next > 0) ? 4u : 0u) |
                                  (nch > 0 ? 8u : 0u) | (root_out ? 32u : 0u));
    const uint32_t qcol = sb_opaque(smem_u32(&qr[0][lane]));             // this lane's column of the runoff / discharge ring (rows 128 bytes apart)
    const uint32_t ecol = sb_opaque(smem_u32(&extv[0][lane]));
    const uint32_t pcol = sb_opaque(smem_u32(&ring_p[0][0][lane]));       // prcp boxes of this warp; the pet boxes lie SB_WARPS * SB_NST boxes further

    // per-lane clock of the routing wavefront: time step, its row of the ring (byte offset), the row that completes this tick
    int t = -delay;
    uint32_t toff = ((uint32_t)(-delay) & (SB_DL - 1)) << 7;
    const int ttshift = delay - dmax;
    const uint32_t ttadd = ((uint32_t)(delay - dmax) & (SB_DL - 1)) << 7;
    const int ngroup = (T + dmax + SB_W - 1) / SB_W;                      // groups of 8 micro-ticks
    uint32_t parity = 0;
    int slot = 0;
#pragma unroll 1
    for (int kb = 0; kb < ngroup; kb++) {
        // ================= reservoirs of time steps 8 kb .. 8 kb + 7, every lane at the same time (md_forward_structure.f90:106-144)
        if (kb < nblk) {
            mbar_wait(&bars[slot], parity);
            float pv[SB_W], ev[SB_W];
            float mn = 0.0f, mx = 0.0f;
            const uint32_t pbox = pcol + (uint32_t)slot * (uint32_t)sizeof(SbBox), ebox = pbox + (uint32_t)(sizeof(SbBox) * SB_WARPS * SB_NST);
#pragma unroll
            for (int i = 0; i < SB_W; i++) {
                pv[i] = sb_lds(pbox + 128 * i);
                ev[i] = sb_lds(ebox + 128 * i);
                mn = fminf(mn, fminf(pv[i], ev[i]));
                mx = fmaxf(mx, fmaxf(pv[i], ev[i]));
            }
            __syncwarp();                                                 // every lane holds the box in registers: refill the slot
            if (lane == 0 && kb + SB_NST < nblk) {
                mbar_expect_tx(&bars[slot], STAGE_BYTES);
                sb_tma_load_2d(&ring_p[slot][0][0], &tm_prcp, tile * 32, (kb + SB_NST) * SB_W, &bars[slot], pol);
                sb_tma_load_2d(&ring_e[slot][0][0], &tm_pet, tile * 32, (kb + SB_NST) * SB_W, &bars[slot], pol);
            }
            if (++slot == SB_NST) { slot = 0; parity ^= 1u; }
            const int tb = kb * SB_W;
            const bool full = tb + SB_W <= T;                             // empty lanes carry default parameters and zero forcing
            const float xm = mx * inv_cp;
            const bool lean = full && __all_sync(FULLM, mn >= 0.0f && xm < 0.25f && fmaf(8.0f, xm, hp) < 15.0f);
            const uint32_t row = qcol + ((uint32_t)(tb & (SB_DL - 1)) << 7);      // 8 consecutive rows: tb is a multiple of 8
            if (lean) {
                if (exc_on) {
#pragma unroll
                    for (int i = 0; i < SB_W; i++) sb_sts(row + 128 * i, sub_step_lean<true>(k, pv[i], ev[i], hp, hft));
                } else {
#pragma unroll
                    for (int i = 0; i < SB_W; i++) sb_sts(row + 128 * i, sub_step_lean<false>(k, pv[i], ev[i], hp, hft));
                }
            } else {
#pragma unroll
                for (int i = 0; i < SB_W; i++) {
                    const bool act = valid && tb + i < T;
                    float hp_n = hp, hft_n = hft, qt;
                    const bool gapless = (pv[i] >= 0.0f) && (ev[i] >= 0.0f);
                    if (__all_sync(FULLM, gapless)) qt = vertical_step_nogap(k, pv[i], ev[i], hp_n, hft_n);
                    else qt = vertical_step<1>(k, pv[i], ev[i], hp_n, hft_n).qt;
                    if (act) { hp = hp_n; hft = hft_n; }
                    sb_sts(row + 128 * i, qt);
                }
            }
            if (save_netp) {
#pragma unroll
                for (int i = 0; i < SB_W; i++)
                    if (tb + i < T) __stcs(np_ + (size_t)(tb + i) * pitch, sb_lds(row + 128 * i));
            }
        }
        // ================= routing, micro-ticks 8 kb .. 8 kb + 7: lane l is at time step t = m - delay(l).  In the interior groups
        // every lane is inside [0, T) at every tick and the guards fall away.
        auto tick = [&](auto guarded_t) {
            constexpr bool GUARDED = decltype(guarded_t)::value;
            const bool active = GUARDED ? (valid && (unsigned)t < (unsigned)T) : true;
            // ---- inflow blocks of other tiles: once per window of this lane (the lane's column of extv is its own)
            if ((fl & 4u) && active && (toff & 0x380u) == 0u) {
                float acc[SB_W];
#pragma unroll
                for (int i = 0; i < SB_W; i++) acc[i] = 0.0f;
                for (int e = 0; e < next; e++) {                          // md_routing_operator.f90:37-53 (inflows of other tiles first)
                    const float *blk = a.X + ((size_t)tp.extlist[eoff + e] * nwin + (t >> 3)) * SB_W;
                    float v[SB_W];
                    int spins = 0;
                    for (;;) {
                        sb_ld8(blk, v);
                        const float chk = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
                        if (chk == chk || a.nowait) break;                // a NaN anywhere makes the sum a NaN
                        __nanosleep(100);
                        if (++spins > (1 << 21)) { atomicExch(a.err, 1 + tile); break; }
                    }
#pragma unroll
                    for (int i = 0; i < SB_W; i++) acc[i] = acc[i] + v[i];
                }
#pragma unroll
                for (int i = 0; i < SB_W; i++) sb_sts(ecol + 128 * i, acc[i]);
            }
            // ---- the in-tile inflows sit in consecutive lanes: segmented sum of last micro-tick's discharge, then the parent reads
            // the lane of its last inflow (md_routing_operator.f90:37-53)
            float x = qprev, v;
            v = __shfl_up_sync(FULLM, x, 1); x = segpos >= 1 ? x + v : x;
            v = __shfl_up_sync(FULLM, x, 2); x = segpos >= 2 ? x + v : x;
            v = __shfl_up_sync(FULLM, x, 4); x = segpos >= 4 ? x + v : x;
            v = __shfl_sync(FULLM, x, lastc);
            float qup = (fl & 8u) ? v : 0.0f;
            if (fl & 4u) qup += sb_lds(ecol + (toff & 0x380u));
            const uint32_t qcell = qcol + toff;
            const float qt = sb_lds(qcell);
            const float hr = hlr + qup * s_q;                             // :55-56, :73
            const float hn = hr * E;                                      // :75
            const float q = fmaf(hr - hn, fa1, qt) * c0;                  // :77, md_forward_structure.f90:155
            if (GUARDED) { hlr = active ? hn : hlr; qprev = active ? q : 0.0f; }
            else { hlr = hn; qprev = q; }
            if (active) {
                sb_sts(qcell, q);
                if (fl & 32u) xo[t] = q;                                  // the root of a subtree: its series goes to another tile
                if (fl & 2u) {
                    if (rowp) rowp[t] = q;
                    if (gauge && !pit)
                        for (int g = a.gauge_first[j]; g >= 0; g = a.gauge_next[g]) a.qsim[(size_t)t * ng + g] = q;   // :206-210
                }
            }
            // the row of time step m - dmax = t + delay - dmax is complete: one coalesced row of the domain series
            if ((fl & 1u) && (!GUARDED || (unsigned)(t + ttshift) < (unsigned)T)) {
                __stcs(qd, sb_lds(qcol + ((toff + ttadd) & (SB_DL * 128u - 1u))));
                qd += pitch;
            }
            t++;
            toff = (toff + 128u) & (SB_DL * 128u - 1u);
        };
        const int m0 = kb * SB_W;
        if (m0 >= dmax && m0 + SB_W <= T) {
#pragma unroll 1
            for (int mi = 0; mi < SB_W; mi++) tick(std::false_type());
        } else {
#pragma unroll 1
            for (int mi = 0; mi < SB_W; mi++) tick(std::true_type());
        }
    }
    // the reader of an exchange block waits for 8 numbers: fill what lies beyond the last time step
    if (root_out && (T & 7) != 0)
        for (int i = T & 7; i < SB_W; i++) xo[(size_t)(nwin - 1) * SB_W + i] = 0.0f;
    if (valid) {
        float *fs = a.fstates + j;
        fs[0] = hp; fs[(size_t)npad] = hft;
        if (!pit) fs[(size_t)2 * npad] = hlr;
    }
}

}  // namespace

size_t sub_smem_bytes() {
    return (2 * sizeof(SbBox) * SB_NST + sizeof(float) * 32 * SB_DL + sizeof(float) * 32 * SB_W + sizeof(uint64_t) * SB_NST) * SB_WARPS;
}

cudaError_t launch_sub_forward(const SbArgs &a, const CUtensorMap &prcp, const CUtensorMap &pet, cudaStream_t s) {
    static bool attr = false;
    const size_t smem = sub_smem_bytes();
    cudaError_t e;
    if (!attr) {
        e = cudaFuncSetAttribute(sub_forward_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    e = cudaMemsetAsync(a.X, 0xff, sizeof(float) * (size_t)std::max(1, a.tp.nslot) * a.nwin * SB_W, s);   // every value a NaN
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(a.err, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)((a.tp.ntile + SB_WARPS - 1) / SB_WARPS);
    sub_forward_kernel<0><<<blocks, SB_WARPS * 32, smem, s>>>(prcp, pet, a);
    return cudaGetLastError();
}

}  // namespace smash
