// struct_kernels.cu -- the runoff-production part of the structures other than gr-a (forward only):
//   gr-b   md_forward_structure.f90:216-398   interception store + gr-a's production / exchange / transfer
//   gr-c   md_forward_structure.f90:400-587   gr-b with a second (slow) transfer store
//   gr-d   md_forward_structure.f90:589-760   production + one transfer store, no exchange
//   vic-a  md_forward_structure.f90:762-931   VIC infiltration, vertical transfer, interflow, baseflow (md_vic_operator.f90)
// Every structure ends in the same upstream_discharge + linear_routing, so this pass only replaces vertical_forward of the
// split engine (split_kernels.cu): one thread per cell for the whole run, stores in registers, forcing boxes [8 steps][32 cells]
// by 2-D TMA into a per-warp mbarrier ring; it leaves the runoff series qt of every cell as a row [cell][time] (and the final
// discharge of the cells without inflow), which route_forward then turns into q.
//
// Reference statements are cited as file:line under /root/reference/smash/solver/.
#include "split_kernels.cuh"

#include "../../include/smash_b200.h"
#include "cell_math.cuh"

namespace smash {
namespace {

constexpr unsigned FULLM = 0xffffffffu;
constexpr int ST_TK = 8;       // time steps per TMA box
constexpr int ST_WARPS = 8;    // warps per CTA, every warp runs its own pipeline
constexpr int ST_NST = 3;      // boxes in flight per warp and array
typedef float StStage[2][ST_TK][32];

__device__ __forceinline__ void st_tma_load_2d(void *dst, const CUtensorMap *tm, int x, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void st_st8(float *p, const float *v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// parameters and stores of one cell; which members are live depends on the structure
struct SPar {
    float ci, cp, inv_cp, cft, cst, exc, b, cusl1, cusl2, clsl, ks, ds, dsm, ws;
    // reciprocals of the capacities (fast math mode: a division by a per-cell constant becomes a product)
    float inv_ci, inv_cft, inv_cst, inv_cusl1, inv_cusl2, inv_clsl;
};
// a / b with b a per-cell constant whose reciprocal is at hand
template <int FAST> __device__ __forceinline__ float cdiv(float a, float b, float inv_b) { return FAST ? a * inv_b : a / b; }
// the two powers of vic_infiltration are subtracted from 1 (from cusl - wusl) right away: the result keeps two to three digits fewer
// than the power, so libm's powf stays in the fast math mode too (__powf moved the France checksum by 7e-4)
struct SSto {
    float hi, hp, hft, hst, husl1, husl2, hlsl;
};

// gr_production md_gr_operator.f90:36-67 (beta = 1000)
template <int FAST> __device__ __forceinline__ void d_production(float pn, float en, float cp, float inv_cp, float &hp, float &pr, float &perc) {
    float hp_imd;
    if (FAST && !(pn > 0.0f && en > 0.0f)) {
        // one of pn, en is zero (always so behind gr_interception and the gr-d statements): tanh(0) = 0 leaves one of ps, es
        const bool wet = pn > 0.0f;
        const float th = ftanh<FAST>((wet ? pn : en) * inv_cp);
        const float num = (wet ? cp * (1.0f - hp * hp) : (hp * cp) * (2.0f - hp)) * th;     // :52, :55
        const float r = __fdividef(num, fmaf(wet ? hp : 1.0f - hp, th, 1.0f));
        hp_imd = fmaf(wet ? r : -r, inv_cp, hp);                                             // :58
    } else {
        const float tp = ftanh<FAST>(pn * inv_cp), te = ftanh<FAST>(en * inv_cp);
        const float ps = fdiv<FAST>((cp * (1.0f - hp * hp)) * tp, 1.0f + hp * tp);           // :52
        const float es = fdiv<FAST>(((hp * cp) * (2.0f - hp)) * te, 1.0f + (1.0f - hp) * te);    // :55
        hp_imd = hp + (ps - es) * inv_cp;                                                    // :58
    }
    pr = (pn > 0.0f) ? pn - (hp_imd - hp) * cp : 0.0f;                                       // :60-62
    const float w = 1.0f + pow4(hp_imd * 0.001f);                                            // :64
    const float pw = (w == 1.0f) ? 1.0f : pow_m025<FAST>(w);
    perc = (hp_imd * cp) * (1.0f - pw);
    hp = hp_imd - perc * inv_cp;                                                             // :66
}

// x -> ((x c)^-4 + c^-4)^(-1/4) / c and the released depth (x - that) * c: gr_transfer md_gr_operator.f90:104-106 and
// vic_interflow md_vic_operator.f90:130-132 with n = 5.  FAST: the form without cancellation of cell_math.cuh.
template <int FAST> __device__ __forceinline__ float d_power_store(float h_imd, float c, float &h_new) {
    if (FAST) {
        const float z = pow4(h_imd);
        const float s2 = fsqrt_fast(1.0f + z), s1 = fsqrt_fast(s2);
        const float rel = h_imd * __fdividef(z, s1 * (s1 + 1.0f) * (s2 + 1.0f));
        h_new = h_imd - rel;
        return rel * c;
    }
    const float x1 = h_imd * c;
    h_new = (1.0f / sqrtf(sqrtf(1.0f / pow4(x1) + 1.0f / pow4(c)))) / c;
    return (h_imd - h_new) * c;
}

// gr_transfer(n = 5) md_gr_operator.f90:81-110
template <int FAST> __device__ __forceinline__ float d_transfer(float prcp, float pr, float ct, float inv_ct, float &ht) {
    float pr_imd = pr;
    if (prcp < 0.0f) {                                                                       // :95-96 forcing gap: emptying
        const float x = ht * ct;
        pr_imd = 1.0f / sqrtf(sqrtf(1.0f / pow4(x) - 1.0f / pow4(ct))) - x;
    }
    const float ht_imd = fmaxf(1.e-6f, FAST ? fmaf(pr_imd, inv_ct, ht) : ht + pr_imd / ct);   // :102
    float ht_new;
    const float q = d_power_store<FAST>(ht_imd, ct, ht_new);
    ht = ht_new;
    return q;
}

// vic_infiltration md_vic_operator.f90:21-77
template <int FAST> __device__ __forceinline__ float d_vic_infiltration(float prcp, const SPar &p, float &husl1, float &husl2) {
    const float cusl1 = p.cusl1, cusl2 = p.cusl2;
    const float bp1 = p.b + 1.0f;
    float ifl;
    if (prcp <= 0.0f) {
        ifl = 0.0f;
    } else {
        const float cusl = cusl1 + cusl2;
        float wusl = husl1 * cusl1 + husl2 * cusl2;
        wusl = fmaxf(1.e-6f, wusl);
        wusl = fminf(cusl - 1e-6f, wusl);
        const float iflm = cusl * bp1;
        const float iflc = iflm * (1.0f - powf(1.0f - fdiv<FAST>(wusl, cusl), fdiv<FAST>(1.0f, bp1)));
        if (iflc + prcp >= iflm) ifl = cusl - wusl;
        else ifl = (cusl - wusl) - cusl * powf(1.0f - fdiv<FAST>(iflc + prcp, iflm), bp1);
        ifl = fminf(prcp, ifl);
    }
    const float ifl_usl1 = fminf((1.0f - husl1) * cusl1, ifl);
    ifl = ifl - ifl_usl1;
    const float ifl_usl2 = fminf((1.0f - husl2) * cusl2, ifl);
    husl1 = husl1 + cdiv<FAST>(ifl_usl1, cusl1, p.inv_cusl1);
    husl2 = husl2 + cdiv<FAST>(ifl_usl2, cusl2, p.inv_cusl2);
    return prcp - (ifl_usl1 + ifl_usl2);
}

// brooks_and_corey_flow md_vic_operator.f90:165-183 as called (:88, :92): residual 0, porosity 1, lambda 1
__device__ __forceinline__ float d_brooks_corey(float ks, float c_upper, float c_lower, float h_upper, float h_lower) {
    const float flow = ks * h_upper;
    const float max_flow = fminf(h_upper * c_upper, c_lower - h_lower * c_lower);
    return fminf(max_flow, flow);
}

// vic_vertical_transfer md_vic_operator.f90:79-114 (linear_evapotranspiration :185-200 inlined: min(c h, e h))
template <int FAST> __device__ __forceinline__ void d_vic_vertical_transfer(float pet, const SPar &p, float &husl1, float &husl2, float &hlsl) {
    float fbc = d_brooks_corey(p.ks, p.cusl1, p.cusl2, husl1, husl2);
    husl1 = husl1 - cdiv<FAST>(fbc, p.cusl1, p.inv_cusl1);
    husl2 = husl2 + cdiv<FAST>(fbc, p.cusl2, p.inv_cusl2);
    fbc = d_brooks_corey(p.ks, p.cusl2, p.clsl, husl2, hlsl);
    husl2 = husl2 - cdiv<FAST>(fbc, p.cusl2, p.inv_cusl2);
    hlsl = hlsl + cdiv<FAST>(fbc, p.clsl, p.inv_clsl);
    float fe = fminf(p.cusl1 * husl1, pet * husl1);
    husl1 = husl1 - cdiv<FAST>(fe, p.cusl1, p.inv_cusl1);
    float pet_remain = fmaxf(0.0f, pet - fe);
    fe = fminf(p.cusl2 * husl2, pet_remain * husl2);
    husl2 = husl2 - cdiv<FAST>(fe, p.cusl2, p.inv_cusl2);
    pet_remain = fmaxf(0.0f, pet_remain - fe);
    fe = fminf(p.clsl * hlsl, pet_remain * hlsl);
    hlsl = hlsl - cdiv<FAST>(fe, p.clsl, p.inv_clsl);
}

// vic_baseflow md_vic_operator.f90:137-163
template <int FAST> __device__ __forceinline__ float d_vic_baseflow(const SPar &p, float &hlsl) {
    float qb;
    if (hlsl <= p.ws) qb = fdiv<FAST>(p.ds * p.dsm, p.ws) * hlsl;
    else qb = fdiv<FAST>(p.dsm * (1.0f - fdiv<FAST>(p.ds, p.ws)) * (hlsl - p.ws), 1.0f - p.ws);
    qb = fminf(p.clsl * hlsl, qb);
    hlsl = hlsl - cdiv<FAST>(qb, p.clsl, p.inv_clsl);
    return qb;
}

// one cell-step of the runoff-production part; returns qt
// NOGAP: the caller has checked that no lane of the warp has a forcing gap in this box (the usual case)
template <int ST, int FAST, bool NOGAP = false> __device__ __forceinline__ float struct_step(const SPar &p, float prcp, float pet, SSto &s) {
    const bool nogap = NOGAP || ((prcp >= 0.0f) && (pet >= 0.0f));
    if (ST == SMASH_STRUCTURE_VIC_A) {
        float runoff = 0.0f;
        if (nogap) {
            runoff = d_vic_infiltration<FAST>(prcp, p, s.husl1, s.husl2);            // md_forward_structure.f90:843
            d_vic_vertical_transfer<FAST>(pet, p, s.husl1, s.husl2, s.hlsl);                       // :851
        }
        float h2;
        const float qi = d_power_store<FAST>(s.husl2, p.cusl2, h2);                          // :861
        s.husl2 = h2;
        const float qb = d_vic_baseflow<FAST>(p, s.hlsl);                                          // :863
        return runoff + qi + qb;                                                             // :866
    }
    float pr = 0.0f, perc = 0.0f, l = 0.0f;
    if (nogap) {
        float ei, pn;
        if (ST == SMASH_STRUCTURE_GR_D) {
            ei = fminf(pet, prcp);                                                           // :666
            pn = fmaxf(0.0f, prcp - ei);                                                     // :668
        } else {                                                                             // gr_interception md_gr_operator.f90:20-34
            ei = fminf(pet, prcp + s.hi * p.ci);
            pn = fmaxf(0.0f, prcp - p.ci * (1.0f - s.hi) - ei);
            s.hi = s.hi + cdiv<FAST>(prcp - ei - pn, p.ci, p.inv_ci);
        }
        const float en = pet - ei;
        d_production<FAST>(pn, en, p.cp, p.inv_cp, s.hp, pr, perc);                          // :306, :490, :676
        if (ST != SMASH_STRUCTURE_GR_D)                                                      // gr_exchange :313, :497
            l = FAST ? p.exc * ((s.hft * s.hft) * s.hft * fsqrt_fast(s.hft)) : ((p.exc == 0.0f) ? 0.0f : p.exc * pow_3p5(s.hft));
    }
    if (ST == SMASH_STRUCTURE_GR_B) {
        const float prr = 0.9f * (pr + perc) + l;                                            // :321
        const float prd = 0.1f * (pr + perc);                                                // :322
        const float qr = d_transfer<FAST>(prcp, prr, p.cft, p.inv_cft, s.hft);                          // :324
        return qr + fmaxf(0.0f, prd + l);                                                    // :326-328
    }
    if (ST == SMASH_STRUCTURE_GR_C) {
        const float prr = 0.9f * 0.6f * (pr + perc) + l;                                     // :505
        const float prl = 0.9f * 0.4f * (pr + perc);                                         // :506
        const float prd = 0.1f * (pr + perc);                                                // :507
        const float qr = d_transfer<FAST>(prcp, prr, p.cft, p.inv_cft, s.hft);                          // :509
        const float ql = d_transfer<FAST>(prcp, prl, p.cst, p.inv_cst, s.hst);                          // :511
        return (qr + ql) + fmaxf(0.0f, prd + l);                                             // :513-515
    }
    return d_transfer<FAST>(prcp, pr + perc, p.cft, p.inv_cft, s.hft);                                  // gr-d :685-689
}

template <int ST, int FAST>
__global__ void __launch_bounds__(ST_WARPS * 32) vertical_struct_kernel(const __grid_constant__ CUtensorMap tm_prcp,
                                                                       const __grid_constant__ CUtensorMap tm_pet, const SplitArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    StStage *stage = reinterpret_cast<StStage *>(smem_raw) + warp * ST_NST;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + sizeof(StStage) * ST_WARPS * ST_NST) + warp * ST_NST;
    const int m = blockIdx.y;
    const int j0 = (blockIdx.x * ST_WARPS + warp) * 32;
    const int n = a.tp.n, npad = a.tp.npad, T = a.T;
    if (j0 >= n) return;
    const int j = j0 + lane;
    const bool valid = j < n;
    const int nst = (T + ST_TK - 1) / ST_TK;
    constexpr uint32_t STAGE_BYTES = sizeof(StStage);

    if (lane == 0) {
        for (int s = 0; s < ST_NST; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < ST_NST && s < nst; s++) {
            mbar_expect_tx(&bars[s], STAGE_BYTES);
            st_tma_load_2d(&stage[s][0][0][0], &tm_prcp, j0, s * ST_TK, &bars[s]);
            st_tma_load_2d(&stage[s][1][0][0], &tm_pet, j0, s * ST_TK, &bars[s]);
        }
    }
    __syncwarp();

    // padding lanes run on the reference's default values (mwd_parameters.f90:150-167, mwd_states.f90:117-126)
    SPar p{1e-6f, 200.0f, 0.005f, 500.0f, 500.0f, 0.0f, 0.3f, 100.0f, 500.0f, 2000.0f, 20.0f, 0.02f, 0.33f, 0.8f, 1e6f, 0.002f, 0.002f, 0.01f,
           0.002f, 0.0005f};
    SSto s{0.01f, 0.01f, 0.01f, 0.01f, 0.01f, 0.01f, 0.01f};
    float hlr = 0.0f, E = 0.0f;
    int fa = 1;
    if (valid) {
        const float *f = a.sfields + (size_t)m * (SMASH_B200_GNP + SMASH_B200_GNS) * npad + j;
        auto P = [&](int k) { return f[(size_t)k * npad]; };
        auto S = [&](int k) { return f[(size_t)(SMASH_B200_GNP + k) * npad]; };
        fa = a.tp.flwacc[j];
        p.ci = P(SMASH_P_CI); p.cp = P(SMASH_P_CP); p.inv_cp = 1.0f / p.cp; p.cft = P(SMASH_P_CFT); p.cst = P(SMASH_P_CST);
        p.exc = P(SMASH_P_EXC); p.b = P(SMASH_P_B); p.cusl1 = P(SMASH_P_CUSL1); p.cusl2 = P(SMASH_P_CUSL2); p.clsl = P(SMASH_P_CLSL);
        p.ks = P(SMASH_P_KS); p.ds = P(SMASH_P_DS); p.dsm = P(SMASH_P_DSM); p.ws = P(SMASH_P_WS);
        p.inv_ci = 1.0f / p.ci; p.inv_cft = 1.0f / p.cft; p.inv_cst = 1.0f / p.cst;
        p.inv_cusl1 = 1.0f / p.cusl1; p.inv_cusl2 = 1.0f / p.cusl2; p.inv_clsl = 1.0f / p.clsl;
        s.hi = S(SMASH_S_HI); s.hp = S(SMASH_S_HP); s.hft = S(SMASH_S_HFT); s.hst = S(SMASH_S_HST);
        s.husl1 = S(SMASH_S_HUSL1); s.husl2 = S(SMASH_S_HUSL2); s.hlsl = S(SMASH_S_HLSL);
        hlr = S(SMASH_S_HLR);
        E = expf(-a.dt / (P(SMASH_P_LR) * 60.0f));                         // md_routing_operator.f90:75
    }
    const bool src = fa <= 1;
    const bool all_valid = j0 + 32 <= n;
    const int gfirst = valid ? a.tp.gauge_first[j] : -1;
    const size_t qpitch = (size_t)a.qpitch;
    float *row = a.rows + ((size_t)m * npad + j) * a.Tp;
    float *qd = (a.save_q && src && valid) ? a.qdom + (size_t)m * T * qpitch + j : nullptr;
    float *np_ = (a.save_netp && valid) ? a.netp + (size_t)m * T * qpitch + j : nullptr;
    float *qsim = a.qsim + (size_t)m * T * a.tp.ng;
    const int ng = a.tp.ng;
    const float c0 = a.dx * a.dx * 0.001f / a.dt;                           // md_forward_structure.f90:155 etc.

    uint32_t parity = 0;
    int slot = 0;
#pragma unroll 1
    for (int st = 0; st < nst; st++) {
        mbar_wait(&bars[slot], parity);
        float pv[ST_TK], ev[ST_TK], qv[ST_TK];
#pragma unroll
        for (int i = 0; i < ST_TK; i++) {
            pv[i] = stage[slot][0][i][lane];
            ev[i] = stage[slot][1][i][lane];
        }
        __syncwarp();   // every lane holds its box in registers: refill the slot
        if (lane == 0 && st + ST_NST < nst) {
            mbar_expect_tx(&bars[slot], STAGE_BYTES);
            st_tma_load_2d(&stage[slot][0][0][0], &tm_prcp, j0, (st + ST_NST) * ST_TK, &bars[slot]);
            st_tma_load_2d(&stage[slot][1][0][0], &tm_pet, j0, (st + ST_NST) * ST_TK, &bars[slot]);
        }
        const int tb = st * ST_TK;
        float mn = 0.0f;
#pragma unroll
        for (int i = 0; i < ST_TK; i++) mn = fminf(mn, fminf(pv[i], ev[i]));
        if (FAST && all_valid && tb + ST_TK <= T && __all_sync(FULLM, mn >= 0.0f)) {
            // whole box, every lane a cell, no forcing gap in the warp: no per-step guards, the stores are updated in place
#pragma unroll
            for (int i = 0; i < ST_TK; i++) {
                const int t = tb + i;
                const float qt = struct_step<ST, FAST, true>(p, pv[i], ev[i], s);
                float q = qt;
                if (np_) np_[(size_t)t * qpitch] = qt;
                if (src) {
                    q = qt * c0;
                    hlr = (hlr + 0.0f) * E;
                    if (qd) qd[(size_t)t * qpitch] = q;
                    if (gfirst >= 0)
                        for (int g = gfirst; g >= 0; g = a.tp.gauge_next[g]) qsim[(size_t)t * ng + g] = q;
                }
                qv[i] = q;
            }
            st_st8(row + (size_t)tb, qv);
            if (++slot == ST_NST) { slot = 0; parity ^= 1u; }
            continue;
        }
#pragma unroll
        for (int i = 0; i < ST_TK; i++) {
            const int t = tb + i;
            SSto sn = s;
            const float qt = struct_step<ST, FAST>(p, pv[i], ev[i], sn);
            float q = qt;
            if (valid && t < T) {
                s = sn;
                if (np_) np_[(size_t)t * qpitch] = qt;
                if (src) {                                                  // final here: no inflow, q = qt dx^2 1e-3 / dt
                    q = FAST ? qt * c0 : qt * a.dx * a.dx * 0.001f / a.dt;
                    hlr = (hlr + 0.0f) * E;                                 // linear_routing with qup = 0, md_routing_operator.f90:73-77
                    if (qd) qd[(size_t)t * qpitch] = q;
                    if (gfirst >= 0)
                        for (int g = gfirst; g >= 0; g = a.tp.gauge_next[g]) qsim[(size_t)t * ng + g] = q;
                }
            }
            qv[i] = q;
        }
        if (valid) st_st8(row + (size_t)tb, qv);
        if (++slot == ST_NST) { slot = 0; parity ^= 1u; }
    }
    if (valid) {
        float *fs = a.sfstates + (size_t)m * SMASH_B200_GNS * npad + j;
        fs[(size_t)SMASH_S_HI * npad] = s.hi; fs[(size_t)SMASH_S_HP * npad] = s.hp; fs[(size_t)SMASH_S_HFT * npad] = s.hft;
        fs[(size_t)SMASH_S_HST * npad] = s.hst; fs[(size_t)SMASH_S_HUSL1 * npad] = s.husl1; fs[(size_t)SMASH_S_HUSL2 * npad] = s.husl2;
        fs[(size_t)SMASH_S_HLSL * npad] = s.hlsl;
        if (src) a.fstates[(size_t)m * 3 * npad + (size_t)2 * npad + j] = hlr;   // the routed cells' hlr comes from route_forward
    }
}

// planes [24][ncell] (16 parameters, 8 states, raster order) -> sfields [m][24][npad] in cell order j; an ensemble member takes
// its sample values (set_sample_to_parameters_states, mw_multiple_run.f90:40-65: the whole plane)
__global__ void gather_struct_fields_kernel(const int32_t *cell, int npad, int nmember, const float *planes, int64_t ncell, const float *sample,
                                            const int32_t *sample_plane, int nvar, float *sfields) {
    constexpr int NP = SMASH_B200_GNP + SMASH_B200_GNS;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)nmember * NP * npad;
    if (i >= total) return;
    const int slot = (int)(i % npad);
    const int f = (int)((i / npad) % NP);
    const int m = (int)(i / ((int64_t)npad * NP));
    const int c = cell[slot];
    float v = (c >= 0) ? planes[(int64_t)f * ncell + c] : 1.0f;
    for (int k = 0; k < nvar; k++) if (sample_plane[k] == f) v = sample[(int64_t)m * nvar + k];
    sfields[i] = v;
}

}  // namespace

cudaError_t launch_gather_struct_fields(const int32_t *cell, int npad, int nmember, const float *planes, int64_t ncell, const float *sample,
                                        const int32_t *sample_plane, int nvar, float *sfields, cudaStream_t s) {
    const int64_t total = (int64_t)nmember * (SMASH_B200_GNP + SMASH_B200_GNS) * npad;
    gather_struct_fields_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(cell, npad, nmember, planes, ncell, sample, sample_plane, nvar,
                                                                             sfields);
    return cudaGetLastError();
}

cudaError_t launch_vertical_struct(const SplitArgs &a, const CUtensorMap &prcp, const CUtensorMap &pet, int structure, int math_mode,
                                   cudaStream_t s) {
    dim3 grid((unsigned)((a.tp.n + ST_WARPS * 32 - 1) / (ST_WARPS * 32)), (unsigned)a.nmember);
    const size_t smem = sizeof(StStage) * ST_WARPS * ST_NST + sizeof(uint64_t) * ST_WARPS * ST_NST;
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, ST_WARPS * 32, smem, s>>>(prcp, pet, a);
        return cudaGetLastError();
    };
    switch (structure) {
    case SMASH_STRUCTURE_GR_B: return math_mode ? go(vertical_struct_kernel<SMASH_STRUCTURE_GR_B, 1>) : go(vertical_struct_kernel<SMASH_STRUCTURE_GR_B, 0>);
    case SMASH_STRUCTURE_GR_C: return math_mode ? go(vertical_struct_kernel<SMASH_STRUCTURE_GR_C, 1>) : go(vertical_struct_kernel<SMASH_STRUCTURE_GR_C, 0>);
    case SMASH_STRUCTURE_GR_D: return math_mode ? go(vertical_struct_kernel<SMASH_STRUCTURE_GR_D, 1>) : go(vertical_struct_kernel<SMASH_STRUCTURE_GR_D, 0>);
    case SMASH_STRUCTURE_VIC_A: return math_mode ? go(vertical_struct_kernel<SMASH_STRUCTURE_VIC_A, 1>) : go(vertical_struct_kernel<SMASH_STRUCTURE_VIC_A, 0>);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace smash
