// cell_math.cuh -- PTX helpers and the per-cell gr-a reservoir arithmetic shared by the fused and the split engine.
// Reference statements are cited as file:line under /root/reference/smash/solver/.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace smash {

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar_addr), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------
// scalar math.  FAST = 0: IEEE division / sqrt, libm tanhf.  FAST = 1: MUFU reciprocal / rsqrt.
// The reference's real exponents are all of the form +-4, +-1/4, 3.5, 2.5 (md_gr_operator.f90:62,77,96,104),
// so powf is replaced by products and square roots (<= 2 ulp, like libm's powf).
// ------------------------------------------------------------------------------------------------
template <int FAST> __device__ __forceinline__ float fdiv(float a, float b) {
    if (FAST) return __fdividef(a, b);
    return a / b;
}
template <int FAST> __device__ __forceinline__ float frcp(float a) {
    if (FAST) return __frcp_rn(a);
    return 1.0f / a;
}
template <int FAST> __device__ __forceinline__ float pow_m025(float x) {  // x ** (-0.25)
    if (FAST) return rsqrtf(sqrtf(x));
    return 1.0f / sqrtf(sqrtf(x));
}
template <int FAST> __device__ __forceinline__ float pow_m125(float x) {  // x ** (-1.25)
    return pow_m025<FAST>(x) * frcp<FAST>(x);
}
__device__ __forceinline__ float pow4(float x) { float x2 = x * x; return x2 * x2; }
template <int FAST> __device__ __forceinline__ float pow_m4(float x) { return frcp<FAST>(pow4(x)); }
__device__ __forceinline__ float pow_3p5(float x) { return (x * x) * x * sqrtf(x); }
__device__ __forceinline__ float pow_2p5(float x) { return (x * x) * sqrtf(x); }
template <int FAST> __device__ __forceinline__ float ftanh(float x) {
    if (FAST) {
        // x >= 0 here (pn, en >= 0 and cp > 0).  Odd Taylor polynomial below 0.25 (truncation < 3e-9 relative),
        // 1 - 2/(e^{2x}+1) above (absolute error ~1e-7 on a value >= 0.24).
        if (x < 0.25f) {
            const float x2 = x * x;
            float p = fmaf(x2, 0.021869488f, -0.053968254f);   // 62/2835, -17/315
            p = fmaf(x2, p, 0.13333334f);                      // 2/15
            p = fmaf(x2, p, -0.33333334f);                     // -1/3
            return fmaf(x * x2, p, x);
        }
        const float e = __expf(2.0f * x);
        return 1.0f - __fdividef(2.0f, e + 1.0f);
    }
    return tanhf(x);
}
__device__ __forceinline__ float mufu_rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fsqrt_fast(float x) { return x * mufu_rsq(x); }   // x > 0, normal range

struct CellConst {
    float cp, inv_cp, cft, inv_cft, cft_m4, exc, lr, E, fa1, den;  // den = 0.001*dx*dx*(flwacc-1)
    float s_q, c0;                                                  // dt/den (0 for sources) and dx*dx*0.001/dt
    float kr = 0.9f, kd = 0.1f;   // shares of pr + perc that go to the transfer store / to the direct branch: 0.9, 0.1 (md_forward_structure.f90:137-138);
                    // 1, 0 with exc = 0 turns every statement into gr_d_forward's (:685-689) and the reverse sweep into GR_D_FORWARD_B
};

// One gr-a cell-step WITHOUT routing (md_forward_structure.f90:106-144).  Updates hp, hft; returns qt and
// the intermediates the adjoint needs.
struct StepOut { float qt, prr, prd, l, pn, en; bool nogap; };

template <int FAST>
__device__ __forceinline__ StepOut vertical_step(const CellConst &k, float prcp, float pet, float &hp, float &hft) {
    StepOut o;
    float pr = 0.0f, perc = 0.0f, l = 0.0f;
    o.pn = 0.0f; o.en = 0.0f;
    o.nogap = (prcp >= 0.0f) && (pet >= 0.0f);                        // :106
    if (o.nogap) {
        float ei = fminf(pet, prcp);                                 // :112
        float pn = fmaxf(0.0f, prcp - ei);                           // :114
        float en = pet - ei;                                         // :116
        o.pn = pn; o.en = en;
        // gr_production md_gr_operator.f90:36-67.  One of pn, en is exactly 0, tanh(0) = 0, so only the
        // non-zero branch of (ps, es) is evaluated; the other is exactly 0.
        bool wet = pn > 0.0f;
        float th = ftanh<FAST>((wet ? pn : en) * k.inv_cp);
        float num = wet ? (k.cp * (1.0f - hp * hp)) * th : ((hp * k.cp) * (2.0f - hp)) * th;   // :52, :55
        float den = wet ? 1.0f + hp * th : 1.0f + (1.0f - hp) * th;
        float r = fdiv<FAST>(num, den);
        float hp_imd = hp + (wet ? r : -r) * k.inv_cp;               // :58
        if (wet) pr = pn - (hp_imd - hp) * k.cp;                     // :60-62
        float w = 1.0f + pow4(hp_imd * 0.001f);                      // :66 (beta = 1000, :122 of caller)
        float pw = (w == 1.0f) ? 1.0f : pow_m025<FAST>(w);
        perc = (hp_imd * k.cp) * (1.0f - pw);
        hp = hp_imd - perc * k.inv_cp;                               // :68
        // gr_exchange md_gr_operator.f90:69-79
        l = (k.exc == 0.0f) ? 0.0f : k.exc * pow_3p5(hft);
    }
    o.prr = k.kr * (pr + perc) + l;                                  // :137
    o.prd = k.kd * (pr + perc);                                      // :138
    o.l = l;
    // gr_transfer(n = 5) md_gr_operator.f90:81-110
    float pr_imd;
    if (prcp < 0.0f) {
        float x = hft * k.cft;
        pr_imd = pow_m025<0>(pow_m4<0>(x) - k.cft_m4) - x;            // :96 (forcing gap: emptying)
    } else {
        pr_imd = o.prr;
    }
    float ht_imd, ht_new, qr;
    if (FAST) {
        // ht = ((ht_imd*ct)^-4 + ct^-4)^(-1/4) / ct  ==  u / s  with u = ht_imd, s = (1 + u^4)^(1/4)   (:104)
        // q  = (ht_imd - ht)*ct = ct*u*(1 - 1/s), evaluated without the cancellation of the reference's form:
        //      1 - 1/s = z / (s (s+1) (s^2+1)),  z = u^4.
        ht_imd = fmaxf(1.e-6f, fmaf(pr_imd, k.inv_cft, hft));       // :102
        const float z = pow4(ht_imd);
        const float s2 = fsqrt_fast(1.0f + z), s1 = fsqrt_fast(s2);
        const float g = __fdividef(z, s1 * (s1 + 1.0f) * (s2 + 1.0f));
        const float rel = ht_imd * g;
        qr = rel * k.cft;                                            // :106
        ht_new = ht_imd - rel;
    } else {
        ht_imd = fmaxf(1.e-6f, hft + pr_imd / k.cft);               // :102
        float x1 = ht_imd * k.cft;
        ht_new = pow_m025<0>(pow_m4<0>(x1) + k.cft_m4) / k.cft;     // :104
        qr = (ht_imd - ht_new) * k.cft;                              // :106
    }
    hft = ht_new;
    float qd = fmaxf(0.0f, o.prd + l);                               // :142
    o.qt = qr + qd;                                                  // :144
    return o;
}

__device__ __forceinline__ CellConst make_const(float cp, float cft, float exc, float lr, int flwacc, float dt, float dx, bool grd = false) {
    CellConst k;
    if (grd) exc = 0.0f;
    k.kr = grd ? 1.0f : 0.9f; k.kd = grd ? 0.0f : 0.1f;
    k.cp = cp; k.inv_cp = 1.0f / cp;                                  // md_gr_operator.f90:47
    k.cft = cft; k.inv_cft = 1.0f / cft; k.cft_m4 = 1.0f / pow4(cft);
    k.exc = exc; k.lr = lr;
    k.E = expf(-dt / (lr * 60.0f));                                   // md_routing_operator.f90:75
    k.fa1 = (float)(flwacc - 1);
    k.den = 0.001f * dx * dx * k.fa1;                                 // md_routing_operator.f90:56
    k.s_q = (flwacc > 1) ? dt / k.den : 0.0f;
    k.c0 = dx * dx * 0.001f / dt;                                     // md_forward_structure.f90:155
    return k;
}

// Branch-light cell-step for the common case (no forcing gap in the warp), FAST math only.
// Same statements as vertical_step<1>; all lanes execute it, the caller commits the state conditionally.
__device__ __forceinline__ float vertical_step_nogap(const CellConst &k, float prcp, float pet, float &hp, float &hft) {
    const float ei = fminf(pet, prcp);                                   // md_forward_structure.f90:112
    const float pn = fmaxf(0.0f, prcp - ei);                             // :114
    const float en = pet - ei;                                           // :116
    const bool wet = pn > 0.0f;
    const float x = (wet ? pn : en) * k.inv_cp;
    float th;
    if (__all_sync(0xffffffffu, x < 0.25f)) {                            // warp-uniform: no divergence
        const float x2 = x * x;
        float p = fmaf(x2, 0.021869488f, -0.053968254f);
        p = fmaf(x2, p, 0.13333334f);
        p = fmaf(x2, p, -0.33333334f);
        th = fmaf(x * x2, p, x);
    } else {
        th = ftanh<1>(x);
    }
    const float num = (wet ? k.cp * (1.0f - hp * hp) : (hp * k.cp) * (2.0f - hp)) * th;     // md_gr_operator.f90:52,55
    const float den = fmaf(wet ? hp : 1.0f - hp, th, 1.0f);
    const float r = num * mufu_rcp(den);
    const float hp_imd = hp + (wet ? r : -r) * k.inv_cp;                 // :58
    const float pr = wet ? pn - (hp_imd - hp) * k.cp : 0.0f;             // :60-62
    float perc = 0.0f;
    if (__any_sync(0xffffffffu, hp_imd > 15.0f)) {                       // below, 1 + (hp_imd/1000)^4 == 1 in float32 (:66)
        const float w = 1.0f + pow4(hp_imd * 0.001f);
        perc = (w == 1.0f) ? 0.0f : (hp_imd * k.cp) * (1.0f - pow_m025<1>(w));
    }
    hp = hp_imd - perc * k.inv_cp;                                       // :68
    const float l = k.exc * ((hft * hft) * hft * fsqrt_fast(hft));       // md_gr_operator.f90:77
    const float prr = fmaf(k.kr, pr + perc, l);                          // md_forward_structure.f90:137
    const float prd = k.kd * (pr + perc);                                // :138
    const float u = fmaxf(1.e-6f, fmaf(prr, k.inv_cft, hft));            // md_gr_operator.f90:102
    const float z = pow4(u);
    const float s2 = fsqrt_fast(1.0f + z), s1 = fsqrt_fast(s2);
    const float g = z * mufu_rcp(s1 * (s1 + 1.0f) * (s2 + 1.0f));        // 1 - (1+u^4)^(-1/4), cancellation-free (:104)
    const float rel = u * g;
    hft = u - rel;
    return fmaf(rel, k.cft, fmaxf(0.0f, prd + l));                       // qt = qr + qd (:106, md_forward_structure.f90:142-144)
}

// Adjoint of one gr-a cell-step without routing (GR_A_FORWARD_B forward_db.f90:8128-8146, GR_TRANSFER_B :6275-6412,
// GR_EXCHANGE_B :6147-6157, GR_PRODUCTION_B :6012-6103).  Inputs: forcing, the pre-step states hp0 / hft0 and the adjoint
// of the cell-step's qt; carried backwards hp_b, hft_b; accumulated cp_b, cft_b, exc_b.
template <int FAST>
__device__ __forceinline__ void vertical_step_b(const CellConst &k, float prcp, float pet, float hp0, float hft0, float qt_b,
                                                float &hp_b, float &hft_b, float &cp_b, float &cft_b, float &exc_b) {
    // recompute the forward intermediates of this cell-step from the taped states
    const bool nogap = (prcp >= 0.0f) && (pet >= 0.0f);
    float pn = 0.0f, en = 0.0f, pr = 0.0f, perc = 0.0f, l = 0.0f;
    float tp_ = 0.0f, te_ = 0.0f, ps = 0.0f, es = 0.0f, hp_imd = hp0, pwx1 = 1.0f, pwr1 = 1.0f;
    if (nogap) {
        const float ei = fminf(pet, prcp);
        pn = fmaxf(0.0f, prcp - ei);
        en = pet - ei;
        tp_ = (pn > 0.0f) ? ftanh<FAST>(pn * k.inv_cp) : 0.0f;
        te_ = (en > 0.0f) ? ftanh<FAST>(en * k.inv_cp) : 0.0f;
        ps = fdiv<FAST>(k.cp * (1.0f - hp0 * hp0) * tp_, 1.0f + hp0 * tp_);
        es = fdiv<FAST>(hp0 * k.cp * (2.0f - hp0) * te_, 1.0f + (1.0f - hp0) * te_);
        hp_imd = hp0 + (ps - es) * k.inv_cp;
        if (pn > 0.0f) pr = pn - (hp_imd - hp0) * k.cp;
        pwx1 = 1.0f + pow4(hp_imd * 0.001f);
        pwr1 = (pwx1 == 1.0f) ? 1.0f : pow_m025<FAST>(pwx1);
        perc = hp_imd * k.cp * (1.0f - pwr1);
        l = (k.exc == 0.0f) ? 0.0f : k.exc * pow_3p5(hft0);
    }
    const float prr = k.kr * (pr + perc) + l;
    const float prd = k.kd * (pr + perc);
    const float qr_b = qt_b, qd_b = qt_b;
    float prd_b = 0.0f, l_b = 0.0f;
    if (0.0f < prd + l) { prd_b = qd_b; l_b = qd_b; }                     // :8128-8137
    // GR_TRANSFER_B(n = 5) :6275-6412
    float prr_b;
    {
        const float ct = k.cft, ht = hft0;
        float pr_imd, g_pwx1 = 0.0f, g_pwx3 = 0.0f;
        const bool gap = prcp < 0.0f;
        if (gap) {
            g_pwx1 = ht * ct;
            g_pwx3 = pow_m4<0>(g_pwx1) - k.cft_m4;
            pr_imd = pow_m025<0>(g_pwx3) - ht * ct;
        } else pr_imd = prr;
        const float hsum = ht + fdiv<FAST>(pr_imd, ct);
        const bool first = 1.e-6f < hsum;
        const float ht_imd = first ? hsum : 1.e-6f;
        const float x1 = ht_imd * ct;
        const float x3 = pow_m4<FAST>(x1) + k.cft_m4;
        const float pwr3 = pow_m025<FAST>(x3);
        const float ht_new = fdiv<FAST>(pwr3, ct);
        float htb = hft_b - ct * qr_b;
        const float pwr3_b = fdiv<FAST>(htb, ct);
        const float x3_b = -0.25f * (pwr3 * frcp<FAST>(x3)) * pwr3_b;               // pwy3*x3**(pwy3-1)
        const float x1_b = -4.0f * (pow_m4<FAST>(x1) * frcp<FAST>(x1)) * x3_b;      // pwy1*x1**(pwy1-1)
        const float ht_imd_b = ct * qr_b + ct * x1_b;
        const float ct2 = ct * ct;
        cft_b += (ht_imd - ht_new) * qr_b + (-4.0f * (k.cft_m4 * frcp<FAST>(ct))) * x3_b -
                 fdiv<FAST>(pwr3 * htb, ct2) + ht_imd * x1_b;
        float pr_imd_b;
        if (first) {
            htb = ht_imd_b;
            pr_imd_b = fdiv<FAST>(ht_imd_b, ct);
            cft_b -= fdiv<FAST>(pr_imd * ht_imd_b, ct2);
        } else { htb = 0.0f; pr_imd_b = 0.0f; }
        if (!gap) prr_b = pr_imd_b;
        else {
            const float gw3_b = (g_pwx3 <= 0.0f) ? 0.0f : -0.25f * pow_m125<0>(g_pwx3) * pr_imd_b;
            const float gw1_b = (g_pwx1 <= 0.0f) ? 0.0f : -4.0f * (pow_m4<0>(g_pwx1) / g_pwx1) * gw3_b;
            htb = htb + ct * gw1_b - ct * pr_imd_b;
            cft_b += (-4.0f * (k.cft_m4 / ct)) * (-gw3_b) - ht * pr_imd_b + ht * gw1_b;
            prr_b = 0.0f;
        }
        hft_b = htb;
    }
    const float pr_b = k.kd * prd_b + k.kr * prr_b;                       // :8143
    float perc_b = pr_b;
    l_b += prr_b;
    if (nogap) {
        // GR_EXCHANGE_B :6147-6157 (pre-transfer hft)
        if (k.exc != 0.0f || l_b != 0.0f) {
            exc_b += pow_3p5(hft0) * l_b;
            hft_b += 3.5f * pow_2p5(hft0) * k.exc * l_b;
        }
        // GR_PRODUCTION_B :6012-6103
        const float cp = k.cp, inv_cp = k.inv_cp, hp = hp0;
        perc_b = perc_b - inv_cp * hp_b;
        float inv_cp_b = -(perc * hp_b);
        cp_b += hp_imd * (1.0f - pwr1) * perc_b;
        const float pwr1_b = -(hp_imd * cp * perc_b);
        const float pwx1_b = (pwx1 == 1.0f) ? -(0.25f * pwr1_b) : -(0.25f * pow_m125<FAST>(pwx1) * pwr1_b);
        float hp_imd_b = hp_b + cp * (1.0f - pwr1) * perc_b + 4.0f * (hp_imd * hp_imd * hp_imd) * pwx1_b * 1.0e-12f;
        float hpb;
        if (pn > 0.0f) {
            hp_imd_b -= cp * pr_b;
            hpb = cp * pr_b;
            cp_b -= (hp_imd - hp) * pr_b;
        } else hpb = 0.0f;
        const float es_b = -(inv_cp * hp_imd_b);
        const float temp3 = (1.0f - hp) * te_ + 1.0f;
        const float temp0e = hp * cp * (2.0f - hp);
        const float temp_b3 = fdiv<FAST>(es_b, temp3);
        const float temp_be = (2.0f - hp) * te_ * temp_b3;
        const float temp_b0e = -fdiv<FAST>(temp0e * te_ * temp_b3, temp3);
        hpb = hpb + hp_imd_b + cp * temp_be - hp * cp * te_ * temp_b3 - te_ * temp_b0e;
        const float ps_b = inv_cp * hp_imd_b;
        const float sech_e = 1.0f - te_ * te_;
        const float temp_b4 = sech_e * temp0e * temp_b3;
        const float temp_b5 = sech_e * (1.0f - hp) * temp_b0e;
        cp_b += hp * temp_be;
        const float temp0p = hp * tp_ + 1.0f;
        const float temp2 = cp * (1.0f - hp * hp);
        const float temp_bp = fdiv<FAST>(ps_b, temp0p);
        const float sech_p = 1.0f - tp_ * tp_;
        const float temp_b0p = sech_p * temp2 * temp_bp;
        const float temp_b1 = -fdiv<FAST>(temp2 * tp_ * temp_bp, temp0p);
        hpb = hpb + tp_ * temp_b1 - 2.0f * hp * cp * tp_ * temp_bp;
        const float temp_b2 = sech_p * hp * temp_b1;
        inv_cp_b = inv_cp_b + (ps - es) * hp_imd_b + en * temp_b5 + en * temp_b4 + pn * temp_b2 + pn * temp_b0p;
        cp_b += (1.0f - hp * hp) * tp_ * temp_bp - fdiv<FAST>(inv_cp_b, cp * cp);
        hp_b = hpb;
    }
}

}  // namespace smash
