// split_kernels.cu -- the split engine: reservoirs and routing of gr_a_forward as separate passes (split_kernels.cuh).
//
// Reference statements are cited as file:line under /root/reference/smash/solver/.
#include "split_kernels.cuh"

#include <algorithm>

#include "cell_math.cuh"

namespace smash {

constexpr unsigned FULL = 0xffffffffu;
constexpr int VT_TK = 8;      // time steps per TMA box (= one 32-byte sector of a row)
constexpr int VT_WARPS = 8;   // warps per CTA of the per-cell passes; every warp runs its own pipeline, no CTA barrier
constexpr int VF_NST = 3;     // forward: boxes in flight per warp and array
constexpr int VB_NST = 2;     // adjoint: boxes in flight per warp and array

// 2-D TMA tile load global -> shared (SASS UTMALDG), completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int x, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
// 32-byte (one sector) vector accesses, sm_100+
__device__ __forceinline__ void ld8(const float *p, float *v) {
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st8(float *p, const float *v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int FAST> __device__ __forceinline__ float scale_q(float v, float c0, float dx, float dt) {
    return FAST ? v * c0 : v * dx * dx * 0.001f / dt;                     // md_forward_structure.f90:155
}

// gr-a cell-step for a stage whose forcing has no gap, whose tanh arguments stay below 0.25 and whose production store
// cannot reach the percolation threshold of float32 (hp_imd <= 15 => 1 + (hp_imd/1000)^4 == 1): the statements of
// vertical_step_nogap (cell_math.cuh) with those warp-uniform branches resolved once per stage.  Bit-identical results.
template <bool EXC>
__device__ __forceinline__ float vertical_step_lean(const CellConst &k, float prcp, float pet, float &hp, float &hft) {
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
    const float prr = fmaf(k.kr, pr, l);                                 // md_forward_structure.f90:137
    const float prd = k.kd * pr;                                         // :138
    const float u = fmaxf(1.e-6f, fmaf(prr, k.inv_cft, hft));            // md_gr_operator.f90:102
    const float z = pow4(u);
    const float s2 = fsqrt_fast(1.0f + z), s1 = fsqrt_fast(s2);
    const float g = z * mufu_rcp(s1 * (s1 + 1.0f) * (s2 + 1.0f));        // 1 - (1+u^4)^(-1/4), cancellation-free (:104)
    const float rel = u * g;
    hft = u - rel;
    return fmaf(rel, k.cft, fmaxf(0.0f, prd + l));                       // qt = qr + qd (:106, md_forward_structure.f90:142-144)
}

// Adjoint of vertical_step_lean (same stage-level preconditions: no gap, tanh arguments below 0.25, hp_imd <= 15 so that
// perc == 0 and pwx1 == 1): the statements of vertical_step_b (cell_math.cuh) with exactly one of the production /
// evaporation branches live (pn and en are never both positive), written with selects instead of branches and with the
// per-cell reciprocals hoisted.  GR_TRANSFER_B forward_db.f90:6275-6412, GR_EXCHANGE_B :6147-6157, GR_PRODUCTION_B :6012-6103.
template <bool EXC>
__device__ __forceinline__ void vertical_step_b_lean(const CellConst &k, float inv_cft2, float inv_cp2, float prcp, float pet, float hp,
                                                     float ht, float qt_b, float &hp_b, float &hft_b, float &cp_b, float &cft_b,
                                                     float &exc_b) {
    // forward intermediates from the taped states
    const float ei = fminf(pet, prcp);
    const float pn = fmaxf(0.0f, prcp - ei);
    const float en = pet - ei;
    const bool wet = pn > 0.0f;
    const float am = wet ? pn : en;
    const float x = am * k.inv_cp;
    const float x2 = x * x;
    float p = fmaf(x2, 0.021869488f, -0.053968254f);
    p = fmaf(x2, p, 0.13333334f);
    p = fmaf(x2, p, -0.33333334f);
    const float th = fmaf(x * x2, p, x);
    const float u1 = wet ? hp : 1.0f - hp;
    const float Nn = wet ? 1.0f - hp * hp : hp * (2.0f - hp);
    const float N = k.cp * Nn;
    const float inv_den = mufu_rcp(fmaf(u1, th, 1.0f));
    const float val = N * th * inv_den;                               // ps (wet) or es (dry), md_gr_operator.f90:52,55
    const float sval = wet ? val : -val;
    const float hp_imd = fmaf(sval, k.inv_cp, hp);                    // :58
    const float pr = wet ? pn - (hp_imd - hp) * k.cp : 0.0f;          // :60-62
    const float h25 = (ht * ht) * fsqrt_fast(ht);                     // hft^2.5
    const float l = EXC ? k.exc * (h25 * ht) : 0.0f;                  // :77
    const float prr = fmaf(k.kr, pr, l);
    const float prd = k.kd * pr;
    // reverse
    const float qr_b = qt_b;
    const bool qd_on = 0.0f < prd + l;                                // forward_db.f90:8128-8137
    const float prd_b = qd_on ? qt_b : 0.0f;
    float l_b = prd_b;
    // GR_TRANSFER_B (n = 5), no gap
    const float ct = k.cft, ict = k.inv_cft;
    const float hsum = fmaf(prr, ict, ht);
    const bool first = 1.e-6f < hsum;
    const float ht_imd = first ? hsum : 1.e-6f;
    const float x1 = ht_imd * ct;
    const float ix1 = mufu_rcp(x1);
    const float x1m4 = pow4(ix1);
    const float x3 = x1m4 + k.cft_m4;
    const float pwr3 = mufu_rsq(fsqrt_fast(x3));
    const float ht_new = pwr3 * ict;
    const float htb = hft_b - ct * qr_b;
    const float pwr3_b = htb * ict;
    const float x3_b = -0.25f * (pwr3 * mufu_rcp(x3)) * pwr3_b;
    const float x1_b = -4.0f * (x1m4 * ix1) * x3_b;
    const float ht_imd_b = ct * qr_b + ct * x1_b;
    cft_b += (ht_imd - ht_new) * qr_b + (-4.0f * (k.cft_m4 * ict)) * x3_b - (pwr3 * htb) * inv_cft2 + ht_imd * x1_b;
    const float prr_b = first ? ht_imd_b * ict : 0.0f;
    if (first) cft_b -= (prr * ht_imd_b) * inv_cft2;
    hft_b = first ? ht_imd_b : 0.0f;
    const float pr_b = fmaf(k.kr, prr_b, k.kd * prd_b);               // :8143
    l_b += prr_b;
    exc_b = fmaf(h25 * ht, l_b, exc_b);                               // GR_EXCHANGE_B (pre-transfer hft); exc_b is fed even where exc == 0
    if (EXC) hft_b = fmaf(3.5f * h25 * k.exc, l_b, hft_b);
    // GR_PRODUCTION_B with perc == 0, pwx1 == 1
    const float perc_b = pr_b - k.inv_cp * hp_b;
    const float pwx1_b = 0.25f * (hp_imd * k.cp * perc_b);
    float hp_imd_b = fmaf(4.0f * (hp_imd * hp_imd * hp_imd) * pwx1_b, 1.0e-12f, hp_b);
    float hpb = 0.0f;
    if (wet) {
        hp_imd_b -= k.cp * pr_b;
        hpb = k.cp * pr_b;
        cp_b -= (hp_imd - hp) * pr_b;
    }
    const float gb = (wet ? k.inv_cp : -k.inv_cp) * hp_imd_b;         // ps_b or es_b
    const float tb = gb * inv_den;
    const float dden = -(N * th * tb) * inv_den;
    const float dNdhp = wet ? -2.0f * hp * k.cp : k.cp * (2.0f - 2.0f * hp);
    hpb = hpb + hp_imd_b + dNdhp * th * tb + (wet ? th : -th) * dden;
    const float sech = 1.0f - th * th;
    const float d_th = sech * N * tb, d_u = sech * u1 * dden;
    const float inv_cp_b = sval * hp_imd_b + am * d_u + am * d_th;
    cp_b += Nn * th * tb - inv_cp_b * inv_cp2;
    hp_b = hpb;
}

// ------------------------------------------------------------------------------------------------
// vertical_forward: md_forward_structure.f90:106-144 for every cell and time step; no inter-cell dependency.
// Source cells (flwacc == 1) are final here: q = qt * dx^2 * 1e-3 / dt (:155 with flwacc - 1 = 0).
// ------------------------------------------------------------------------------------------------
typedef float FwdStage[2][VT_TK][32];

template <int FAST, int TAPE>
__global__ void __launch_bounds__(VT_WARPS * 32) vertical_forward_kernel(const __grid_constant__ CUtensorMap tm_prcp,
                                                                        const __grid_constant__ CUtensorMap tm_pet,
                                                                        const SplitArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FwdStage *stage = reinterpret_cast<FwdStage *>(smem_raw) + warp * VF_NST;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + sizeof(FwdStage) * VT_WARPS * VF_NST) + warp * VF_NST;
    const int m = blockIdx.y;
    const int j0 = (blockIdx.x * VT_WARPS + warp) * 32;
    const int n = a.tp.n, npad = a.tp.npad, T = a.T;
    if (j0 >= n) return;
    const int j = j0 + lane;
    const bool valid = j < n;
    const int t_begin = a.t_begin, t_end = a.t_end;                       // whole run: 0, T; streamed runs: one window
    const int nst = (t_end - t_begin + VT_TK - 1) / VT_TK;
    constexpr uint32_t STAGE_BYTES = sizeof(FwdStage);

    if (lane == 0) {
        for (int s = 0; s < VF_NST; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < VF_NST && s < nst; s++) {
            mbar_expect_tx(&bars[s], STAGE_BYTES);
            tma_load_2d(&stage[s][0][0][0], &tm_prcp, j0, t_begin + s * VT_TK, &bars[s]);
            tma_load_2d(&stage[s][1][0][0], &tm_pet, j0, t_begin + s * VT_TK, &bars[s]);
        }
    }
    __syncwarp();

    const float *fld = a.fields + (size_t)m * NFIELD * npad + j;
    float hp = 0.01f, hft = 0.01f, hlr = 0.0f;
    int fa = 1;
    CellConst k = make_const(200.0f, 500.0f, 0.0f, 5.0f, 1, a.dt, a.dx);
    if (valid) {
        fa = a.tp.flwacc[j];
        k = make_const(fld[(size_t)F_CP * npad], fld[(size_t)F_CFT * npad], fld[(size_t)F_EXC * npad], fld[(size_t)F_LR * npad], fa,
                       a.dt, a.dx, a.grd != 0);
        hp = fld[(size_t)F_HP * npad]; hft = fld[(size_t)F_HFT * npad]; hlr = fld[(size_t)F_HLR * npad];
        if (t_begin > 0) {                                               // a later window: the states the previous one left
            const float *fs = a.fstates + (size_t)m * 3 * npad + j;
            hp = fs[0]; hft = fs[(size_t)npad];
            if (fa <= 1) hlr = fs[(size_t)2 * npad];
        }
    }
    const bool src = fa <= 1;
    const int gfirst = valid ? a.tp.gauge_first[j] : -1;
    const bool all_valid = j0 + 32 <= n;
    const bool exc_on = __any_sync(FULL, k.exc != 0.0f);
    const size_t qpitch = (size_t)a.qpitch;
    float *row = a.rows + ((size_t)m * npad + j) * a.Tp;
    float *qd = (a.save_q && src && valid) ? a.qdom + (size_t)m * T * qpitch + j : nullptr;
    float *np_ = (a.save_netp && valid) ? a.netp + (size_t)m * T * qpitch + j : nullptr;
    const long long tape_row0 = (long long)m * T - a.tape_t0;             // tape row of time step 0 (checkpointed runs tape one window)
    float *thp = TAPE ? a.tape_hp + tape_row0 * npad + j : nullptr;
    float *thft = TAPE ? a.tape_hft + tape_row0 * npad + j : nullptr;
    float *qsim = a.qsim + (size_t)m * T * a.tp.ng;
    const int ng = a.tp.ng;
    const float c0 = k.c0, E = k.E, inv_cp = k.inv_cp;

    uint32_t parity = 0;
    int slot = 0;
#pragma unroll 1
    for (int st = 0; st < nst; st++) {
        mbar_wait(&bars[slot], parity);
        float pv[VT_TK], ev[VT_TK], qv[VT_TK];
        float mn = 0.0f, mx = 0.0f;
#pragma unroll
        for (int i = 0; i < VT_TK; i++) {
            pv[i] = stage[slot][0][i][lane];
            ev[i] = stage[slot][1][i][lane];
            mn = fminf(mn, fminf(pv[i], ev[i]));
            mx = fmaxf(mx, fmaxf(pv[i], ev[i]));
        }
        __syncwarp();   // every lane holds its stage in registers: refill the slot
        if (lane == 0 && st + VF_NST < nst) {
            mbar_expect_tx(&bars[slot], STAGE_BYTES);
            tma_load_2d(&stage[slot][0][0][0], &tm_prcp, j0, t_begin + (st + VF_NST) * VT_TK, &bars[slot]);
            tma_load_2d(&stage[slot][1][0][0], &tm_pet, j0, t_begin + (st + VF_NST) * VT_TK, &bars[slot]);
        }
        const int tb = t_begin + st * VT_TK;
        const bool full = all_valid && tb + VT_TK <= t_end;
        const float xm = mx * inv_cp;
        const bool lean = FAST && full && __all_sync(FULL, mn >= 0.0f && xm < 0.25f && fmaf(8.0f, xm, hp) < 15.0f);
        if (lean) {
#pragma unroll
            for (int i = 0; i < VT_TK; i++) {
                if (TAPE) { thp[(size_t)(tb + i) * npad] = hp; thft[(size_t)(tb + i) * npad] = hft; }
                const float qt = exc_on ? vertical_step_lean<true>(k, pv[i], ev[i], hp, hft) : vertical_step_lean<false>(k, pv[i], ev[i], hp, hft);
                if (np_) np_[(size_t)(tb + i) * qpitch] = qt;
                float q = qt;
                if (src) {
                    q = qt * c0;
                    hlr = (hlr + 0.0f) * E;                                // linear_routing with qup = 0, md_routing_operator.f90:73-77
                    if (qd) qd[(size_t)(tb + i) * qpitch] = q;
                    if (gfirst >= 0)
                        for (int g = gfirst; g >= 0; g = a.tp.gauge_next[g]) qsim[(size_t)(tb + i) * ng + g] = q;   // :206-210
                }
                qv[i] = q;
            }
        } else {
#pragma unroll
            for (int i = 0; i < VT_TK; i++) {
                const int t = tb + i;
                const bool act = valid && t < t_end;
                if (TAPE && act) { thp[(size_t)t * npad] = hp; thft[(size_t)t * npad] = hft; }
                float hp_n = hp, hft_n = hft, qt;
                const bool gapless = (pv[i] >= 0.0f) && (ev[i] >= 0.0f);
                if (FAST && __all_sync(FULL, gapless)) qt = vertical_step_nogap(k, pv[i], ev[i], hp_n, hft_n);
                else qt = vertical_step<FAST>(k, pv[i], ev[i], hp_n, hft_n).qt;
                float q = qt;
                if (act) {
                    hp = hp_n; hft = hft_n;
                    if (np_) np_[(size_t)t * qpitch] = qt;
                    if (src) {
                        q = scale_q<FAST>(qt, c0, a.dx, a.dt);
                        hlr = (hlr + 0.0f) * E;
                        if (qd) qd[(size_t)t * qpitch] = q;
                        if (gfirst >= 0)
                            for (int g = gfirst; g >= 0; g = a.tp.gauge_next[g]) qsim[(size_t)t * ng + g] = q;
                    }
                }
                qv[i] = q;
            }
        }
        if (valid) st8(row + (size_t)tb, qv);
        if (++slot == VF_NST) { slot = 0; parity ^= 1u; }
    }
    if (valid) {
        float *fs = a.fstates + (size_t)m * 3 * npad + j;
        fs[0] = hp; fs[(size_t)npad] = hft;
        if (src) fs[(size_t)2 * npad] = hlr;
    }
}

// ------------------------------------------------------------------------------------------------
// row helpers: a lane owns S consecutive time steps of a row
// ------------------------------------------------------------------------------------------------
template <int S> __device__ __forceinline__ void ld_row(const float *p, float (&v)[S]) {
#pragma unroll
    for (int i = 0; i < S / 8; i++) ld8(p + 8 * i, &v[8 * i]);
}
template <int S> __device__ __forceinline__ void st_row(float *p, const float (&v)[S]) {
#pragma unroll
    for (int i = 0; i < S / 8; i++) st8(p + 8 * i, &v[8 * i]);
}
// every lane asks for the 128-byte lines of its own row (window part): DRAM -> L2 ahead of use
template <int S> __device__ __forceinline__ void prefetch_row(const float *row_window) {
#pragma unroll
    for (int i = 0; i < S; i++) prefetch_l2(row_window + 32 * i);
}
__device__ __forceinline__ void wait_flag(const int *flag, int epoch, int lane) {
    if (lane == 0)
        while (ld_acquire(flag) < epoch) __nanosleep(32);
    __syncwarp();
}
// the row stores of every lane happen before the barrier, the release store of lane 0 after it: cumulativity makes
// them visible to whoever acquires the flag (no separate fence.sc, which costs microseconds here)
__device__ __forceinline__ void publish_flag(int *flag, int value, int lane) {
    __syncwarp();
    if (lane == 0) st_release(flag, value);
}
__device__ __forceinline__ int claim_ticket(unsigned int *ticket, int lane) {
    int tk = 0;
    if (lane == 0) tk = (int)atomicAdd(ticket, 1u);
    return __shfl_sync(FULL, tk, 0);
}

struct RouteConst { float fa1, s_q, E, c0, lr, h0; };
__device__ __forceinline__ RouteConst route_const(const SplitArgs &a, int m, int j, const float *carry) {
    RouteConst c;
    const int fa = a.tp.flwacc[j];
    c.lr = a.fields[((size_t)m * NFIELD + F_LR) * a.tp.npad + j];
    c.h0 = carry[(size_t)m * a.tp.npad + j];
    c.fa1 = (float)(fa - 1);
    c.s_q = (fa > 1) ? a.dt / (0.001f * a.dx * a.dx * c.fa1) : 0.0f;      // md_routing_operator.f90:55-56
    c.E = expf(-a.dt / (c.lr * 60.0f));                                   // :75
    c.c0 = a.dx * a.dx * 0.001f / a.dt;                                   // md_forward_structure.f90:155
    return c;
}
__device__ __forceinline__ RouteConst shfl_const(const RouteConst &c, int src) {
    RouteConst r;
    r.fa1 = __shfl_sync(FULL, c.fa1, src); r.s_q = __shfl_sync(FULL, c.s_q, src); r.E = __shfl_sync(FULL, c.E, src);
    r.c0 = __shfl_sync(FULL, c.c0, src); r.lr = __shfl_sync(FULL, c.lr, src); r.h0 = __shfl_sync(FULL, c.h0, src);
    return r;
}

// ------------------------------------------------------------------------------------------------
// route_forward: upstream_discharge + linear_routing (md_routing_operator.f90:17-79) + the discharge of the
// cell (md_forward_structure.f90:155) on whole time windows.  Per cell the routing state obeys
//   hlr(t) = (hlr(t-1) + qup(t)) * E,   qrout = hr_imd - hlr,   q = (qt + qrout * (flwacc - 1)) * dx^2 * 1e-3 / dt
// which is linear in time: every lane runs its S steps from a zero state, a warp scan of the affine maps
// (E^S, end value) gives every lane its true carry-in, a second pass produces the results.
// ------------------------------------------------------------------------------------------------
template <int S, int TAPE>
__device__ __forceinline__ void route_cell(const SplitArgs &a, const RouteConst &c, int m, int j, int w, int lane, int t_first,
                                           bool gauge, float (&x)[S], const float (&qt)[S], float (&r)[S]) {
    const float E = c.E;
    const float h0 = c.h0;
    float h = (lane == 0) ? h0 : 0.0f;
    float A = 1.0f;
#pragma unroll
    for (int s = 0; s < S; s++) {
        x[s] = x[s] * c.s_q;
        h = (h + x[s]) * E;
        A *= E;
    }
    float Bv = h;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {                     // every lane's map has the same slope E^S: slope of d lanes = A
        const float Bo = __shfl_up_sync(FULL, Bv, d);
        if (lane >= d) Bv = fmaf(A, Bo, Bv);
        A *= A;
    }
    h = __shfl_up_sync(FULL, Bv, 1);
    if (lane == 0) h = h0;
    const int tcap = min(a.T, (w + 1) * a.W) - 1;          // last valid step of this window: its state is carried on
    float hcap = 0.0f;
#pragma unroll
    for (int s = 0; s < S; s++) {
        const float hr = h + x[s];                          // :73
        const float hn = hr * E;                            // :75
        const float qrout = hr - hn;                        // :77
        r[s] = fmaf(qrout, c.fa1, qt[s]) * c.c0;            // md_forward_structure.f90:155
        x[s] = hr;
        if (t_first + s == tcap) hcap = hn;
        h = hn;
    }
    const float hfin = __shfl_sync(FULL, hcap, (tcap - w * a.W) / S);
    float *rows = a.rows + ((size_t)m * a.tp.npad + j) * a.Tp + t_first;
    st_row<S>(rows, r);
    if (TAPE) st_row<S>(a.rows_hr + ((size_t)m * a.tp.npad + j) * a.Tp + t_first, x);
    if (lane == 0) {
        a.hcar[(size_t)m * a.tp.npad + j] = hfin;
        if (w == a.nwin - 1) a.fstates[((size_t)m * 3 + 2) * a.tp.npad + j] = hfin;
    }
    if (gauge) {
        float *qsim = a.qsim + (size_t)m * a.T * a.tp.ng;
        for (int g = a.tp.gauge_first[j]; g >= 0; g = a.tp.gauge_next[g])
#pragma unroll
            for (int s = 0; s < S; s++)
                if (t_first + s < a.T) qsim[(size_t)(t_first + s) * a.tp.ng + g] = r[s];          // :206-210
    }
}

// inflows of a pit-pair cell other than its partner, summed in the reference's order
template <int S>
__device__ __forceinline__ void gather_laterals(const SplitArgs &a, const int *done, int epoch, const float *rows_lane,
                                                const int4 rec, int lane, float (&lat)[S]) {
#pragma unroll
    for (int s = 0; s < S; s++) lat[s] = 0.0f;
    const int nup = rec.y >> 8;
#pragma unroll 1
    for (int e = 0; e < nup; e++) {
        const int2 u = a.tp.tup[rec.z + e];
        if (u.y <= UP_HEAVY) continue;
        if (u.y >= 0) wait_flag(done + u.y, epoch, lane);
        float v[S];
        ld_row<S>(rows_lane + (size_t)u.x * a.Tp, v);
#pragma unroll
        for (int s = 0; s < S; s++) lat[s] = lat[s] + v[s];
    }
}

// pit pair (A earlier in path, B its partner): A gathers q_B of the previous step, B gathers q_A of the current
// step.  The coupled recurrence is run step by step, lane after lane.
template <int S, int TAPE>
__device__ __noinline__ void route_pair(const SplitArgs &a, int m, const int4 recA, const int4 recB, int w, int lane, int t_first,
                                        const int *done, int epoch) {
    const size_t npad = a.tp.npad;
    const int jA = recA.x, jB = recB.x;
    const float *rows_lane = a.rows + (size_t)m * npad * a.Tp + t_first;
    float latA[S], latB[S], qA[S], qB[S];
    gather_laterals<S>(a, done, epoch, rows_lane, recA, lane, latA);
    gather_laterals<S>(a, done, epoch, rows_lane, recB, lane, latB);
    ld_row<S>(rows_lane + (size_t)jA * a.Tp, qA);
    ld_row<S>(rows_lane + (size_t)jB * a.Tp, qB);
    const RouteConst cA = route_const(a, m, jA, a.hcar), cB = route_const(a, m, jB, a.hcar);
    float hA = cA.h0, hB = cB.h0;
    float qBp = 0.0f;                                                     // q_B of the step before the window
    if (w > 0) qBp = a.qprev ? a.qprev[(size_t)m * npad + jB] : a.rows[((size_t)m * npad + jB) * a.Tp + (size_t)w * a.W - 1];
#pragma unroll 1
    for (int L = 0; L < 32; L++) {
        float sA = hA, sB = hB, sq = qBp;
        if (lane == L) {
#pragma unroll
            for (int s = 0; s < S; s++) {
                const float hrA = sA + (latA[s] + sq) * cA.s_q;
                const float hnA = hrA * cA.E;
                const float qa = fmaf(hrA - hnA, cA.fa1, qA[s]) * cA.c0;
                const float hrB = sB + (latB[s] + qa) * cB.s_q;
                const float hnB = hrB * cB.E;
                const float qb = fmaf(hrB - hnB, cB.fa1, qB[s]) * cB.c0;
                qA[s] = qa; qB[s] = qb; latA[s] = hrA; latB[s] = hrB;
                if (t_first + s < a.T) { sA = hnA; sB = hnB; sq = qb; }
            }
        }
        hA = __shfl_sync(FULL, sA, L); hB = __shfl_sync(FULL, sB, L); qBp = __shfl_sync(FULL, sq, L);
    }
    float *rows = a.rows + (size_t)m * npad * a.Tp + t_first;
    st_row<S>(rows + (size_t)jA * a.Tp, qA);
    st_row<S>(rows + (size_t)jB * a.Tp, qB);
    if (a.fuse_export) {                                                 // the routing warps leave the pit cells to their own kernel
        float *qd = a.qdom + (size_t)m * a.T * a.qpitch;
#pragma unroll
        for (int s = 0; s < S; s++)
            if (t_first + s < a.T) {
                qd[(size_t)(t_first + s) * a.qpitch + jA] = qA[s];
                qd[(size_t)(t_first + s) * a.qpitch + jB] = qB[s];
            }
    }
    if (TAPE) {
        float *hr = a.rows_hr + (size_t)m * npad * a.Tp + t_first;
        st_row<S>(hr + (size_t)jA * a.Tp, latA);
        st_row<S>(hr + (size_t)jB * a.Tp, latB);
    }
    if (lane == 0) {
        a.hcar[(size_t)m * npad + jA] = hA; a.hcar[(size_t)m * npad + jB] = hB;
        if (w == a.nwin - 1) {
            a.fstates[((size_t)m * 3 + 2) * npad + jA] = hA;
            a.fstates[((size_t)m * 3 + 2) * npad + jB] = hB;
        }
    }
    float *qsim = a.qsim + (size_t)m * a.T * a.tp.ng;
    for (int g = a.tp.gauge_first[jA]; g >= 0; g = a.tp.gauge_next[g])
#pragma unroll
        for (int s = 0; s < S; s++)
            if (t_first + s < a.T) qsim[(size_t)(t_first + s) * a.tp.ng + g] = qA[s];
    for (int g = a.tp.gauge_first[jB]; g >= 0; g = a.tp.gauge_next[g])
#pragma unroll
        for (int s = 0; s < S; s++)
            if (t_first + s < a.T) qsim[(size_t)(t_first + s) * a.tp.ng + g] = qB[s];
}

// One warp per task.  The records of up to 32 cells of the chain are fetched with one coalesced load, lane i then
// owns the scalars of cell i (constants, carried state) and one inflow entry of the group; the tributaries of the whole
// group are awaited once, lane-parallel, and their rows requested from DRAM, before the warp routes the cells one after
// the other.
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
template <int N> __device__ __forceinline__ void ld_vec(const float *p, float (&v)[N]) {      // N even, p 8-byte aligned
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        const float2 x = __ldcg(reinterpret_cast<const float2 *>(p) + i);
        v[2 * i] = x.x; v[2 * i + 1] = x.y;
    }
}
template <int N> __device__ __forceinline__ void st_vec(float *p, const float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) reinterpret_cast<float2 *>(p)[i] = make_float2(v[2 * i], v[2 * i + 1]);
}
template <int N> __device__ __forceinline__ void lds_vec(const float *p, float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        const float2 x = reinterpret_cast<const float2 *>(p)[i];
        v[2 * i] = x.x; v[2 * i + 1] = x.y;
    }
}

// Staging area of one warp: the next cell's own row and its first two tributary rows travel L2 -> shared memory
// (cp.async, no registers) while the current cell is routed.  [slot][float4 index][lane]: every lane reads and writes
// only its own column, so no warp barrier is involved.
template <int S> struct RowStage { float4 v[3][S / 4][32]; };
template <int S> __device__ __forceinline__ void stage_row(RowStage<S> &st, int slot, int lane, const float *src_lane) {
#pragma unroll
    for (int i = 0; i < S / 4; i++) cp_async16(&st.v[slot][i][lane], src_lane + 4 * i);
}
template <int S> __device__ __forceinline__ void staged_row(const RowStage<S> &st, int slot, int lane, float (&v)[S]) {
#pragma unroll
    for (int i = 0; i < S / 4; i++) {
        const float4 x = st.v[slot][i][lane];
        v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
    }
}

// ---- dynamic scheduling of the ticketed chains -------------------------------------------------------------------
// A finished task tells the ticketed chain that gathers its last cell; the chain whose last tributary chain has finished
// is pushed to the ready queue of its class.  Called by one thread after the task's done flag has been released.
__device__ __forceinline__ void notify_consumer(const SplitArgs &a, int task) {
    if (!a.dyn) return;
    const int c = a.cons[task];
    if (c < 0) return;
    __threadfence();                                         // release: this task's rows before the decrement
    if (atomicSub(a.ndep + c, 1) == 1) {
        __threadfence();                                     // acquire: the other tributaries' decrements (and rows) before the push
        const int q = a.qid[c];
        const unsigned s = atomicAdd(a.qctl + 16 + q, 1u);
        st_release(a.queue + a.qoff[q] + (int)s, c);
    }
}
// Next ready chain, queues in priority order; -1 when every chain has been handed out.  A warp never holds a task while
// it waits here, and a chain in a queue has no unfinished tributary: nothing ever spins on a chain that has not started.
// block = false: returns -2 instead of waiting when no chain is ready at the moment (the caller has other work).
__device__ __forceinline__ int pop_ready(const SplitArgs &a, int lane, bool block = true) {
    for (;;) {
        unsigned h = 0, t = 0, cap = 0;
        if (lane < a.dyn_nq) {
            h = *reinterpret_cast<volatile unsigned int *>(a.qctl + lane);
            t = *reinterpret_cast<volatile unsigned int *>(a.qctl + 16 + lane);
            cap = (unsigned)(a.qoff[lane + 1] - a.qoff[lane]);
        }
        const unsigned alive = __ballot_sync(FULL, h < cap);
        if (!alive) return -1;
        const unsigned avail = __ballot_sync(FULL, h < t && h < cap);
        if (!avail) {
            if (!block) return -2;
            __nanosleep(256);
            continue;
        }
        const int q = __ffs(avail) - 1;
        unsigned s = 0;
        if (lane == 0) s = atomicAdd(a.qctl + q, 1u);
        s = __shfl_sync(FULL, s, 0);
        const unsigned capq = __shfl_sync(FULL, cap, q);
        if (s >= capq) continue;                             // the queue ran out between the look and the claim
        int task = -1;
        if (lane == 0)
            while ((task = ld_acquire(a.queue + a.qoff[q] + (int)s)) < 0) __nanosleep(64);   // claimed ahead of the push
        return __shfl_sync(FULL, task, 0);
    }
}

template <int S, int TAPE>
__device__ __forceinline__ void route_chain_warp(const SplitArgs &a, RowStage<S> &stg, int m, int task, int w, int lane, int t_first,
                                            int epoch) {
    const SplitTopo &tp = a.tp;
    const int cb = tp.task_begin[task], ce = tp.task_begin[task + 1];
    int *done = a.done + (size_t)m * tp.ntask;
    const float *rows_lane = a.rows + (size_t)m * tp.npad * a.Tp + t_first;
    const float *rows_win = a.rows + (size_t)m * tp.npad * a.Tp + (size_t)w * a.W;
    float r[S];
#pragma unroll
    for (int s = 0; s < S; s++) r[s] = 0.0f;
    int ngr = 0;
    const bool prof = a.dbg_prof != nullptr && task >= tp.nchain - tp.nded;
    long long t_start = 0, t_wait = 0;
    if (prof) t_start = clock64();
#pragma unroll 1
    for (int g0 = cb; g0 < ce; g0 += ngr) {
        ngr = min(32, ce - g0);
        int4 rec = make_int4(-1, 0, 0, 0);
        if (lane < ngr) rec = tp.tcell[g0 + lane];
        {   // the group ends where its inflow entries would no longer fit the 32 lanes that hold them
            int cum = (lane < ngr) ? (rec.y >> 8) : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(FULL, cum, d);
                if (lane >= d) cum += o;
            }
            ngr = max(1, __popc(__ballot_sync(FULL, lane < ngr && cum <= 32)));
            if (lane >= ngr) rec = make_int4(-1, 0, 0, 0);
        }
        RouteConst ci = {0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
        if (lane < ngr) {
            prefetch_row<S>(rows_win + (size_t)rec.x * a.Tp);
            if (rec.y & 1) ci = route_const(a, m, rec.x, a.hcar);
        }
        const int up0 = __shfl_sync(FULL, rec.z, 0);
        const int nup_all = __shfl_sync(FULL, rec.z + (rec.y >> 8), ngr - 1) - up0;
        int2 ent = make_int2(-1, UP_HEAVY);
        long long tw0 = 0;
        if (prof) tw0 = clock64();
        if (lane < nup_all) {
            ent = tp.tup[up0 + lane];
            if (ent.y >= 0)
                while (ld_acquire(done + ent.y) < epoch) __nanosleep(64);   // tributary of another task
            if (ent.y > UP_HEAVY) prefetch_row<S>(rows_win + (size_t)ent.x * a.Tp);
        }
        __syncwarp();
        if (prof) t_wait += clock64() - tw0;
        int ns1 = -1, ns2 = -1;                                         // group entry indices staged in slots 1 and 2
        auto stage_cell = [&](int c) {
            const int jn = __shfl_sync(FULL, rec.x, c);
            const int mn = __shfl_sync(FULL, rec.y, c);
            const int un = __shfl_sync(FULL, rec.z, c) - up0;
            stage_row<S>(stg, 0, lane, rows_lane + (size_t)jn * a.Tp);
            ns1 = -1; ns2 = -1;
            const int nupn = (mn & 1) ? (mn >> 8) : 0;
            for (int e = 0; e < nupn && ns2 < 0; e++) {
                const int idx = un + e;
                if (idx >= 32) break;
                const int ux = __shfl_sync(FULL, ent.x, idx), uy = __shfl_sync(FULL, ent.y, idx);
                if (uy > UP_HEAVY) {
                    stage_row<S>(stg, ns1 < 0 ? 1 : 2, lane, rows_lane + (size_t)ux * a.Tp);
                    if (ns1 < 0) ns1 = idx; else ns2 = idx;
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        stage_cell(0);
#pragma unroll 1
        for (int c = 0; c < ngr; c++) {
            const int j = __shfl_sync(FULL, rec.x, c);
            const int meta = __shfl_sync(FULL, rec.y, c);
            const int uo = __shfl_sync(FULL, rec.z, c) - up0;
            const int s1 = ns1, s2 = ns2;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            float qt[S];
            staged_row<S>(stg, 0, lane, qt);
            if (meta & 1) {
                float x[S];
#pragma unroll
                for (int s = 0; s < S; s++) x[s] = 0.0f;
                const int nup = meta >> 8;
#pragma unroll 1
                for (int e = 0; e < nup; e++) {                         // md_routing_operator.f90:37-53, same order
                    const int idx = uo + e;
                    int2 u;
                    u.x = __shfl_sync(FULL, ent.x, idx & 31); u.y = __shfl_sync(FULL, ent.y, idx & 31);
                    if (idx >= 32) {                                     // a single cell with more entries than lanes: never on a D8 mesh
                        u = tp.tup[up0 + idx];
                        if (u.y >= 0) wait_flag(done + u.y, epoch, lane);
                    }
                    if (u.y <= UP_HEAVY) {
#pragma unroll
                        for (int s = 0; s < S; s++) x[s] = x[s] + r[s];
                    } else {
                        float v[S];
                        if (idx == s1) staged_row<S>(stg, 1, lane, v);
                        else if (idx == s2) staged_row<S>(stg, 2, lane, v);
                        else ld_row<S>(rows_lane + (size_t)u.x * a.Tp, v);
#pragma unroll
                        for (int s = 0; s < S; s++) x[s] = x[s] + v[s];
                    }
                }
                if (c + 1 < ngr) stage_cell(c + 1);                     // travels while this cell is routed
                const RouteConst cc = shfl_const(ci, c);
                route_cell<S, TAPE>(a, cc, m, j, w, lane, t_first, (meta & 2) != 0, x, qt, r);
            } else {
                if (c + 1 < ngr) stage_cell(c + 1);
#pragma unroll
                for (int s = 0; s < S; s++) r[s] = qt[s];               // source cell at the chain head: already final
            }
        }
    }
    publish_flag(done + task, epoch, lane);
    if (lane == 0) notify_consumer(a, task);
    if (prof && lane == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        unsigned long long *o = a.dbg_prof + 8 * (size_t)(task - (tp.nchain - tp.nded));
        o[0] = (unsigned long long)(ce - cb); o[1] = (unsigned long long)(clock64() - t_start); o[2] = (unsigned long long)t_wait; o[3] = gt;
    }
}

// One CTA (4 warps) per task, the time window spread over its 128 threads (S4 consecutive steps each), so that a cell
// on the critical path of a river costs S4-step passes, one warp scan and two CTA barriers.  Shared memory holds the
// next cell's own row and its first two tributary rows (cp.async while the current cell is routed).
struct CellPlan {                // decoded once per group, lane-parallel, then read by every thread (broadcast)
    int j, meta;                 // meta: bit 0 routed, bit 1 gauge, bits 8-11 terms of the inflow sum, bits 12-15 position of the
    int lat[3];                  //       chain predecessor among them (15: none), bit 16: more than three tributaries
    int up_off, nup;             // lat: first three tributary cells in summation order (-1: none); full entry list for the rest
    float fa1, s_q, E, c0, h0;
};
template <int S4> struct ChainShared {
    float stg[4][128 * S4];      // staged row windows: slot 0 own qt row, slots 1-3 first three tributaries
    float warp_b[4];             // scan: end value of each warp's segment
    CellPlan plan[32];
};

template <int S4, int TAPE>
__device__ __forceinline__ void route_chain_cta(const SplitArgs &a, ChainShared<S4> &sh, int m, int task, int w, int t_first, int epoch) {
    constexpr int W = 128 * S4;
    const SplitTopo &tp = a.tp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cb = tp.task_begin[task], ce = tp.task_begin[task + 1];
    int *done = a.done + (size_t)m * tp.ntask;
    const float *rows_lane = a.rows + (size_t)m * tp.npad * a.Tp + t_first;
    const float *rows_win = a.rows + (size_t)m * tp.npad * a.Tp + (size_t)w * W;
    const int tcap = min(a.T, (w + 1) * W) - 1;                  // last valid step of this window: its state is carried on
    float r[S4];
#pragma unroll
    for (int s = 0; s < S4; s++) r[s] = 0.0f;
    const bool prof = a.dbg_prof != nullptr && task >= tp.nchain - tp.nded;
    long long t_start = 0, t_wait = 0;
    if (prof) t_start = clock64();
    int ngr = 0;
#pragma unroll 1
    for (int g0 = cb; g0 < ce; g0 += ngr) {
        ngr = min(32, ce - g0);
        long long tw0 = 0;
        if (prof) tw0 = clock64();
        __syncthreads();                                                 // the previous group's plans are no longer read
        if (warp == 0) {
            // lane = cell: decode its record and inflow entries, await its tributaries, ask DRAM for its rows
            if (lane < ngr) {
                const int4 rec = tp.tcell[g0 + lane];
                CellPlan P;
                P.j = rec.x; P.up_off = rec.z; P.nup = (rec.y & 1) ? (rec.y >> 8) : 0;
                P.lat[0] = P.lat[1] = P.lat[2] = -1;
                int nlat = 0, hpos = 15;
                for (int e = 0; e < P.nup; e++) {
                    const int2 u = tp.tup[rec.z + e];
                    if (u.y <= UP_HEAVY) { hpos = e; continue; }
                    if (u.y >= 0)
                        while (ld_acquire(done + u.y) < epoch) __nanosleep(64);   // tributary of another task
                    prefetch_row<4 * S4>(rows_win + (size_t)u.x * a.Tp);
                    if (nlat < 3) P.lat[nlat] = u.x;
                    nlat++;
                }
                P.meta = (rec.y & 3) | (P.nup << 8) | (hpos << 12) | (nlat > 3 ? 1 << 16 : 0);
                RouteConst ci = {0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
                if (rec.y & 1) ci = route_const(a, m, rec.x, a.hcar);
                P.fa1 = ci.fa1; P.s_q = ci.s_q; P.E = ci.E; P.c0 = ci.c0; P.h0 = ci.h0;
                prefetch_row<4 * S4>(rows_win + (size_t)rec.x * a.Tp);
                sh.plan[lane] = P;
            }
        }
        __syncthreads();
        if (prof) t_wait += clock64() - tw0;
        auto stage_window = [&](int slot, const float *row_window) {    // W floats = 32 * S4 chunks of 16 bytes
            for (int k = tid; k < 32 * S4; k += 128) cp_async16(&sh.stg[slot][4 * k], row_window + 4 * k);
        };
        auto stage_cell = [&](int c) {
            const CellPlan &P = sh.plan[c];
            stage_window(0, rows_win + (size_t)P.j * a.Tp);
#pragma unroll
            for (int i = 0; i < 3; i++)
                if (P.lat[i] >= 0) stage_window(1 + i, rows_win + (size_t)P.lat[i] * a.Tp);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        stage_cell(0);
#pragma unroll 1
        for (int c = 0; c < ngr; c++) {
            const CellPlan &P = sh.plan[c];
            const int j = P.j, meta = P.meta;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                             // staged rows of this cell are in shared memory
            float qt[S4];
            lds_vec<S4>(&sh.stg[0][tid * S4], qt);
            if (meta & 1) {
                float x[S4];
#pragma unroll
                for (int s = 0; s < S4; s++) x[s] = 0.0f;
                const int nup = meta >> 8 & 15, hpos = meta >> 12 & 15;
                if (!(meta >> 16 & 1)) {
#pragma unroll 1
                    for (int k = 0; k < nup; k++) {                     // md_routing_operator.f90:37-53, same order
                        if (k == hpos) {
#pragma unroll
                            for (int s = 0; s < S4; s++) x[s] = x[s] + r[s];
                        } else {
                            float v[S4];
                            lds_vec<S4>(&sh.stg[1 + k - (k > hpos ? 1 : 0)][tid * S4], v);
#pragma unroll
                            for (int s = 0; s < S4; s++) x[s] = x[s] + v[s];
                        }
                    }
                } else {                                                 // more than three tributaries (rare): rows from memory
#pragma unroll 1
                    for (int k = 0; k < nup; k++) {
                        const int2 u = tp.tup[P.up_off + k];
                        if (u.y <= UP_HEAVY) {
#pragma unroll
                            for (int s = 0; s < S4; s++) x[s] = x[s] + r[s];
                        } else {
                            float v[S4];
                            ld_vec<S4>(rows_lane + (size_t)u.x * a.Tp, v);
#pragma unroll
                            for (int s = 0; s < S4; s++) x[s] = x[s] + v[s];
                        }
                    }
                }
                const float E = P.E, h0 = P.h0;
                // local pass from a zero state (thread 0 starts from the carried state), slopes by squaring
                float h = (tid == 0) ? h0 : 0.0f;
                float A = 1.0f;
#pragma unroll
                for (int s = 0; s < S4; s++) {
                    x[s] = x[s] * P.s_q;
                    h = fmaf(h, E, x[s] * E);                            // carries only: one dependent FMA per step
                    A *= E;
                }
                float Bv = h, pw = 1.0f;                                 // pw -> (slope of one thread)^lane
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const float Bo = __shfl_up_sync(FULL, Bv, d);
                    if (lane >= d) Bv = fmaf(A, Bo, Bv);
                    if (lane & d) pw *= A;
                    A *= A;
                }
                if (lane == 31) sh.warp_b[warp] = Bv;                    // A is now the slope of a whole warp
                float hin = __shfl_up_sync(FULL, Bv, 1);
                if (lane == 0) hin = 0.0f;
                const float fa1 = P.fa1, c0 = P.c0;
                __syncthreads();                                         // warp totals visible; everybody is done with the staged rows
                if (c + 1 < ngr) stage_cell(c + 1);                      // travels while this cell is finished
                float cw = 0.0f;
                for (int k = 0; k < warp; k++) cw = fmaf(A, cw, sh.warp_b[k]);
                h = fmaf(pw, cw, hin);
                if (tid == 0) h = h0;
                float hr[S4];
#pragma unroll
                for (int s = 0; s < S4; s++) {
                    hr[s] = h + x[s];                                    // md_routing_operator.f90:73
                    const float hn = hr[s] * E;                          // :75
                    r[s] = fmaf(hr[s] - hn, fa1, qt[s]) * c0;           // :77, md_forward_structure.f90:155
                    if (t_first + s == tcap) {
                        a.hcar[(size_t)m * tp.npad + j] = hn;
                        if (w == a.nwin - 1) a.fstates[((size_t)m * 3 + 2) * tp.npad + j] = hn;
                    }
                    h = hn;
                }
                st_vec<S4>(a.rows + ((size_t)m * tp.npad + j) * a.Tp + t_first, r);
                if (TAPE) st_vec<S4>(a.rows_hr + ((size_t)m * tp.npad + j) * a.Tp + t_first, hr);
                if (meta & 2) {
                    float *qsim = a.qsim + (size_t)m * a.T * tp.ng;
                    for (int g = tp.gauge_first[j]; g >= 0; g = tp.gauge_next[g])
#pragma unroll
                        for (int s = 0; s < S4; s++)
                            if (t_first + s < a.T) qsim[(size_t)(t_first + s) * tp.ng + g] = r[s];   // :206-210
                }
            } else {
                __syncthreads();
                if (c + 1 < ngr) stage_cell(c + 1);
#pragma unroll
                for (int s = 0; s < S4; s++) r[s] = qt[s];               // source cell at the chain head: already final
            }
        }
    }
    __syncthreads();                                                     // every thread's row stores precede the release
    if (tid == 0) { st_release(done + task, epoch); notify_consumer(a, task); }
    if (prof && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        unsigned long long *o = a.dbg_prof + 8 * (size_t)(task - (tp.nchain - tp.nded));
        o[0] = (unsigned long long)(ce - cb); o[1] = (unsigned long long)(clock64() - t_start); o[2] = (unsigned long long)t_wait; o[3] = gt;
    }
}

// Export of a tile of 32 consecutive cells to the domain layout qdom[t][cell], done by a routing warp once the chains that
// own the tile's cells have finished this window (fuse_export): the transposition runs in the shadow of the serial walks
// down the main rivers instead of in a kernel of its own.  Source cells were written by the reservoir pass already.
// true when every chain that owns a cell of the tile has finished this window (one look, no waiting)
__device__ __forceinline__ bool tile_ready(const SplitArgs &a, int m, int j0, int lane, int epoch) {
    const SplitTopo &tp = a.tp;
    const int j = j0 + lane;
    bool ok = true;
    if (j < tp.n && (tp.deep ? tp.deep[j] != 0 : tp.flwacc[j] > 1)) {
        const int task = tp.cell_task[j];
        if (task >= 0 && task < tp.nchain) ok = ld_acquire(a.done + (size_t)m * tp.ntask + task) >= epoch;
    }
    return __all_sync(FULL, ok);
}
template <int S>
__device__ __forceinline__ void export_tile(const SplitArgs &a, float *tile, int m, int j0, int w, int lane, int epoch, bool wait = true) {
    constexpr int TC = 2 * S, P = TC + 1;                    // steps per chunk; padded pitch of the shared tile [32][P]
    const SplitTopo &tp = a.tp;
    const int j = j0 + lane;
    const int task = j < tp.n ? tp.cell_task[j] : -1;
    const bool mine = j < tp.n && (tp.deep ? tp.deep[j] != 0 : tp.flwacc[j] > 1) && task < tp.nchain;   // pit pairs are routed (and exported) afterwards
    if (wait && mine && task >= 0) {
        const int *flag = a.done + (size_t)m * tp.ntask + task;
        while (ld_acquire(flag) < epoch) __nanosleep(128);
    }
    __syncwarp();
    const int t_lo = w * a.W, t_hi = min(a.T, (w + 1) * a.W);
    const float *rows = a.rows + ((size_t)m * tp.npad + j0) * a.Tp;
    float *qd = a.qdom + (size_t)m * a.T * a.qpitch + j;
#pragma unroll 1
    for (int t0 = t_lo; t0 < t_hi; t0 += TC) {
#pragma unroll 8
        for (int k = 0; k < TC; k++) {                       // element e = cell * TC + step: consecutive lanes, consecutive steps
            const int e = k * 32 + lane, i = e / TC, sidx = e - i * TC;
            tile[i * P + sidx] = (j0 + i < tp.n) ? __ldcs(rows + (size_t)i * a.Tp + t0 + sidx) : 0.0f;   // rows are padded to Tp
        }
        __syncwarp();
        if (mine) {
            const int nt = min(TC, t_hi - t0);
#pragma unroll 8
            for (int tt = 0; tt < nt; tt++) qd[(size_t)(t0 + tt) * a.qpitch] = tile[lane * P + tt];
        }
        __syncwarp();
    }
}

// Tasks [0, nchain - nded) are claimed through the ticket, in dependency order, one warp per chain (lane = S steps).
// The nded longest chains (main rivers) have a CTA each (the first CTAs of the grid; thread = S/4 steps): they start at
// once and advance as their tributaries finish instead of queueing behind the rest of their basin -- a long chain is a
// serial walk and sets the kernel's duration, so it gets four warps per cell and no queue.
template <int S, int TAPE>
__global__ void __launch_bounds__(128, 3) route_forward_kernel(const SplitArgs a, const int w, const int ded_blocks) {
    union Shared {
        RowStage<S> warp_stage[4];
        ChainShared<S / 4> cta;
    };
    __shared__ __align__(16) Shared sh;
    const SplitTopo &tp = a.tp;
    const int epoch = w + 1;
    if ((int)blockIdx.x < ded_blocks) {
        const int t_first = w * a.W + threadIdx.x * (S / 4);
        const int task = tp.nchain - tp.nded + blockIdx.x;
        for (int m = 0; m < a.nmember; m++) route_chain_cta<S / 4, TAPE>(a, sh.cta, m, task, w, t_first, epoch);
        return;
    }
    const int lane = threadIdx.x & 31;
    RowStage<S> &stg = sh.warp_stage[threadIdx.x >> 5];
    const int t_first = w * a.W + lane * S;
    const int nticket = tp.nchain - tp.nded;                // pit pairs are routed by route_pairs_kernel afterwards
    const int total = nticket * a.nmember;
    if (a.dyn) {
        // no chain ready at the moment: take an export tile whose chains are done instead of waiting (one tile may be held
        // back per warp until its chains finish)
        const bool exporter = a.fuse_export > (int)(threadIdx.x >> 5);
        const int j_first = (a.first_routed / 32) * 32;
        const int ntile = (tp.n - j_first + 31) / 32;
        float *tile = reinterpret_cast<float *>(&stg);
        int pend = -1;
        bool tiles_left = exporter;
        for (;;) {
            const int task = pop_ready(a, lane, !exporter);
            if (task >= 0) { route_chain_warp<S, TAPE>(a, stg, 0, task, w, lane, t_first, epoch); continue; }
            if (task == -1) break;                               // every chain has been handed out
            if (pend < 0 && tiles_left) {
                pend = claim_ticket(a.ticket + 1, lane);
                if (pend >= ntile) { pend = -1; tiles_left = false; }
            }
            if (pend >= 0 && tile_ready(a, 0, j_first + 32 * pend, lane, epoch)) {
                export_tile<S>(a, tile, 0, j_first + 32 * pend, w, lane, epoch, false);
                pend = -1;
                continue;
            }
            __nanosleep(200);
        }
        if (pend >= 0) export_tile<S>(a, tile, 0, j_first + 32 * pend, w, lane, epoch);
    } else {
        int tk_next = claim_ticket(a.ticket, lane);
        for (;;) {
            const int tk = tk_next;
            if (tk >= total) break;
            tk_next = claim_ticket(a.ticket, lane);            // the next ticket travels while this task is routed
            const int m = tk / nticket, task = tk - m * nticket;
            route_chain_warp<S, TAPE>(a, stg, m, task, w, lane, t_first, epoch);
        }
    }
    if (a.fuse_export > (int)(threadIdx.x >> 5)) {            // fuse_export = warps per CTA that take export tiles
        // every chain has been claimed: an export tile only ever waits for chains that are running or done
        const int j_first = (a.first_routed / 32) * 32;
        const int ntile = (tp.n - j_first + 31) / 32;
        const int total_e = ntile * a.nmember;
        static_assert(sizeof(RowStage<S>) >= 32 * (2 * S + 1) * sizeof(float), "export tile does not fit the warp's staging area");
        float *tile = reinterpret_cast<float *>(&stg);
        for (;;) {
            const int tk = claim_ticket(a.ticket + 1, lane);
            if (tk >= total_e) break;
            const int m = tk / ntile, tl = tk - m * ntile;
            export_tile<S>(a, tile, m, j_first + 32 * tl, w, lane, epoch);
        }
    }
}

// pit pairs: terminal cells (they only drain into each other), routed after all chains; one warp per pair
template <int S, int TAPE>
__global__ void __launch_bounds__(128) route_pairs_kernel(const SplitArgs a, const int w) {
    const SplitTopo &tp = a.tp;
    const int lane = threadIdx.x & 31;
    const int npair = tp.ntask - tp.nchain;
    const int p = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (p >= npair * a.nmember) return;
    const int m = p / npair, task = tp.nchain + (p - m * npair);
    const int cb = tp.task_begin[task];
    int *done = a.done + (size_t)m * tp.ntask;
    route_pair<S, TAPE>(a, m, tp.tcell[cb], tp.tcell[cb + 1], w, lane, w * a.W + lane * S, done, w + 1);
    publish_flag(done + task, w + 1, lane);
}

// ------------------------------------------------------------------------------------------------
// route_members: the routing pass for ensembles (compute_multiple_run: many members, one small mesh).  The members share
// the mesh, so lane = member: a warp routes one cell for 32 members at once, every lane following the same control flow,
// strictly sequentially in time -- the reference's arithmetic and summation order (md_routing_operator.f90:17-79,
// md_forward_structure.f90:155), no scan.  A CTA owns a batch of 32 members; its warps claim the routed cells in path
// order (a topological order) from a counter in shared memory and wait for the cell's routed inflows through flags in
// shared memory, so independent branches of the river network are routed concurrently.
// ------------------------------------------------------------------------------------------------
template <int TAPE>
__global__ void __launch_bounds__(512) route_members_kernel(const SplitArgs a) {
    extern __shared__ int sm_i[];            // [0] next cell, [1 ..] done flag per routed cell
    volatile int *done = sm_i + 1;
    const SplitTopo &tp = a.tp;
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * 32 + lane;
    const bool on = m < a.nmember;
    const int mm = on ? m : a.nmember - 1;   // idle lanes shadow the last member (loads only)
    const int T = a.T, nblk = (T + 7) / 8;
    for (int i = threadIdx.x; i <= tp.nrouted; i += blockDim.x) sm_i[i] = 0;
    __syncthreads();
    const size_t mrow = (size_t)mm * tp.npad * a.Tp;
    float *qsim = a.qsim + (size_t)mm * T * tp.ng;
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(&sm_i[0], 1);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= tp.nrouted) break;
        const int j = tp.rlist[idx];
        const int ub = tp.up_begin[j], nup = tp.up_begin[j + 1] - ub;
        const float *src[8];
#pragma unroll
        for (int e = 0; e < 8; e++) src[e] = nullptr;
        for (int e = 0; e < nup && e < 8; e++) {
            const int sc = tp.up[ub + e].src;
            const int ri = tp.rindex[sc];
            if (ri >= 0) {
                if (lane == 0)
                    while (done[ri] == 0) __nanosleep(100);
                __syncwarp();
            }
            src[e] = a.rows + mrow + (size_t)sc * a.Tp;
        }
        __threadfence();                     // rows published by other warps before their flag are read after it
        const int fa = tp.flwacc[j];
        const float lr = a.fields[((size_t)mm * NFIELD + F_LR) * tp.npad + j];
        const float fa1 = (float)(fa - 1);
        const float s_q = a.dt / (0.001f * a.dx * a.dx * fa1);                 // md_routing_operator.f90:55-56
        const float E = expf(-a.dt / (lr * 60.0f));                           // :75
        const float c0 = a.dx * a.dx * 0.001f / a.dt;                         // md_forward_structure.f90:155
        float h = a.fields[((size_t)mm * NFIELD + F_HLR) * tp.npad + j];
        float *rowj = a.rows + mrow + (size_t)j * a.Tp;
        float *hrj = TAPE ? a.rows_hr + mrow + (size_t)j * a.Tp : nullptr;
        const int gfirst = tp.gauge_first[j];
        // loads run one block ahead of the arithmetic: own row and the first two inflows in registers
        float nqt[8], nv0[8], nv1[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { nv0[i] = 0.0f; nv1[i] = 0.0f; }
        ld8(rowj, nqt);
        if (nup > 0) ld8(src[0], nv0);
        if (nup > 1) ld8(src[1], nv1);
#pragma unroll 1
        for (int b = 0; b < nblk; b++) {
            const int tb = b * 8;
            float qt[8], x[8], v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { qt[i] = nqt[i]; x[i] = nv0[i]; v[i] = nv1[i]; }
            if (b + 1 < nblk) {
                ld8(rowj + tb + 8, nqt);
                if (nup > 0) ld8(src[0] + tb + 8, nv0);
                if (nup > 1) ld8(src[1] + tb + 8, nv1);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = (0.0f + x[i]) + v[i];            // :37-53, neighbour order (absent inflows are 0)
#pragma unroll
            for (int e = 2; e < 8; e++)
                if (e < nup) {
                    ld8(src[e] + tb, v);
#pragma unroll
                    for (int i = 0; i < 8; i++) x[i] = x[i] + v[i];
                }
            if ((b & 3) == 0 && b + 8 < nblk) {                                 // next lines of the rows: DRAM -> L2 ahead of use
                prefetch_l2(rowj + tb + 64);
#pragma unroll
                for (int e = 0; e < 8; e++)
                    if (e < nup) prefetch_l2(src[e] + tb + 64);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float hr = h + x[i] * s_q;                                // :55-56, :73
                const float hn = hr * E;                                        // :75
                qt[i] = fmaf(hr - hn, fa1, qt[i]) * c0;                         // :77, md_forward_structure.f90:155
                x[i] = hr;
                if (tb + i < T) h = hn;
            }
            if (on) {
                st8(rowj + tb, qt);
                if (TAPE) st8(hrj + tb, x);
                if (gfirst >= 0)
                    for (int g = gfirst; g >= 0; g = tp.gauge_next[g])
#pragma unroll
                        for (int i = 0; i < 8; i++)
                            if (tb + i < T) qsim[(size_t)(tb + i) * tp.ng + g] = qt[i];     // :206-210
            }
        }
        if (on) a.fstates[((size_t)m * 3 + 2) * tp.npad + j] = h;
        __syncwarp();
        if (lane == 0) { __threadfence(); done[idx] = 1; }
    }
}

// ------------------------------------------------------------------------------------------------
// rows_to_domain: q rows of the routed cells -> qsim_domain layout [t][cell]
// ------------------------------------------------------------------------------------------------
constexpr int RD_T = 128;   // time steps per tile
__global__ void __launch_bounds__(256) rows_to_domain_kernel(const SplitArgs a, const int j_first) {
    __shared__ float tile[32][RD_T + 1];
    const int m = blockIdx.z;
    const int j0 = j_first + blockIdx.x * 32, t0 = a.t_begin + blockIdx.y * RD_T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < 32; i += 8) {
        const int j = j0 + i;
        if (j >= a.tp.n) break;
        const float *row = a.rows + ((size_t)m * a.tp.npad + j) * a.Tp + t0;
#pragma unroll
        for (int k = 0; k < RD_T / 32; k++) tile[i][k * 32 + lane] = __ldcs(row + k * 32 + lane);   // rows are padded to Tp
    }
    __syncthreads();
    const int j = j0 + lane;
    if (j < a.tp.n && (a.tp.deep ? a.tp.deep[j] != 0 : a.tp.flwacc[j] > 1))
        for (int i = warp; i < RD_T; i += 8) {
            const int t = t0 + i;
            if (t < a.t_end) a.qdom[((size_t)m * a.T + t) * a.qpitch + j] = tile[lane][i];
        }
}

// ------------------------------------------------------------------------------------------------
// route_adjoint: LINEAR_ROUTING_B (forward_db.f90:6628-6652), UPSTREAM_DISCHARGE_B (:6520-6564) and the q_b part of
// GR_A_FORWARD_B (:8104-8118) on whole time windows, chains in reverse order, cells from the chain tail upwards.
//   q_b(c,t) = sum_g qsim_b(g,t) + w(down(c), t),     qrout_b = (flwacc-1) * c0 * q_b
//   G(t) = hr_imd_b(t) = qrout_b(t) + E * (G(t+1) - qrout_b(t)),   lr_b += dt * E * hr_imd(t) * (G(t+1) - qrout_b(t)) / (60 lr^2)
//   w(c,t) = s * G(t)  is what every inflow of c adds to its own q_b
// ------------------------------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ void add_seeds(const SplitArgs &a, int m, int j, int t_first, float (&qb)[S]) {
    const float *sb = a.qsim_b + (size_t)m * a.T * a.tp.ng;
    for (int g = a.tp.gauge_first[j]; g >= 0; g = a.tp.gauge_next[g])
#pragma unroll
        for (int s = 0; s < S; s++)
            if (t_first + s < a.T) qb[s] = qb[s] + sb[(size_t)(t_first + s) * a.tp.ng + g];        // :8104-8108
}

template <int S>
__device__ __forceinline__ void route_cell_b(const SplitArgs &a, const RouteConst &c, int m, int j, int w, int lane, int t_first,
                                             float (&qb)[S], const float (&hr)[S], float (&wv)[S]) {
    const float E = c.E;
    const float g0 = c.h0;                                    // hlr_b carried in from the next window
    const float k1 = c.fa1 * c.c0;
    float G = (lane == 31) ? g0 : 0.0f;
    float A = 1.0f;
#pragma unroll
    for (int s = S - 1; s >= 0; s--) {
        qb[s] = (t_first + s < a.T) ? k1 * qb[s] : 0.0f;     // qrout_b :8118
        G = fmaf(E, G - qb[s], qb[s]);
        A *= E;
    }
    float Bv = G;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float Bo = __shfl_down_sync(FULL, Bv, d);
        if (lane + d < 32) Bv = fmaf(A, Bo, Bv);
        A *= A;
    }
    G = __shfl_down_sync(FULL, Bv, 1);
    if (lane == 31) G = g0;
    float lrs = 0.0f;
#pragma unroll
    for (int s = S - 1; s >= 0; s--) {
        const float hrb = G - qb[s];                          // :6640
        const float Gn = qb[s] + E * hrb;                     // hr_imd_b :6641
        if (t_first + s < a.T) lrs = fmaf(hr[s], hrb, lrs);  // arg1_b / E :6643
        wv[s] = c.s_q * Gn;                                   // :6547
        G = Gn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lrs += __shfl_xor_sync(FULL, lrs, o);
    st_row<S>(a.rows_w + ((size_t)m * a.tp.npad + j) * a.Tp + t_first, wv);
    if (lane == 0) {
        float *g = a.grad + (size_t)m * NFIELD * a.tp.npad + j;
        g[(size_t)F_LR * a.tp.npad] += a.dt * (E * lrs) / ((c.lr * c.lr) * 60.0f);                // :6644
        a.gcar[(size_t)m * a.tp.npad + j] = G;                // lane 0 ends at the first step of the window
        if (w == 0) g[(size_t)F_HLR * a.tp.npad] = G;
    }
}

template <int S>
__device__ __noinline__ void route_pair_b(const SplitArgs &a, int m, int jA, int jB, int w, int lane, int t_first) {
    const size_t npad = a.tp.npad;
    float sdA[S], sdB[S], hrA[S], hrB[S];
#pragma unroll
    for (int s = 0; s < S; s++) { sdA[s] = 0.0f; sdB[s] = 0.0f; }
    add_seeds<S>(a, m, jA, t_first, sdA);
    add_seeds<S>(a, m, jB, t_first, sdB);
    ld_row<S>(a.rows_hr + ((size_t)m * npad + jA) * a.Tp + t_first, hrA);
    ld_row<S>(a.rows_hr + ((size_t)m * npad + jB) * a.Tp + t_first, hrB);
    const RouteConst cA = route_const(a, m, jA, a.gcar), cB = route_const(a, m, jB, a.gcar);
    const float kA = cA.fa1 * cA.c0, kB = cB.fa1 * cB.c0;
    float GA = cA.h0, GB = cB.h0;
    float wAn = 0.0f;                                                     // w_A(t + 1) at the last step of the window
    if (w < a.nwin - 1) wAn = a.wnext ? a.wnext[(size_t)m * npad + jA] : a.rows_w[((size_t)m * npad + jA) * a.Tp + (size_t)(w + 1) * a.W];
    float lrA = 0.0f, lrB = 0.0f;
#pragma unroll 1
    for (int L = 31; L >= 0; L--) {
        float sA = GA, sB = GB, sw = wAn;
        if (lane == L) {
#pragma unroll
            for (int s = S - 1; s >= 0; s--) {
                const bool on = t_first + s < a.T;
                const float qrB = kB * (sdB[s] + sw);           // B's discharge fed A one step later
                const float hbB = sB - qrB;
                const float GnB = qrB + cB.E * hbB;
                const float wB = cB.s_q * GnB;
                const float qrA = kA * (sdA[s] + wB);           // A's discharge fed B in the same step
                const float hbA = sA - qrA;
                const float GnA = qrA + cA.E * hbA;
                const float wA = cA.s_q * GnA;
                if (on) {
                    lrB = fmaf(hrB[s], hbB, lrB); lrA = fmaf(hrA[s], hbA, lrA);
                    sA = GnA; sB = GnB; sw = wA;
                }
                sdA[s] = on ? wA : 0.0f; sdB[s] = on ? wB : 0.0f;
            }
        }
        GA = __shfl_sync(FULL, sA, L); GB = __shfl_sync(FULL, sB, L); wAn = __shfl_sync(FULL, sw, L);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lrA += __shfl_xor_sync(FULL, lrA, o); lrB += __shfl_xor_sync(FULL, lrB, o); }
    st_row<S>(a.rows_w + ((size_t)m * npad + jA) * a.Tp + t_first, sdA);
    st_row<S>(a.rows_w + ((size_t)m * npad + jB) * a.Tp + t_first, sdB);
    if (lane == 0) {
        float *g = a.grad + (size_t)m * NFIELD * npad;
        g[(size_t)F_LR * npad + jA] += a.dt * (cA.E * lrA) / ((cA.lr * cA.lr) * 60.0f);
        g[(size_t)F_LR * npad + jB] += a.dt * (cB.E * lrB) / ((cB.lr * cB.lr) * 60.0f);
        a.gcar[(size_t)m * npad + jA] = GA; a.gcar[(size_t)m * npad + jB] = GB;
        if (w == 0) { g[(size_t)F_HLR * npad + jA] = GA; g[(size_t)F_HLR * npad + jB] = GB; }
    }
}

template <int S>
__global__ void __launch_bounds__(128) route_adjoint_kernel(const SplitArgs a, const int w) {
    const int lane = threadIdx.x & 31;
    const SplitTopo &tp = a.tp;
    const int t_first = w * a.W + lane * S;
    const int epoch = a.nwin - w;
    const int total = tp.ntask * a.nmember;
    int tk_next = claim_ticket(a.ticket, lane);
    for (;;) {
        const int tk = tk_next;
        if (tk >= total) break;
        tk_next = claim_ticket(a.ticket, lane);
        const int m = tk / tp.ntask, task = tp.ntask - 1 - (tk - m * tp.ntask);
        const int cb = tp.task_begin[task], ce = tp.task_begin[task + 1];
        int *rdone = a.rdone + (size_t)m * tp.ntask;
        if (task >= tp.nchain) {
            route_pair_b<S>(a, m, tp.task_cells[cb], tp.task_cells[cb + 1], w, lane, t_first);
            publish_flag(rdone + task, (epoch << 16) | 0xffff, lane);
            continue;
        }
        const size_t mrow = (size_t)m * tp.npad * a.Tp;
        float wv[S];
        {   // w of the consumer of the chain tail (another task), zero at an outlet
            const int jt = tp.task_cells[ce - 1];
            const int d = tp.down[jt];
            if (d >= 0) {
                // the consumer's chain streams its progress (cells done, counted from its tail)
                wait_flag(rdone + tp.down_task[jt], (epoch << 16) | tp.down_need[jt], lane);
                ld_row<S>(a.rows_w + mrow + (size_t)d * a.Tp + t_first, wv);
            } else {
#pragma unroll
                for (int s = 0; s < S; s++) wv[s] = 0.0f;
            }
        }
        int ndone = 0;
#pragma unroll 1
        for (int g1 = ce; g1 > cb; g1 -= 32) {
            const int g0 = max(cb, g1 - 32), ngr = g1 - g0;
            int4 rec = make_int4(-1, 0, 0, 0);
            if (lane < ngr) rec = tp.tcell[g0 + lane];
            RouteConst ci = {0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
            if (lane < ngr && (rec.y & 1)) {
                prefetch_row<S>(a.rows_hr + mrow + (size_t)rec.x * a.Tp + (size_t)w * a.W);
                ci = route_const(a, m, rec.x, a.gcar);
            }
#pragma unroll 1
            for (int c = ngr - 1; c >= 0; c--) {
                const int j = __shfl_sync(FULL, rec.x, c);
                const int meta = __shfl_sync(FULL, rec.y, c);
                if (!(meta & 1)) break;                            // source cell at the chain head: nothing to route
                float hr[S], qb[S];
                ld_row<S>(a.rows_hr + mrow + (size_t)j * a.Tp + t_first, hr);
#pragma unroll
                for (int s = 0; s < S; s++) qb[s] = wv[s];
                if (meta & 2) add_seeds<S>(a, m, j, t_first, qb);
                const RouteConst cc = shfl_const(ci, c);
                route_cell_b<S>(a, cc, m, j, w, lane, t_first, qb, hr, wv);
                if ((++ndone & 7) == 0) publish_flag(rdone + task, (epoch << 16) | min(ndone, 0xfffe), lane);   // tributaries may start
            }
        }
        publish_flag(rdone + task, (epoch << 16) | 0xffff, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// vertical_adjoint: reverse time loop of the reservoirs of one cell (vertical_step_b, cell_math.cuh).
// qt_b(c,t) = dx^2 * 1e-3 / dt * (sum_g qsim_b(g,t) + w(down(c), t)), forward_db.f90:8104-8118.
// ------------------------------------------------------------------------------------------------
typedef float BwdStage[4][VT_TK][32];

template <int FAST>
__global__ void __launch_bounds__(VT_WARPS * 32) vertical_adjoint_kernel(const __grid_constant__ CUtensorMap tm_prcp,
                                                                        const __grid_constant__ CUtensorMap tm_pet,
                                                                        const __grid_constant__ CUtensorMap tm_hp,
                                                                        const __grid_constant__ CUtensorMap tm_hft,
                                                                        const SplitArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    BwdStage *stage = reinterpret_cast<BwdStage *>(smem_raw) + warp * VB_NST;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + sizeof(BwdStage) * VT_WARPS * VB_NST) + warp * VB_NST;
    const int m = blockIdx.y;
    const int j0 = (blockIdx.x * VT_WARPS + warp) * 32;
    const int n = a.tp.n, npad = a.tp.npad, T = a.T;
    if (j0 >= n) return;
    const int j = j0 + lane;
    const bool valid = j < n;
    const int t_begin = a.t_begin, t_end = a.t_end;                      // whole run: 0, T; checkpointed runs: one window
    const int st_lo = t_begin / VT_TK;
    const int nst = (t_end + VT_TK - 1) / VT_TK - st_lo;
    constexpr uint32_t STAGE_BYTES = sizeof(BwdStage);
    const int yb = m * T - a.tape_t0;   // tape row of time step 0 of member m

    auto issue = [&](int slot, int st) {
        mbar_expect_tx(&bars[slot], STAGE_BYTES);
        tma_load_2d(&stage[slot][0][0][0], &tm_prcp, j0, st * VT_TK, &bars[slot]);
        tma_load_2d(&stage[slot][1][0][0], &tm_pet, j0, st * VT_TK, &bars[slot]);
        tma_load_2d(&stage[slot][2][0][0], &tm_hp, j0, yb + st * VT_TK, &bars[slot]);
        tma_load_2d(&stage[slot][3][0][0], &tm_hft, j0, yb + st * VT_TK, &bars[slot]);
    };
    if (lane == 0) {
        for (int s = 0; s < VB_NST; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < VB_NST && s < nst; s++) issue(s, st_lo + nst - 1 - s);
    }
    __syncwarp();

    const float *fld = a.fields + (size_t)m * NFIELD * npad + j;
    CellConst k = make_const(1.0f, 1.0f, 0.0f, 1.0f, 1, a.dt, a.dx);
    int d = -1, dlag = 0, gfirst = -1;
    if (valid) {
        k = make_const(fld[(size_t)F_CP * npad], fld[(size_t)F_CFT * npad], fld[(size_t)F_EXC * npad], fld[(size_t)F_LR * npad],
                       a.tp.flwacc[j], a.dt, a.dx, a.grd != 0);
        d = a.tp.down[j]; dlag = a.tp.down_lag[j]; gfirst = a.tp.gauge_first[j];
    }
    const float *wrow = (d >= 0) ? a.rows_w + ((size_t)m * npad + d) * a.Tp : nullptr;
    const float *sb = a.qsim_b + (size_t)m * T * a.tp.ng;
    const int ng = a.tp.ng;
    float hp_b = 0.0f, hft_b = 0.0f, cp_b = 0.0f, cft_b = 0.0f, exc_b = 0.0f;
    if (valid && t_end < T) {                                            // a later window ran already: its adjoint states and sums
        const float *g = a.grad + (size_t)m * NFIELD * npad + j;
        cp_b = g[(size_t)F_CP * npad]; cft_b = g[(size_t)F_CFT * npad]; exc_b = g[(size_t)F_EXC * npad];
        hp_b = g[(size_t)F_HP * npad]; hft_b = g[(size_t)F_HFT * npad];
    }
    const float wnx = (d >= 0 && dlag && a.wnext && t_end < T) ? a.wnext[(size_t)m * npad + d] : 0.0f;
    const bool all_valid = j0 + 32 <= n;
    const bool exc_on = __any_sync(FULL, k.exc != 0.0f);
    const float inv_cft2 = k.inv_cft * k.inv_cft, inv_cp2 = k.inv_cp * k.inv_cp;
    const float c0 = k.c0;

    auto load_w = [&](int st, float *v) {
#pragma unroll
        for (int i = 0; i < VT_TK; i++) v[i] = 0.0f;
        if (wrow) {
            if (!dlag) ld8(wrow + (size_t)st * VT_TK, v);
            else {   // late cell of a pit pair: its discharge fed the partner one step later
#pragma unroll
                for (int i = 0; i < VT_TK; i++) {
                    const int t = st * VT_TK + i + 1;
                    float x = 0.0f;
                    if (t < t_end) x = __ldcg(wrow + t);
                    else if (t < T) x = a.wnext ? wnx : __ldcg(wrow + t);     // first step of the next window
                    v[i] = x;
                }
            }
        }
    };
    float wn[VT_TK];
    load_w(st_lo + nst - 1, wn);
    uint32_t parity = 0;
    int slot = 0;
#pragma unroll 1
    for (int it = 0; it < nst; it++) {
        const int st = st_lo + nst - 1 - it;
        float wq[VT_TK];
#pragma unroll
        for (int i = 0; i < VT_TK; i++) wq[i] = wn[i];
        if (st > st_lo) load_w(st - 1, wn);
        mbar_wait(&bars[slot], parity);
        float pv[VT_TK], ev[VT_TK], hpv[VT_TK], hfv[VT_TK];
#pragma unroll
        for (int i = 0; i < VT_TK; i++) {
            pv[i] = stage[slot][0][i][lane]; ev[i] = stage[slot][1][i][lane];
            hpv[i] = stage[slot][2][i][lane]; hfv[i] = stage[slot][3][i][lane];
        }
        __syncwarp();
        if (lane == 0 && it + VB_NST < nst) issue(slot, st - VB_NST);
        float mn = 0.0f, mx = 0.0f, hmx = 0.0f;
#pragma unroll
        for (int i = 0; i < VT_TK; i++) {
            mn = fminf(mn, fminf(pv[i], ev[i]));
            mx = fmaxf(mx, fmaxf(pv[i], ev[i]));
            hmx = fmaxf(hmx, hpv[i]);
        }
        const float xm = mx * k.inv_cp;
        const bool lean = FAST && all_valid && (st + 1) * VT_TK <= t_end &&  // warp-uniform part first: every lane votes
                          __all_sync(FULL, gfirst < 0 && mn >= 0.0f && xm < 0.25f && hmx + xm < 15.0f);
        if (lean) {
#pragma unroll
            for (int i = VT_TK - 1; i >= 0; i--) {
                const float qt_b = c0 * wq[i];                                                     // :8114
                if (exc_on) vertical_step_b_lean<true>(k, inv_cft2, inv_cp2, pv[i], ev[i], hpv[i], hfv[i], qt_b, hp_b, hft_b, cp_b, cft_b, exc_b);
                else vertical_step_b_lean<false>(k, inv_cft2, inv_cp2, pv[i], ev[i], hpv[i], hfv[i], qt_b, hp_b, hft_b, cp_b, cft_b, exc_b);
            }
        } else
#pragma unroll
        for (int i = VT_TK - 1; i >= 0; i--) {
            const int t = st * VT_TK + i;
            if (valid && t < t_end) {
                float q_b = wq[i];
                if (gfirst >= 0)
                    for (int g = gfirst; g >= 0; g = a.tp.gauge_next[g]) q_b += sb[(size_t)t * ng + g];
                const float qt_b = fdiv<FAST>(a.dx * a.dx * 0.001f * q_b, a.dt);                   // :8114
                vertical_step_b<FAST>(k, pv[i], ev[i], hpv[i], hfv[i], qt_b, hp_b, hft_b, cp_b, cft_b, exc_b);
            }
        }
        if (++slot == VB_NST) { slot = 0; parity ^= 1u; }
    }
    if (valid) {
        float *g = a.grad + (size_t)m * NFIELD * npad + j;
        g[(size_t)F_CP * npad] = cp_b; g[(size_t)F_CFT * npad] = cft_b;
        g[(size_t)F_EXC * npad] = a.grd ? 0.0f : exc_b;                       // gr-d has no exchange: parameters_b%exc stays 0 (forward_db.f90:9604-9797)
        g[(size_t)F_HP * npad] = hp_b; g[(size_t)F_HFT * npad] = hft_b;
    }
}

// ------------------------------------------------------------------------------------------------
// layout kernels
// ------------------------------------------------------------------------------------------------
__global__ void pack_columns_kernel(const float *raw, int64_t stride, const int32_t *idx, int n, int npad, int T, float *out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npad) return;
    const int src = j < n ? idx[j] : -1;
    for (int t = blockIdx.y; t < T; t += gridDim.y) out[(size_t)t * npad + j] = src >= 0 ? raw[(int64_t)t * stride + src] : 0.0f;
}
__global__ void scatter_columns_kernel(const float *src, int64_t pitch, const int32_t *idx, int n, int T, int64_t stride, float *out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int dst = idx[j];
    if (dst < 0) return;
    for (int t = blockIdx.y; t < T; t += gridDim.y) out[(int64_t)t * stride + dst] = src[(size_t)t * pitch + j];
}
__global__ void sum_domain_kernel(const float *src, int64_t pitch, int n, int T, double *out) {
    double acc = 0.0;
    for (int t = blockIdx.y; t < T; t += gridDim.y)
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) acc += (double)src[(size_t)t * pitch + j];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int split_pick_window(int T, int *S, int *nwin, bool small) {
    if (small) { *S = 8; *nwin = (T + 255) / 256; return 256; }
    const int nw = (T + 1023) / 1024;
    const int need = (T + nw - 1) / nw;
    int s = 8;
    while (32 * s < need) s += 8;
    *S = s; *nwin = nw;
    return 32 * s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map_2d(CUtensorMap *tm, const float *base, uint64_t cols, uint64_t rows, uint64_t pitch_elems, const char **err) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || !p) {
            cudaGetLastError();
            *err = "cuTensorMapEncodeTiled is not available from the driver";
            return 1;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    if ((pitch_elems * sizeof(float)) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) % 16) != 0) {
        *err = "tensor map: base and row pitch must be multiples of 16 bytes";
        return 1;
    }
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {pitch_elems * sizeof(float)};
    const cuuint32_t box[2] = {32, (cuuint32_t)VT_TK};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled failed"; return 1; }
    return 0;
}

cudaError_t launch_vertical_forward(const SplitArgs &a, const CUtensorMap &prcp, const CUtensorMap &pet, int math_mode, bool tape,
                                    cudaStream_t s) {
    dim3 grid((unsigned)((a.tp.n + VT_WARPS * 32 - 1) / (VT_WARPS * 32)), (unsigned)a.nmember);
    const size_t smem = sizeof(FwdStage) * VT_WARPS * VF_NST + sizeof(uint64_t) * VT_WARPS * VF_NST;
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, VT_WARPS * 32, smem, s>>>(prcp, pet, a);
        return cudaGetLastError();
    };
    if (math_mode) return tape ? go(vertical_forward_kernel<1, 1>) : go(vertical_forward_kernel<1, 0>);
    return tape ? go(vertical_forward_kernel<0, 1>) : go(vertical_forward_kernel<0, 0>);
}

template <typename K> static cudaError_t persistent_grid(K kern, int *blocks) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0);
    if (e != cudaSuccess) return e;
    *blocks = sms * (per_sm > 0 ? per_sm : 1);
    return cudaSuccess;
}

template <int S> static cudaError_t route_forward_windows(const SplitArgs &a, bool tape, cudaStream_t s, int w_begin, int w_end) {
    int blocks = 0;
    cudaError_t e = tape ? persistent_grid(route_forward_kernel<S, 1>, &blocks) : persistent_grid(route_forward_kernel<S, 0>, &blocks);
    if (e != cudaSuccess) return e;
    const int ded_blocks = a.tp.nded;                      // one CTA per dedicated chain (route_graph.cpp keeps nded small)
    const long long total = (long long)(a.tp.nchain - a.tp.nded) * a.nmember;
    const long long need = (total + 3) / 4 + ded_blocks;
    const int npair = (a.tp.ntask - a.tp.nchain) * a.nmember;
    if (blocks > need) blocks = need > 0 ? (int)need : 1;
    if (blocks <= ded_blocks) {
        // every chain is a dedicated one (e.g. a straight channel): one more CTA for the export tiles; the dedicated CTAs
        // must all be resident
        int limit = 0;
        e = tape ? persistent_grid(route_forward_kernel<S, 1>, &limit) : persistent_grid(route_forward_kernel<S, 0>, &limit);
        if (e != cudaSuccess) return e;
        if (ded_blocks + 1 > limit) return cudaErrorLaunchOutOfResources;
        blocks = ded_blocks + 1;
    }
    for (int w = w_begin; w < w_end; w++) {
        e = cudaMemsetAsync(a.ticket, 0, 2 * sizeof(unsigned int), s);
        if (e != cudaSuccess) return e;
        if (a.dyn) {                                         // ready queues, tributary counters and queue heads / tails of this window
            const size_t nt = (size_t)(a.tp.nchain - a.tp.nded);
            e = cudaMemcpyAsync(a.queue, a.queue0, nt * sizeof(int), cudaMemcpyDeviceToDevice, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(a.ndep, a.ndep0, nt * sizeof(int), cudaMemcpyDeviceToDevice, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(a.qctl, a.qctl0, 32 * sizeof(unsigned int), cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) return e;
        }
        if (a.tp.nchain > 0 || a.fuse_export > 1) {                       // a graph of pit pairs only has nothing for this kernel
            if (tape) route_forward_kernel<S, 1><<<blocks, 128, 0, s>>>(a, w, ded_blocks);
            else route_forward_kernel<S, 0><<<blocks, 128, 0, s>>>(a, w, ded_blocks);
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
        if (npair > 0) {
            if (tape) route_pairs_kernel<S, 1><<<(npair + 3) / 4, 128, 0, s>>>(a, w);
            else route_pairs_kernel<S, 0><<<(npair + 3) / 4, 128, 0, s>>>(a, w);
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

cudaError_t launch_route_members(const SplitArgs &a, bool tape, cudaStream_t s) {
    if (a.tp.nrouted == 0) return cudaSuccess;
    const size_t smem = sizeof(int) * (size_t)(a.tp.nrouted + 1);
    const unsigned grid = (unsigned)((a.nmember + 31) / 32);
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) return cudaErrorInvalidValue;
    if (tape) route_members_kernel<1><<<grid, 512, smem, s>>>(a);
    else route_members_kernel<0><<<grid, 512, smem, s>>>(a);
    e = cudaGetLastError();
    return e;
}

static cudaError_t route_forward_range(const SplitArgs &a, bool tape, cudaStream_t s, int w_begin, int w_end, bool reset = false) {
    if (a.tp.ntask == 0) return cudaSuccess;
    if (w_begin == 0 || reset) {
        cudaError_t e = cudaMemsetAsync(a.done, 0, 2 * sizeof(int) * (size_t)a.tp.ntask * a.nmember, s);   // flags + block counters
        if (e != cudaSuccess) return e;
    }
    switch (a.W / 32) {
        case 8: return route_forward_windows<8>(a, tape, s, w_begin, w_end);
        case 16: return route_forward_windows<16>(a, tape, s, w_begin, w_end);
        case 24: return route_forward_windows<24>(a, tape, s, w_begin, w_end);
        default: return route_forward_windows<32>(a, tape, s, w_begin, w_end);
    }
}
cudaError_t launch_route_forward(const SplitArgs &a, bool tape, cudaStream_t s) { return route_forward_range(a, tape, s, 0, a.nwin); }
cudaError_t launch_route_forward_window(const SplitArgs &a, int w, bool tape, cudaStream_t s, bool reset) {
    return route_forward_range(a, tape, s, w, w + 1, reset);
}

template <int S> static cudaError_t route_adjoint_windows(const SplitArgs &a, cudaStream_t s, int w_hi, int w_lo) {
    int blocks = 0;
    cudaError_t e = persistent_grid(route_adjoint_kernel<S>, &blocks);
    if (e != cudaSuccess) return e;
    const long long total = (long long)a.tp.ntask * a.nmember;
    const int need = (int)((total + 3) / 4);
    if (blocks > need) blocks = need > 0 ? need : 1;
    for (int w = w_hi; w >= w_lo; w--) {
        e = cudaMemsetAsync(a.ticket, 0, sizeof(unsigned int), s);
        if (e != cudaSuccess) return e;
        route_adjoint_kernel<S><<<blocks, 128, 0, s>>>(a, w);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static cudaError_t route_adjoint_range(const SplitArgs &a, cudaStream_t s, int w_hi, int w_lo) {
    if (a.tp.ntask == 0) return cudaSuccess;
    if (w_hi == a.nwin - 1) {
        cudaError_t e = cudaMemsetAsync(a.rdone, 0, sizeof(int) * (size_t)a.tp.ntask * a.nmember, s);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(a.gcar, 0, sizeof(float) * (size_t)a.tp.npad * a.nmember, s);
        if (e != cudaSuccess) return e;
    }
    switch (a.W / 32) {
        case 8: return route_adjoint_windows<8>(a, s, w_hi, w_lo);
        case 16: return route_adjoint_windows<16>(a, s, w_hi, w_lo);
        case 24: return route_adjoint_windows<24>(a, s, w_hi, w_lo);
        default: return route_adjoint_windows<32>(a, s, w_hi, w_lo);
    }
}
cudaError_t launch_route_adjoint(const SplitArgs &a, cudaStream_t s) { return route_adjoint_range(a, s, a.nwin - 1, 0); }
cudaError_t launch_route_adjoint_window(const SplitArgs &a, int w, cudaStream_t s) { return route_adjoint_range(a, s, w, w); }

__global__ void first_step_kernel(const float *rows, int pitch, int npad, float *out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < npad) out[j] = rows[(size_t)j * pitch];
}
cudaError_t launch_first_step(const float *rows, int pitch, int npad, float *out, cudaStream_t s) {
    first_step_kernel<<<(npad + 255) / 256, 256, 0, s>>>(rows, pitch, npad, out);
    return cudaGetLastError();
}

cudaError_t launch_rows_to_domain(const SplitArgs &a, cudaStream_t s) {
    const int j_first = (a.first_routed / 32) * 32;
    if (j_first >= a.tp.n) return cudaSuccess;
    dim3 grid((unsigned)((a.tp.n - j_first + 31) / 32), (unsigned)((a.t_end - a.t_begin + RD_T - 1) / RD_T), (unsigned)a.nmember);
    rows_to_domain_kernel<<<grid, 256, 0, s>>>(a, j_first);
    return cudaGetLastError();
}

cudaError_t launch_vertical_adjoint(const SplitArgs &a, const CUtensorMap &prcp, const CUtensorMap &pet, const CUtensorMap &hp,
                                    const CUtensorMap &hft, int math_mode, cudaStream_t s) {
    dim3 grid((unsigned)((a.tp.n + VT_WARPS * 32 - 1) / (VT_WARPS * 32)), (unsigned)a.nmember);
    const size_t smem = sizeof(BwdStage) * VT_WARPS * VB_NST + sizeof(uint64_t) * VT_WARPS * VB_NST;
    cudaError_t e;
    if (math_mode) {
        e = cudaFuncSetAttribute(vertical_adjoint_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        vertical_adjoint_kernel<1><<<grid, VT_WARPS * 32, smem, s>>>(prcp, pet, hp, hft, a);
    } else {
        e = cudaFuncSetAttribute(vertical_adjoint_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        vertical_adjoint_kernel<0><<<grid, VT_WARPS * 32, smem, s>>>(prcp, pet, hp, hft, a);
    }
    return cudaGetLastError();
}

cudaError_t launch_pack_columns(const float *raw, int64_t stride, const int32_t *idx, int n, int npad, int T, float *out,
                                cudaStream_t s) {
    dim3 grid((unsigned)((npad + 255) / 256), (unsigned)(T < 64 ? T : 64));
    pack_columns_kernel<<<grid, 256, 0, s>>>(raw, stride, idx, n, npad, T, out);
    return cudaGetLastError();
}
cudaError_t launch_scatter_columns(const float *src, int64_t pitch, const int32_t *idx, int n, int T, int64_t stride, float *out,
                                   cudaStream_t s) {
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)(T < 64 ? T : 64));
    scatter_columns_kernel<<<grid, 256, 0, s>>>(src, pitch, idx, n, T, stride, out);
    return cudaGetLastError();
}
// out[t * stride + didx[i]] = src[t * pitch + sidx[i]], i < n: a handful of columns from one layout into another
__global__ void copy_columns_kernel(const float *src, int64_t pitch, const int32_t *sidx, const int32_t *didx, int n, int T, int64_t stride,
                                    float *out) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const int a = sidx[i], b = didx[i];
    for (int t = threadIdx.x; t < T; t += blockDim.x) out[(int64_t)t * stride + b] = src[(int64_t)t * pitch + a];
}
cudaError_t launch_copy_columns(const float *src, int64_t pitch, const int32_t *sidx, const int32_t *didx, int n, int T, int64_t stride,
                                float *out, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    copy_columns_kernel<<<n, 256, 0, s>>>(src, pitch, sidx, didx, n, T, stride, out);
    return cudaGetLastError();
}
cudaError_t launch_sum_domain(const float *src, int64_t pitch, int n, int T, double *out, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), s);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)std::min(1024, (n + 255) / 256), (unsigned)(T < 64 ? T : 64));
    sum_domain_kernel<<<grid, 256, 0, s>>>(src, pitch, n, T, out);
    return cudaGetLastError();
}

}  // namespace smash
