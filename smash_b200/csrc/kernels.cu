// kernels.cu -- hand-written sm_100a kernels of the smash gr-a forward / adjoint solver.
//
// Execution model: one CTA per block of B consecutive cells of the post-ordered
// drainage forest, one thread per cell, the whole time loop inside the kernel.  Reservoir states hp, hft,
// hlr live in registers.  At tick d a cell works on time step t = d - off (in-block skew), so every
// in-block inflow was produced exactly one tick earlier and is exchanged through a double buffer in shared
// memory with one __syncthreads per tick; inflows from other blocks come from HBM/L2 behind per-block
// progress flags (blocks only ever wait for lower-numbered blocks, numbering by an atomic ticket).
// Forcing rows [block][tick][prcp|pet][B] (and tape rows in the reverse sweep) are streamed by TMA 1-D
// bulk copies (cp.async.bulk, SASS UBLKCP) into an mbarrier-guarded shared-memory ring.
//
// Reference statements are cited as file:line under /root/reference/smash/solver/.
#include "kernels.cuh"
#include "cell_math.cuh"

#include <cstdio>

namespace smash {

// ------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------
constexpr int QX_PAD = 32;   // qx[buf][B + QX_PAD]; slot B always holds 0 (absent inflow)

// Time is processed in chunks of CHF ticks.  Phase V of a chunk runs the reservoir arithmetic of every lane for
// the chunk's ticks with NO block-wide synchronisation (it has no inter-cell dependency, md_forward_structure.f90:106-144)
// and parks qt in shared memory; phase R then routes the chunk tick by tick (cheap: a few loads, one FMA chain and one
// __syncthreads per tick).  Warps therefore only wait for each other on the cheap part.
constexpr int CHF = 8;                 // ticks per chunk, forward
constexpr int FRING = 2 * CHF;         // forcing rows resident per CTA (two chunks)

// MULTI = 0: the mesh fits one block (small catchments, ensembles): no cross-block code at all.
template <int FAST, int TAPE, int MULTI>
__global__ void __launch_bounds__(512, 2) forward_kernel(const SolverArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const DeviceTopology &tp = a.tp;
    const int B = tp.B;
    const int tid = threadIdx.x;
    const int QS = B + QX_PAD;
    float *ring = reinterpret_cast<float *>(smem_raw);                 // [FRING][2][B]
    float *qts = ring + FRING * 2 * B;                                 // [CHF][B]  qt of the current chunk
    float *xq = qts + CHF * B;                                         // MULTI: [2][CHF][B] prefetched cross-block inflows
    int *seen_s = reinterpret_cast<int *>(xq + (MULTI ? 2 * CHF * B : 0));   // MULTI: [2][B] cached producer progress
    float *qx = reinterpret_cast<float *>(seen_s + (MULTI ? 2 * B : 0));     // [2][B + QX_PAD]
    uint64_t *bars = reinterpret_cast<uint64_t *>(qx + 2 * QS);        // [FRING]
    __shared__ unsigned int s_ticket;

    if (tid == 0) {
        s_ticket = atomicAdd(a.ticket, 1u);
        for (int s = 0; s < FRING; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * QS; i += B) qx[i] = 0.0f;
    if (MULTI) {
        seen_s[tid] = 0; seen_s[B + tid] = 0;
        for (int i = 0; i < 2 * CHF; i++) xq[i * B + tid] = 0.0f;
    }
    __syncthreads();
    const int vb = (int)s_ticket;
    const int member = vb / tp.nblocks;
    const int blk = vb - member * tp.nblocks;
    const int slot = blk * B + tid;
    const int nticks = tp.T + tp.hmax[blk];
    const int64_t row0 = tp.tick_base[blk];
    const unsigned bf = tp.bflags[blk];
    const bool has_late = bf & BLK_LATE;
    const bool publish = (bf & BLK_PUBLISH) || a.save_q;
    const float *frc = a.forcing + row0 * 2 * B;
    const uint32_t row_bytes = 2u * B * sizeof(float);

    if (tid == 0) {
        for (int s = 0; s < FRING && s < nticks; s++) {
            mbar_expect_tx(&bars[s], row_bytes);
            tma_load_1d(ring + (size_t)s * 2 * B, frc + (size_t)s * 2 * B, row_bytes, &bars[s]);
        }
    }

    const bool valid = tp.cell[slot] >= 0;
    const int off = tp.off[slot];
    const int fa = valid ? tp.flwacc[slot] : 1;
    const bool late = tp.late[slot];
    const int ub = tp.up_begin[slot], ue = tp.up_begin[slot + 1];
    const int gfirst = tp.gauge_first[slot];
    const float *fld = a.fields + (size_t)member * NFIELD * tp.nslots + slot;
    float hp = 0.01f, hft = 0.01f, hlr = 0.0f;
    CellConst k = make_const(200.0f, 500.0f, 0.0f, 5.0f, 1, a.dt, a.dx);
    if (valid) {
        k = make_const(fld[(size_t)F_CP * tp.nslots], fld[(size_t)F_CFT * tp.nslots], fld[(size_t)F_EXC * tp.nslots],
                       fld[(size_t)F_LR * tp.nslots], fa, a.dt, a.dx);
        hp = fld[(size_t)F_HP * tp.nslots]; hft = fld[(size_t)F_HFT * tp.nslots]; hlr = fld[(size_t)F_HLR * tp.nslots];
    }
    // inflow lanes of the common case (<= 6 in-block producers, none in another block, not a pit-pair member), as byte
    // offsets into the exchange buffer: absent entries point at the zero slot B, so the sum keeps the reference's
    // order (i = 1..8, md_routing_operator.f90:37-53) without branches
    int u0 = B * 4, u1 = B * 4, u2 = B * 4;
    unsigned ux = (unsigned)B | ((unsigned)B << 10) | ((unsigned)B << 20);   // entries 4..6, 10 bits each (B <= 512)
    // One cross-block producer is folded into the fast path when it is the first or the last term of the sum (0 + x
    // and commutativity keep the reference's result bit for bit); other patterns take the generic loop below.
    int n_ext = 0, n_in = 0, xmode = 0;   // xmode: 0 none, 1 cross-block term first, 2 last
    bool simple = !late;
    {
        unsigned lanes[6];
        for (int e = ub; e < ue; e++) {
            const UpEntry u = tp.up[e];
            if (u.cur) simple = false;
            if (u.a < 0) {
                n_ext++;
                if (e == ub) xmode = 1; else if (e == ue - 1) xmode = 2; else simple = false;
            } else {
                if (n_in < 6) lanes[n_in] = (unsigned)u.a;
                n_in++;
            }
        }
        if (n_in > 6 || n_ext > 1 || (!MULTI && n_ext > 0)) simple = false;
        if (simple) {
            if (n_in > 0) u0 = lanes[0] * 4;
            if (n_in > 1) u1 = lanes[1] * 4;
            if (n_in > 2) u2 = lanes[2] * 4;
            const unsigned a3 = n_in > 3 ? lanes[3] : B, a4 = n_in > 4 ? lanes[4] : B, a5 = n_in > 5 ? lanes[5] : B;
            ux = a3 | (a4 << 10) | (a5 << 20);
        }
    }
    const bool many = n_in > 3;
    const bool xfirst = simple && xmode == 1, xlast = simple && xmode == 2;
    const float c_dx = a.dx, c_dt = a.dt;
    const int T = tp.T, ng = tp.ng;
    float *qdom = a.qdom ? a.qdom + ((size_t)member * tp.total_ticks + row0) * B + tid : nullptr;
    float *netp = (a.save_netp && a.netp) ? a.netp + ((size_t)member * tp.total_ticks + row0) * B + tid : nullptr;
    float *tape = TAPE ? a.tape + ((size_t)member * tp.total_ticks + row0) * 4 * B + tid : nullptr;
    float *qsim = a.qsim + (size_t)member * tp.T * tp.ng;
    const int *prog_m = a.prog + (size_t)member * tp.nblocks;
    int *prog_mine = a.prog + (size_t)member * tp.nblocks + blk;
    const bool do_flag = MULTI && (bf & BLK_PUBLISH);

    uint32_t parity = 0;
    const bool do_net = netp != nullptr;
    float *qbuf0 = qx, *qbuf1 = qx + QS;       // qbuf0 = buffer written at even ticks
#pragma unroll 1
    for (int d0 = 0; d0 < nticks; d0 += CHF) {
        const int nd = min(CHF, nticks - d0);
        const int sbase = d0 % FRING;           // d0 is a multiple of CHF and FRING = 2*CHF: 0 or CHF
        // ---------------- phase V: reservoirs, no block synchronisation
        {
            const float *rp = ring + (size_t)sbase * 2 * B + tid;
            float *qp = qts + tid;
            float *tv = tape;
            uint32_t bar_addr = smem_u32(&bars[sbase]);
            int t = d0 - off;
#pragma unroll 1
            for (int i = 0; i < nd; i++) {
                mbar_wait_addr(bar_addr, parity);
                const float prcp = rp[0];
                const float pet = rp[B];
                const bool act = valid && (unsigned)t < (unsigned)T;
                if (TAPE && act) { tv[0] = hp; tv[B] = hft; }
                float hp_n = hp, hft_n = hft, qt;
                const bool gapless = (prcp >= 0.0f) && (pet >= 0.0f);
                if (FAST && __all_sync(0xffffffffu, gapless)) qt = vertical_step_nogap(k, prcp, pet, hp_n, hft_n);
                else qt = vertical_step<FAST>(k, prcp, pet, hp_n, hft_n).qt;
                if (act) { hp = hp_n; hft = hft_n; }
                qp[0] = qt;
                rp += 2 * B; qp += B; bar_addr += 8; t++;
                if (TAPE) tv += 4 * B;
            }
        }
        // ---------------- phase R: routing, one barrier per tick
        if (MULTI && n_ext > 0 && valid) {
            int j = 0;
#pragma unroll 1
            for (int e = ub; e < ue && j < 2; e++) {
                if (tp.up[e].a >= 0) continue;
                const ExtRef x = tp.ext[-tp.up[e].a - 1];
                const int i_lo = max(0, off + x.lag - d0);            // first tick of the chunk with t - lag >= 0
                const int i_hi = min(nd - 1, T - 1 + off - d0);       // last tick of the chunk with t <= T - 1
                if (i_lo <= i_hi) {
                    const int need = d0 + i_hi + x.dtick + 1;
                    int seen = seen_s[j * B + tid];
                    if (seen < need && !a.debug_nowait) {
                        do {
                            seen = ld_acquire(prog_m + x.blk);
                            if (seen < need) __nanosleep(100);
                        } while (seen < need);
                        seen_s[j * B + tid] = seen;
                    }
                }
                const float *src = a.qdom + (size_t)member * tp.total_ticks * B + x.base + (int64_t)d0 * B;
                for (int i = 0; i < nd; i++)
                    xq[(j * CHF + i) * B + tid] = (i >= i_lo && i <= i_hi) ? __ldcg(src + (size_t)i * B) : 0.0f;
                j++;
            }
        }
        {
            const float *qp = qts + tid;
            float *tr = tape;
            float *qd = qdom, *np_ = netp;
            int t = d0 - off;
#pragma unroll 1
            for (int i = 0; i < nd; i++) {
                const int d = d0 + i;
                const bool act = valid && (unsigned)t < (unsigned)T;
                const float *qprev = (d & 1) ? qbuf0 : qbuf1;
                float *qcur = (d & 1) ? qbuf1 : qbuf0;
                const float qt = qp[0];
                float q = 0.0f, qup = 0.0f;
                float xv[8];
                if (simple) {
                    const char *qb = reinterpret_cast<const char *>(qprev);
                    const float xval = MULTI ? xq[i * B + tid] : 0.0f;      // prefetched cross-block inflow (0 if none)
                    qup = ((((MULTI && xfirst) ? xval : 0.0f) + *reinterpret_cast<const float *>(qb + u0)) +
                           *reinterpret_cast<const float *>(qb + u1)) +
                          *reinterpret_cast<const float *>(qb + u2);          // upstream_discharge md_routing_operator.f90:37-53
                    if (many)
                        qup = ((qup + qprev[ux & 1023u]) + qprev[(ux >> 10) & 1023u]) + qprev[ux >> 20];
                    if (MULTI && xlast) qup = qup + xval;
                } else if (act) {
                    int nx = 0;
                    for (int e = ub; e < ue; e++) {                          // cross-block inflows
                        const int ea = tp.up[e].a;
                        if (ea < 0) {
                            float v;
                            if (nx < 2) v = xq[(nx * CHF + i) * B + tid];    // prefetched at the start of phase R
                            else {                                          // rare: third and later producer, read per tick
                                const ExtRef x = tp.ext[-ea - 1];
                                v = 0.0f;
                                if (t - x.lag >= 0) {
                                    const int need = d + x.dtick + 1;
                                    while (ld_acquire(prog_m + x.blk) < need) __nanosleep(40);
                                    v = __ldcg(a.qdom + (size_t)member * tp.total_ticks * B + x.base + (int64_t)d * B);
                                }
                            }
                            xv[nx & 7] = v;
                            nx++;
                        }
                    }
                    if (!late) {
                        int jx = 0;
                        for (int e = ub; e < ue; e++) {
                            const int ea = tp.up[e].a;
                            qup = qup + (ea >= 0 ? qprev[ea] : xv[(jx++) & 7]);
                        }
                    }
                }
                if (!late) {
                    qup = FAST ? qup * k.s_q : ((fa > 1) ? (qup * c_dt) / k.den : 0.0f);    // :55-56
                    // linear_routing md_routing_operator.f90:62-79
                    const float hr_imd = hlr + qup;
                    const float hlr_n = hr_imd * k.E;
                    const float qrout = hr_imd - hlr_n;
                    q = FAST ? fmaf(qrout, k.fa1, qt) * k.c0
                             : (qt + qrout * k.fa1) * c_dx * c_dx * 0.001f / c_dt;          // md_forward_structure.f90:155
                    if (act) {
                        if (TAPE) { tr[2 * B] = hlr; tr[3 * B] = qup; }
                        hlr = hlr_n;
                        qcur[tid] = q;
                    }
                }
                if (has_late) {
                    __syncthreads();
                    if (act && late) {
                        qup = 0.0f;
                        int jx = 0;
                        for (int e = ub; e < ue; e++) {
                            const UpEntry u = tp.up[e];
                            qup = qup + (u.a >= 0 ? (u.cur ? qcur : qprev)[u.a] : xv[(jx++) & 7]);
                        }
                        qup = FAST ? qup * k.s_q : (qup * c_dt) / k.den;
                        if (TAPE) { tr[2 * B] = hlr; tr[3 * B] = qup; }
                        const float hr_imd = hlr + qup;
                        hlr = hr_imd * k.E;
                        const float qrout = hr_imd - hlr;
                        q = FAST ? fmaf(qrout, k.fa1, qt) * k.c0 : (qt + qrout * k.fa1) * c_dx * c_dx * 0.001f / c_dt;
                        qcur[tid] = q;
                    }
                }
                if (act) {
                    if (publish) qd[0] = q;
                    if (do_net) np_[0] = qt;
                    if (gfirst >= 0)
                        for (int g = gfirst; g >= 0; g = tp.gauge_next[g]) qsim[(size_t)t * ng + g] = q;   // :206-210
                }
                __syncthreads();
                if (tid == 0) {
                    if (do_flag && i == nd - 1) { __threadfence(); st_release(prog_mine, d + 1); }   // one release per chunk
                    if (i == 0) {
                        // every thread is past phase V of this chunk: its forcing rows can be refilled for chunk + 2
                        for (int r = 0; r < CHF; r++) {
                            const int dn = d0 + FRING + r;
                            if (dn < nticks) {
                                mbar_expect_tx(&bars[sbase + r], row_bytes);
                                tma_load_1d(ring + (size_t)(sbase + r) * 2 * B, frc + (size_t)dn * 2 * B, row_bytes, &bars[sbase + r]);
                            }
                        }
                    }
                }
                qp += B; t++;
                if (publish) qd += B;
                if (do_net) np_ += B;
                if (TAPE) tr += 4 * B;
            }
        }
        if (publish) qdom += (size_t)CHF * B;
        if (do_net) netp += (size_t)CHF * B;
        if (TAPE) tape += (size_t)CHF * 4 * B;
        if (sbase) parity ^= 1u;
    }
    if (valid) {
        float *fs = a.fstates + (size_t)member * 3 * tp.nslots + slot;
        fs[0] = hp; fs[(size_t)tp.nslots] = hft; fs[(size_t)2 * tp.nslots] = hlr;
    }
}

// ------------------------------------------------------------------------------------------------
// reverse (adjoint) kernel: GR_A_FORWARD_B reverse sweep, forward_db.f90:8102-8174, hand-written.
// Carried backwards per cell: hp_b, hft_b, hlr_b; accumulated: cp_b, cft_b, exc_b, lr_b.
// Exchange between cells: w = s * qup_b of the downstream cell (UPSTREAM_DISCHARGE_B :6553-6558) read
// as a GATHER from the single downstream cell -- no atomics.
// ------------------------------------------------------------------------------------------------
template <int FAST>
__global__ void __launch_bounds__(512) reverse_kernel(const SolverArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const DeviceTopology &tp = a.tp;
    const int B = tp.B;
    const int tid = threadIdx.x;
    float *ring = reinterpret_cast<float *>(smem_raw);                 // [RING_STAGES][6][B]: prcp, pet, hp0, hft0, hlr0, qup
    float *wx = ring + RING_STAGES * 6 * B;                            // [2][B]
    uint64_t *bars = reinterpret_cast<uint64_t *>(wx + 2 * B);
    __shared__ unsigned int s_ticket;

    if (tid == 0) {
        s_ticket = atomicAdd(a.ticket, 1u);
        for (int s = 0; s < RING_STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    wx[tid] = 0.0f;
    wx[B + tid] = 0.0f;
    __syncthreads();
    const int vb = (int)s_ticket;
    const int member = vb / tp.nblocks;
    const int blk = tp.nblocks - 1 - (vb - member * tp.nblocks);       // reverse block order
    const int slot = blk * B + tid;
    const int nticks = tp.T + tp.hmax[blk];
    const int64_t row0 = tp.tick_base[blk];
    const unsigned bf = tp.bflags[blk];
    const bool has_pair = bf & BLK_LATE;
    const float *frc = a.forcing + row0 * 2 * B;
    const float *tpe = a.tape + ((size_t)member * tp.total_ticks + row0) * 4 * B;
    const uint32_t frc_bytes = 2u * B * sizeof(float), tape_bytes = 4u * B * sizeof(float);

    // ring slot j holds tick d = nticks - 1 - j
    if (tid == 0) {
        for (int j = 0; j < RING_STAGES && j < nticks; j++) {
            const int d = nticks - 1 - j;
            mbar_expect_tx(&bars[j], frc_bytes + tape_bytes);
            tma_load_1d(ring + (size_t)j * 6 * B, frc + (size_t)d * 2 * B, frc_bytes, &bars[j]);
            tma_load_1d(ring + (size_t)j * 6 * B + 2 * B, tpe + (size_t)d * 4 * B, tape_bytes, &bars[j]);
        }
    }

    const bool valid = tp.cell[slot] >= 0;
    const int off = tp.off[slot];
    const int fa = tp.flwacc[slot];
    const bool early = tp.early[slot];
    const int dkind = tp.down_kind[slot], dlane = tp.down_lane[slot];
    const int gfirst = tp.gauge_first[slot];
    const float *fld = a.fields + (size_t)member * NFIELD * tp.nslots + slot;
    CellConst k = make_const(1.0f, 1.0f, 0.0f, 1.0f, 1, a.dt, a.dx);
    if (valid)
        k = make_const(fld[(size_t)F_CP * tp.nslots], fld[(size_t)F_CFT * tp.nslots], fld[(size_t)F_EXC * tp.nslots],
                       fld[(size_t)F_LR * tp.nslots], fa, a.dt, a.dx);
    ExtRef rx = {0, 0, 0, 0, 0};
    if (dkind == 3) rx = tp.rext[dlane];
    const float c_dx = a.dx, c_dt = a.dt;
    const float s_w = (fa > 1) ? c_dt / k.den : 0.0f;                    // forward_db.f90:6547
    float hp_b = 0.0f, hft_b = 0.0f, hlr_b = 0.0f, cp_b = 0.0f, cft_b = 0.0f, exc_b = 0.0f, lr_b = 0.0f;
    const float *qsim_b = a.qsim_b + (size_t)member * tp.T * tp.ng;
    float *wdom = a.wdom + ((size_t)member * tp.total_ticks + row0) * B + tid;
    const int *rprog_m = a.rprog + (size_t)member * tp.nblocks;

    for (int j = 0; j < nticks; j++) {
        const int d = nticks - 1 - j;
        const int stage = j % RING_STAGES;
        mbar_wait(&bars[stage], (j / RING_STAGES) & 1);
        const float *rw = ring + (size_t)stage * 6 * B + tid;
        const float prcp = rw[0], pet = rw[B], hp0 = rw[2 * B], hft0 = rw[3 * B], hlr0 = rw[4 * B], qup = rw[5 * B];
        const int t = d - off;
        const bool act = valid && t >= 0 && t < tp.T;
        const int cur = j & 1;
        float w_out = 0.0f;

        for (int phase = 0; phase < (has_pair ? 2 : 1); phase++) {
            if (phase == 1) __syncthreads();
            if (!(act && ((early && has_pair) ? phase == 1 : phase == 0))) continue;
            // adjoint of the discharge of this cell-step
            float q_b = 0.0f;
            for (int g = gfirst; g >= 0; g = tp.gauge_next[g]) q_b += qsim_b[(size_t)t * tp.ng + g];   // :8104-8108
            if (dkind == 1) q_b += wx[(cur ^ 1) * B + dlane];
            else if (dkind == 2) q_b += wx[cur * B + dlane];
            else if (dkind == 3) {
                if (t + rx.lag < tp.T) {
                    const int dp = d + rx.dtick;
                    while (ld_acquire(rprog_m + rx.blk) > dp) __nanosleep(40);
                    q_b += __ldcg(a.wdom + (size_t)member * tp.total_ticks * B + rx.base + (int64_t)d * B);
                }
            }
            const float temp_b = fdiv<FAST>(c_dx * c_dx * 0.001f * q_b, c_dt);   // :8114
            const float qt_b = temp_b;
            const float qrout_b = k.fa1 * temp_b;                                 // :8118
            // LINEAR_ROUTING_B :6628-6652
            float hrb = hlr_b - qrout_b;
            const float hr_imd_b = qrout_b + k.E * hrb;
            const float arg1_b = k.E * (hlr0 + qup) * hrb;
            lr_b += fdiv<FAST>(c_dt * arg1_b, (k.lr * k.lr) * 60.0f);
            hlr_b = hr_imd_b;
            w_out = s_w * hr_imd_b;                                               // UPSTREAM_DISCHARGE_B :6547
            vertical_step_b<FAST>(k, prcp, pet, hp0, hft0, qt_b, hp_b, hft_b, cp_b, cft_b, exc_b);
            wx[cur * B + tid] = w_out;
        }
        if (act && (bf & BLK_RPUBLISH)) wdom[(size_t)d * B] = w_out;
        __syncthreads();
        if (tid == 0) {
            if (bf & BLK_RPUBLISH) { __threadfence(); st_release(a.rprog + (size_t)member * tp.nblocks + blk, d); }
            const int jn = j + RING_STAGES;
            if (jn < nticks) {
                const int dn = nticks - 1 - jn;
                mbar_expect_tx(&bars[stage], frc_bytes + tape_bytes);
                tma_load_1d(ring + (size_t)stage * 6 * B, frc + (size_t)dn * 2 * B, frc_bytes, &bars[stage]);
                tma_load_1d(ring + (size_t)stage * 6 * B + 2 * B, tpe + (size_t)dn * 4 * B, tape_bytes, &bars[stage]);
            }
        }
    }
    if (valid) {
        float *g = a.grad + (size_t)member * NFIELD * tp.nslots + slot;
        g[(size_t)F_CP * tp.nslots] = cp_b; g[(size_t)F_CFT * tp.nslots] = cft_b; g[(size_t)F_EXC * tp.nslots] = exc_b;
        g[(size_t)F_LR * tp.nslots] = lr_b; g[(size_t)F_HP * tp.nslots] = hp_b; g[(size_t)F_HFT * tp.nslots] = hft_b;
        g[(size_t)F_HLR * tp.nslots] = hlr_b;
    }
}

// ------------------------------------------------------------------------------------------------
// cost kernel: compute_jobs (mwd_cost.f90:37-156) + COMPUTE_JOBS_B (forward_db.f90:2553-2715).
// One CTA per member, one thread per gauge, sequential sums in the reference's order; no FMA contraction
// (explicit _rn intrinsics) so that the sums round like the scalar Fortran.
// ------------------------------------------------------------------------------------------------
#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
#define FD(a, b) __fdiv_rn((a), (b))

struct Series {
    const float *qsim, *qobs;
    int ng, g, start, n;
    float fs, fo;   // qs = qsim*dt/area*1e3 ; qo = qobs*dt/(flwacc*dx*dx)*1e3
    float dt, area, dden;
    __device__ float qs(int i) const { return FM(FD(FM(qsim[(size_t)(start + i) * ng + g], dt), area), 1e3f); }
    __device__ float qo(int i) const { return FM(FD(FM(qobs[(size_t)(start + i) * ng + g], dt), dden), 1e3f); }
};

struct Moments { int n; float sum_x, sum_y, sum_xx, sum_yy, sum_xy, se, lg; };

// sums over one chunk of precomputed (x, y) = (qo, qs) values, in the reference's order (mwd_cost.f90:350-490)
__device__ void moments_acc(Moments &m, const float *xs, const float *ys, int len) {
    for (int i = 0; i < len; i++) {
        const float x = xs[i], y = ys[i];
        if (x >= 0.0f) {
            m.n++;
            m.sum_x = FA(m.sum_x, x);
            m.sum_y = FA(m.sum_y, y);
            m.sum_xx = FA(m.sum_xx, FM(x, x));
            m.sum_yy = FA(m.sum_yy, FM(y, y));
            m.sum_xy = FA(m.sum_xy, FM(x, y));
            const float df = FS(x, y);
            m.se = FA(m.se, FM(df, df));
        }
        if (x > 0.0f && y > 0.0f) { const float lq = logf(FD(y, x)); m.lg = FA(m.lg, FM(FM(x, lq), lq)); }
    }
}

__device__ float nse_of(const Moments &m) {          // mwd_cost.f90:350-401
    const float mean_x = FD(m.sum_x, (float)m.n);
    const float num = FA(FS(m.sum_xx, FM(2.0f, m.sum_xy)), m.sum_yy);
    const float den = FS(m.sum_xx, FM(FM((float)m.n, mean_x), mean_x));
    return FD(num, den);
}
struct Kge { float r, a, b, mean_x, mean_y, var_x, var_y, cov, val; };
__device__ Kge kge_of(const Moments &m) {            // mwd_cost.f90:403-490
    Kge k;
    const float n = (float)m.n;
    k.mean_x = FD(m.sum_x, n); k.mean_y = FD(m.sum_y, n);
    k.var_x = FS(FD(m.sum_xx, n), FM(k.mean_x, k.mean_x));
    k.var_y = FS(FD(m.sum_yy, n), FM(k.mean_y, k.mean_y));
    k.cov = FS(FD(m.sum_xy, n), FM(k.mean_x, k.mean_y));
    k.r = FD(FD(k.cov, sqrtf(k.var_x)), sqrtf(k.var_y));
    k.a = FD(sqrtf(k.var_y), sqrtf(k.var_x));
    k.b = FD(k.mean_y, k.mean_x);
    const float r1 = FS(k.r, 1.0f), b1 = FS(k.b, 1.0f), a1 = FS(k.a, 1.0f);
    k.val = sqrtf(FA(FA(FM(r1, r1), FM(b1, b1)), FM(a1, a1)));
    return k;
}

// One CTA per member, one thread per gauge for the order-dependent sums.  The element-wise part (unit conversion with
// its IEEE divisions) is done by the whole CTA into shared memory, chunk by chunk, so the sequential part is adds only.
// ---- signature objectives (mwd_cost.f90:770-970): |s(qs) / s(qo) - 1| per gauge, one thread per gauge, straight from
// global memory (a few thousand steps).  Forward only.
struct SigSeries {           // qo / qs / po of one gauge from step `start` on (mwd_cost.f90:82-92)
    const float *qsim, *qobs, *mprcp;
    const int32_t *mask;
    int ng, g, start;
    float dt, area, dden;
    __device__ float qs(int i) const { return FM(FD(FM(qsim[(size_t)(start + i) * ng + g], dt), area), 1e3f); }
    __device__ float qo(int i) const { return FM(FD(FM(qobs[(size_t)(start + i) * ng + g], dt), dden), 1e3f); }
    __device__ float po(int i) const { return mprcp[(size_t)(start + i) * ng + g]; }
    __device__ int ev(int i) const { return mask[(size_t)(start + i) * ng + g]; }
};
__device__ void sig_heap_sort(float *a, int n) {             // heap_sort, mwd_cost.f90:594-673 (any correct sort gives the same array)
    for (int start = n / 2 - 1; start >= 0; start--) {
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && a[child] < a[child + 1]) child++;
            if (a[root] >= a[child]) break;
            const float t = a[root]; a[root] = a[child]; a[child] = t;
            root = child;
        }
    }
    for (int end = n - 1; end > 0; end--) {
        const float t0 = a[0]; a[0] = a[end]; a[end] = t0;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && a[child] < a[child + 1]) child++;
            if (a[root] >= a[child]) break;
            const float t = a[root]; a[root] = a[child]; a[child] = t;
            root = child;
        }
    }
}
__device__ float sig_quantile(float *dat, int n, float p) {  // quantile, mwd_cost.f90:675-720 (sorts dat in place)
    if (n <= 0) return 0.0f;
    if (n == 1) return dat[0];
    sig_heap_sort(dat, n);
    const float frac = FA(FM((float)(n - 1), p), 1.0f);
    if (frac <= 1.0f) return dat[0];
    if (frac >= (float)n) return dat[n - 1];
    const float q1 = dat[(int)frac - 1], q2 = dat[(int)frac];
    return FA(q1, FM(FS(q2, q1), FS(frac, (float)(int)frac)));
}
__device__ float signature_dev(const SigSeries &S, int n, int stype, float *scratch /* 2 * n floats */) {
    float res = 0.0f, num = 0.0f, den = 0.0f;
    if (stype >= 12) {                                       // Erc, Elt, Epf: per event of mask_event (:799-893)
        int n_event = 0;
        for (int i = n - 1; i >= 0; i--) if (S.ev(i) > 0) { n_event = S.ev(i); break; }
        for (int e = 1; e <= n_event; e++) {
            int start = -1, cnt = 0;
            for (int j = 0; j < n; j++) if (S.ev(j) == e) { if (start < 0) start = j; cnt++; }
            if (start < 0) start = 0;
            float sum_qo = 0.f, sum_qs = 0.f, sum_po = 0.f, max_qo = 0.f, max_qs = 0.f, max_po = 0.f;
            int imax_qo = 0, imax_qs = 0, imax_po = 0;
            for (int j = start; j < start + cnt && j < n; j++) {
                const float qo = S.qo(j), qs = S.qs(j), po = S.po(j);
                if (qo >= 0.0f && po >= 0.0f) {
                    sum_qo = FA(sum_qo, qo); sum_qs = FA(sum_qs, qs); sum_po = FA(sum_po, po);
                    if (qo > max_qo) { max_qo = qo; imax_qo = j + 1; }
                    if (qs > max_qs) { max_qs = qs; imax_qs = j + 1; }
                    if (po > max_po) { max_po = po; imax_po = j + 1; }
                }
            }
            if (stype == 14) { num = max_qs; den = max_qo; }
            else if (stype == 13) { num = (float)(imax_qs - imax_po); den = (float)(imax_qo - imax_po); }
            else if (sum_po > 0.0f) { num = FD(sum_qs, sum_po); den = FD(sum_qo, sum_po); }
            if (den > 0.0f) res = FA(res, fabsf(FS(FD(num, den), 1.0f)));
        }
        if (n_event > 0) res = FD(res, (float)n_event);
    } else {
        if (stype == 7) {                                    // Crc (:897-918)
            float sum_qo = 0.f, sum_qs = 0.f, sum_po = 0.f;
            for (int i = 0; i < n; i++) {
                const float qo = S.qo(i), po = S.po(i);
                if (qo >= 0.0f && po >= 0.0f) { sum_qo = FA(sum_qo, qo); sum_qs = FA(sum_qs, S.qs(i)); sum_po = FA(sum_po, po); }
            }
            if (sum_po > 0.0f) { num = FD(sum_qs, sum_po); den = FD(sum_qo, sum_po); }
        } else if (scratch) {                                // Cfp2 / 10 / 50 / 90: flow_percentile (:722-768)
            const float p = stype == 8 ? 0.02f : stype == 9 ? 0.1f : stype == 10 ? 0.5f : 0.9f;
            float *a = scratch, *b = scratch + n;
            int j = 0;
            for (int i = 0; i < n; i++) {
                const float qo = S.qo(i), qs = S.qs(i);
                if (qo >= 0.0f && qs >= 0.0f) { a[j] = qo; b[j] = qs; j++; }
            }
            num = sig_quantile(b, j, p);
            den = sig_quantile(a, j, p);
        }
        if (den > 0.0f) res = fabsf(FS(FD(num, den), 1.0f));
    }
    return res;
}

__global__ void cost_kernel(const CostArgs c, const int CH) {
    extern __shared__ float sh[];            // [ng] gauge_jobs, [ng] gauge_jobs_b, [ng][7] adjoint coefficients, [ng][CH] x, [ng][CH] y
    float *gj = sh, *gjb = sh + c.ng, *coef = sh + 2 * c.ng, *xs = sh + 9 * c.ng, *ys = xs + (size_t)c.ng * CH;
    const int m = blockIdx.x;
    const int n = c.T - c.start;
    const int g = threadIdx.x;               // launch_cost: blockDim.x >= ng
    const float *qsim = c.qsim + (size_t)m * c.T * c.ng;
    const bool mine = g < c.ng && (c.wgauge[g] > 0.0f || c.wgauge[g] < 0.0f);
    Moments mo = {0, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < n; c0 += CH) {
        const int len = min(CH, n - c0);
        for (int idx = threadIdx.x; idx < len * c.ng; idx += blockDim.x) {
            const int i = idx / c.ng, gg = idx - i * c.ng;
            const size_t at = (size_t)(c.start + c0 + i) * c.ng + gg;
            const float dden = FM(FM((float)c.gauge_flwacc[gg], c.dx), c.dx);
            xs[(size_t)gg * CH + i] = FM(FD(FM(c.qobs[at], c.dt), dden), 1e3f);        // qo, mwd_cost.f90:88-92
            ys[(size_t)gg * CH + i] = FM(FD(FM(qsim[at], c.dt), c.area[gg]), 1e3f);   // qs, :84-86
        }
        __syncthreads();
        if (mine) moments_acc(mo, xs + (size_t)g * CH, ys + (size_t)g * CH, len);
        __syncthreads();
    }
    if (g < c.ng) {
        float gauge_jobs = 0.0f;
        if (mine) {
            float j_imd = 0.0f;
            for (int j = 0; j < c.njf; j++) {
                if (mo.n > 0) {
                    switch (c.jobs_fun[j]) {
                        case 1: j_imd = nse_of(mo); break;
                        case 2: j_imd = kge_of(mo).val; break;
                        case 3: { const float v = kge_of(mo).val; j_imd = FM(v, v); } break;
                        case 4: j_imd = mo.se; break;
                        case 5: j_imd = sqrtf(FD(mo.se, (float)mo.n)); break;
                        case 6: j_imd = mo.lg; break;
                        default:
                            if (c.jobs_fun[j] >= 7 && c.jobs_fun[j] <= 14 && c.mean_prcp && c.mask_event) {
                                const SigSeries S{qsim, c.qobs, c.mean_prcp, c.mask_event, c.ng, g, c.start, c.dt, c.area[g],
                                                  FM(FM((float)c.gauge_flwacc[g], c.dx), c.dx)};
                                j_imd = signature_dev(S, n, c.jobs_fun[j], c.scratch ? c.scratch + ((size_t)m * c.ng + g) * 2 * c.T : nullptr);
                            }
                            break;
                    }
                }
                gauge_jobs = FA(gauge_jobs, FM(c.wjobs_fun[j], j_imd));
            }
        }
        gj[g] = gauge_jobs;
        gjb[g] = 0.0f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // gauge aggregation mwd_cost.f90:139-154 (+ QUANTILE_B for negative weights)
        float jobs = 0.0f;
        int nmed = 0;
        for (int g = 0; g < c.ng; g++) {
            const float wg = c.wgauge[g];
            if (wg > 0.0f) { jobs = FA(jobs, FM(wg, gj[g])); gjb[g] = FM(wg, c.jobs_b); }
            else if (wg < 0.0f) nmed++;
        }
        if (nmed > 0) {
            // median by selection (n is tiny): rank of each value with index tie-break
            const float frac = FA(FM((float)(nmed - 1), 0.5f), 1.0f);
            int i1 = (int)frac - 1, i2 = i1 + 1;
            float w2 = FS(frac, (float)(int)frac);
            if (nmed == 1 || frac <= 1.0f) { i1 = i2 = 0; w2 = 0.0f; }
            else if (frac >= (float)nmed) { i1 = i2 = nmed - 1; w2 = 0.0f; }
            int g1 = -1, g2 = -1;
            for (int g = 0; g < c.ng; g++) {
                if (!(c.wgauge[g] < 0.0f)) continue;
                int rk = 0;
                for (int h = 0; h < c.ng; h++)
                    if (c.wgauge[h] < 0.0f && (gj[h] < gj[g] || (gj[h] == gj[g] && h < g))) rk++;
                if (rk == i1) g1 = g;
                if (rk == i2) g2 = g;
            }
            jobs = FA(gj[g1], FM(FS(gj[g2], gj[g1]), w2));
            for (int g = 0; g < c.ng; g++) if (c.wgauge[g] < 0.0f) gjb[g] = 0.0f;
            gjb[g1] = FA(gjb[g1], FM(FS(1.0f, w2), c.jobs_b));
            gjb[g2] = FA(gjb[g2], FM(w2, c.jobs_b));
        }
        c.cost_jobs[m] = jobs;
    }
    if (c.qsim_b == nullptr) return;
    __syncthreads();
    if (g < c.ng) {
        float *cf = coef + 7 * g;
        for (int k = 0; k < 7; k++) cf[k] = 0.0f;
    }
    if (mine && mo.n > 0) {
        const float gauge_jobs_b = gjb[g];
        // accumulate y_b(i) = cx*x(i) + cy*y(i) + c0 + (se / log terms)
        float cx = 0.f, cy = 0.f, c0 = 0.f, cse = 0.f, clg = 0.f;
        for (int j = c.njf - 1; j >= 0; j--) {
            const float jb = FM(c.wjobs_fun[j], gauge_jobs_b);
            switch (c.jobs_fun[j]) {
                case 1: {   // NSE_B forward_db.f90:3505-3545
                    const float mean_x = FD(mo.sum_x, (float)mo.n);
                    const float den = FS(mo.sum_xx, FM(FM((float)mo.n, mean_x), mean_x));
                    const float num_b = FD(jb, den);
                    cx = FA(cx, -FM(2.0f, num_b)); cy = FA(cy, FM(2.0f, num_b));
                } break;
                case 2: case 3: {   // KGE_B :3805-3829 + KGE_COMPONENTS_B :3657-3729
                    const Kge k = kge_of(mo);
                    const float res_b = (c.jobs_fun[j] == 3) ? FM(FM(2.0f, k.val), jb) : jb;
                    const float r1 = FS(k.r, 1.0f), b1 = FS(k.b, 1.0f), a1 = FS(k.a, 1.0f);
                    const float arg1 = FA(FA(FM(r1, r1), FM(b1, b1)), FM(a1, a1));
                    const float arg1_b = (arg1 == 0.0f) ? 0.0f : FD(res_b, FM(2.0f, sqrtf(arg1)));
                    const float r_b = FM(FM(2.0f, r1), arg1_b), b_b = FM(FM(2.0f, b1), arg1_b), a_b = FM(FM(2.0f, a1), arg1_b);
                    const float sx = sqrtf(k.var_x), sy = sqrtf(k.var_y);
                    const float result1_b = FD(a_b, sx);
                    float var_y_b = (k.var_y == 0.0f) ? 0.0f : FD(result1_b, FM(2.0f, sy));
                    const float temp_b = FD(r_b, FM(sx, sy));
                    const float cov_b = temp_b;
                    const float result2_b = -FD(FM(k.cov, temp_b), sy);
                    if (!(k.var_y == 0.0f)) var_y_b = FA(var_y_b, FD(result2_b, FM(2.0f, sy)));
                    const float mean_y_b = FS(FS(FD(b_b, k.mean_x), FM(k.mean_x, cov_b)), FM(FM(2.0f, k.mean_y), var_y_b));
                    const float nn = (float)mo.n;
                    cx = FA(cx, FD(cov_b, nn)); cy = FA(cy, FM(2.0f, FD(var_y_b, nn))); c0 = FA(c0, FD(mean_y_b, nn));
                } break;
                case 4: cse = FA(cse, jb); break;
                case 5: {
                    const float sm = FD(mo.se, (float)mo.n);
                    const float sb = (sm == 0.0f) ? 0.0f : FD(jb, FM(2.0f, sqrtf(sm)));
                    cse = FA(cse, FD(sb, (float)mo.n));
                } break;
                case 6: clg = FA(clg, jb); break;
                default: break;
            }
        }
        float *cf = coef + 7 * g;
        cf[0] = cx; cf[1] = cy; cf[2] = c0; cf[3] = cse; cf[4] = clg; cf[5] = FD(FM(c.dt, 1e3f), c.area[g]); cf[6] = 1.0f;   // :2709-2712
    }
    __syncthreads();
    float *qb = c.qsim_b + (size_t)m * c.T * c.ng;
    for (int idx = threadIdx.x; idx < c.T * c.ng; idx += blockDim.x) {
        const int t = idx / c.ng, gg = idx - t * c.ng;
        const float *cf = coef + 7 * gg;
        float out = 0.0f;
        if (t >= c.start && cf[6] != 0.0f) {
            const float dden = FM(FM((float)c.gauge_flwacc[gg], c.dx), c.dx);
            const float x = FM(FD(FM(c.qobs[idx], c.dt), dden), 1e3f), y = FM(FD(FM(qsim[idx], c.dt), c.area[gg]), 1e3f);
            float yb = 0.0f;
            if (x >= 0.0f) yb = FA(FA(FM(x, cf[0]), FM(y, cf[1])), cf[2]) - FM(FM(2.0f, FS(x, y)), cf[3]);
            if (cf[4] != 0.0f && x > 0.0f && y > 0.0f) yb = FA(yb, FD(FM(FM(FM(2.0f, x), logf(FD(y, x))), cf[4]), y));
            out = FM(cf[5], yb);
        }
        qb[idx] = out;
    }
}

// ------------------------------------------------------------------------------------------------
// layout kernels
// ------------------------------------------------------------------------------------------------
__global__ void relayout_forcing_kernel(DeviceTopology tp, const int32_t *src, const float *prcp_raw, const float *pet_raw,
                                        int64_t stride, float *forcing) {
    const int blk = blockIdx.x;
    const int B = tp.B;
    const int nticks = tp.T + tp.hmax[blk];
    const int64_t row0 = tp.tick_base[blk];
    for (int lane = threadIdx.x; lane < B; lane += blockDim.x) {
        const int slot = blk * B + lane;
        const int sidx = src[slot];
        const int off = tp.off[slot];
        for (int d = blockIdx.y; d < nticks; d += gridDim.y) {
            const int t = d - off;
            float p = 0.0f, e = 0.0f;
            if (sidx >= 0 && t >= 0 && t < tp.T) { p = prcp_raw[(int64_t)t * stride + sidx]; e = pet_raw[(int64_t)t * stride + sidx]; }
            forcing[(row0 + d) * 2 * B + lane] = p;
            forcing[(row0 + d) * 2 * B + B + lane] = e;
        }
    }
}

__global__ void gather_fields_kernel(DeviceTopology tp, int nmember, const float *planes, int64_t ncell, const float *sample,
                                     const int32_t *sample_field, int nvar, float *fields) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)nmember * NFIELD * tp.nslots;
    if (i >= total) return;
    const int slot = (int)(i % tp.nslots);
    const int f = (int)((i / tp.nslots) % NFIELD);
    const int m = (int)(i / ((int64_t)tp.nslots * NFIELD));
    const int c = tp.cell[slot];
    float v = (c >= 0) ? planes[(int64_t)f * ncell + c] : 1.0f;
    for (int j = 0; j < nvar; j++) if (sample_field[j] == f) v = sample[(int64_t)m * nvar + j];
    fields[i] = v;
}

__global__ void unskew_kernel(DeviceTopology tp, const int32_t *dst, const float *skewed, int64_t out_stride, float *out) {
    const int blk = blockIdx.x;
    const int B = tp.B;
    const int64_t row0 = tp.tick_base[blk];
    for (int lane = threadIdx.x; lane < B; lane += blockDim.x) {
        const int slot = blk * B + lane;
        const int di = dst[slot];
        if (di < 0) continue;
        const int off = tp.off[slot];
        for (int t = blockIdx.y; t < tp.T; t += gridDim.y) out[(int64_t)t * out_stride + di] = skewed[(row0 + t + off) * B + lane];
    }
}

__global__ void checksum_kernel(DeviceTopology tp, const float *skewed, double *out) {
    const int blk = blockIdx.x;
    const int B = tp.B;
    const int64_t row0 = tp.tick_base[blk];
    double acc = 0.0;
    for (int lane = threadIdx.x; lane < B; lane += blockDim.x) {
        const int slot = blk * B + lane;
        if (tp.cell[slot] < 0) continue;
        const int off = tp.off[slot];
        for (int t = blockIdx.y; t < tp.T; t += gridDim.y) acc += (double)skewed[(row0 + t + off) * B + lane];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// ------------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------------
static size_t fwd_smem(int B, bool multi) { return (size_t)((FRING * 2 + CHF + 2 + (multi ? 2 * CHF + 2 : 0)) * B + 2 * QX_PAD) * sizeof(float) + FRING * sizeof(uint64_t); }
static size_t rev_smem(int B) { return (size_t)(RING_STAGES * 6 + 2) * B * sizeof(float) + RING_STAGES * sizeof(uint64_t); }

template <typename K> static cudaError_t launch_solver(K kern, const SolverArgs &a, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(a.ticket, 0, sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((size_t)a.tp.nblocks * a.nmember));
    kern<<<grid, a.tp.B, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_forward(const SolverArgs &a, int math_mode, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(a.prog, 0, sizeof(int) * (size_t)a.tp.nblocks * a.nmember, s);
    if (e != cudaSuccess) return e;
    const bool multi = a.tp.nblocks > 1;
    const size_t smem = fwd_smem(a.tp.B, multi);
    if (multi) {
        if (a.tape_on) return math_mode ? launch_solver(forward_kernel<1, 1, 1>, a, smem, s) : launch_solver(forward_kernel<0, 1, 1>, a, smem, s);
        return math_mode ? launch_solver(forward_kernel<1, 0, 1>, a, smem, s) : launch_solver(forward_kernel<0, 0, 1>, a, smem, s);
    }
    if (a.tape_on) return math_mode ? launch_solver(forward_kernel<1, 1, 0>, a, smem, s) : launch_solver(forward_kernel<0, 1, 0>, a, smem, s);
    return math_mode ? launch_solver(forward_kernel<1, 0, 0>, a, smem, s) : launch_solver(forward_kernel<0, 0, 0>, a, smem, s);
}

cudaError_t launch_reverse(const SolverArgs &a, int math_mode, cudaStream_t s) {
    // rprog starts above every tick index
    cudaError_t e = cudaMemsetAsync(a.rprog, 0x7f, sizeof(int) * (size_t)a.tp.nblocks * a.nmember, s);
    if (e != cudaSuccess) return e;
    const size_t smem = rev_smem(a.tp.B);
    return math_mode ? launch_solver(reverse_kernel<1>, a, smem, s) : launch_solver(reverse_kernel<0>, a, smem, s);
}

cudaError_t launch_cost(const CostArgs &c, cudaStream_t s) {
    if (c.ng <= 0 || c.nmember <= 0) return cudaSuccess;
    if (c.ng > 1024) return cudaErrorInvalidValue;           // one thread per gauge
    const int threads = c.ng <= 128 ? 128 : ((c.ng + 31) / 32) * 32;
    int ch = (int)((40960 / sizeof(float) - 9 * (size_t)c.ng) / (2 * (size_t)c.ng));
    ch = ch > 2048 ? 2048 : (ch < 32 ? 32 : ch / 32 * 32);
    const size_t smem = (9 * (size_t)c.ng + 2 * (size_t)c.ng * ch) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    cost_kernel<<<c.nmember, threads, smem, s>>>(c, ch);
    return cudaGetLastError();
}

cudaError_t launch_relayout_forcing(const DeviceTopology &tp, const int32_t *src_index, const float *prcp_raw,
                                    const float *pet_raw, int64_t raw_stride, float *forcing, cudaStream_t s) {
    dim3 grid(tp.nblocks, tp.nblocks > 512 ? 8 : 64);
    relayout_forcing_kernel<<<grid, tp.B > 256 ? 256 : tp.B, 0, s>>>(tp, src_index, prcp_raw, pet_raw, raw_stride, forcing);
    return cudaGetLastError();
}

cudaError_t launch_gather_fields(const DeviceTopology &tp, int nmember, const float *planes, int64_t ncell, const float *sample,
                                 const int32_t *sample_field, int nvar, float *fields, cudaStream_t s) {
    const int64_t total = (int64_t)nmember * NFIELD * tp.nslots;
    gather_fields_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(tp, nmember, planes, ncell, sample, sample_field, nvar, fields);
    return cudaGetLastError();
}

cudaError_t launch_unskew(const DeviceTopology &tp, const int32_t *dst_index, const float *skewed, int64_t out_stride,
                          float fill, float *out, cudaStream_t s) {
    (void)fill;
    dim3 grid(tp.nblocks, tp.nblocks > 512 ? 8 : 64);
    unskew_kernel<<<grid, tp.B > 256 ? 256 : tp.B, 0, s>>>(tp, dst_index, skewed, out_stride, out);
    return cudaGetLastError();
}

cudaError_t launch_checksum(const DeviceTopology &tp, const float *skewed, double *out, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), s);
    if (e != cudaSuccess) return e;
    dim3 grid(tp.nblocks, tp.nblocks > 512 ? 8 : 64);
    checksum_kernel<<<grid, 256, 0, s>>>(tp, skewed, out);
    return cudaGetLastError();
}

}  // namespace smash
