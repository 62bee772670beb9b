/*
 * smash_oracle.c -- CPU ORACLE (test infrastructure, NOT the product; see smash_oracle.h).
 *
 * Restates, statement by statement and in the same floating-point evaluation order, the
 * reference's gr-a forward solver, its cost function and the Tapenade reverse sweep, and the
 * forward solvers of gr-b, gr-c, gr-d and vic-a (parity of those four: UNPINNED, see smash_oracle.h).
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/smash/solver/).  Compile with -ffp-contract=off so that no FMA is formed
 * that the scalar Fortran would not have formed.
 *
 * Defined behaviour where the reference has none (SURVEY.md 7 "Undefined behaviour"):
 *   - the automatic array q of gr_a_forward (md_forward_structure.f90:48) is zero-initialised;
 *   - j_imd of compute_jobs (mwd_cost.f90:66) starts at 0.
 */
#include "smash_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if ORACLE_DOUBLE
#define R(x) x
#define POW pow
#define TANH tanh
#define EXP exp
#define SQRT sqrt
#define LOG log
#define FABS fabs
#else
#define R(x) x##f
#define POW powf
#define TANH tanhf
#define EXP expf
#define SQRT sqrtf
#define LOG logf
#define FABS fabsf
#endif

typedef OSYM(OProblem) Prob;

#define IDX(P, row, col) ((size_t)((row)-1) + (size_t)((col)-1) * (size_t)(P)->nrow)

static inline oreal pow4i(oreal x) { /* x**4, integer exponent: gfortran expands to (x*x)*(x*x) */
    oreal x2 = x * x;
    return x2 * x2;
}

/* ============================================================================================
 * GR operators -- operator/md_gr_operator.f90
 * ========================================================================================== */

/* md_gr_operator.f90:36-67 */
static void gr_production(oreal pn, oreal en, oreal cp, oreal beta, oreal *hp, oreal *pr, oreal *perc) {
    oreal inv_cp = R(1.0) / cp;
    *pr = R(0.0);
    oreal ps = cp * (R(1.0) - (*hp) * (*hp)) * TANH(pn * inv_cp) / (R(1.0) + (*hp) * TANH(pn * inv_cp));
    oreal es = ((*hp) * cp) * (R(2.0) - (*hp)) * TANH(en * inv_cp) /
               (R(1.0) + (R(1.0) - (*hp)) * TANH(en * inv_cp));
    oreal hp_imd = (*hp) + (ps - es) * inv_cp;
    if (pn > 0) *pr = pn - (hp_imd - (*hp)) * cp;
    *perc = (hp_imd * cp) * (R(1.0) - POW(R(1.0) + pow4i(hp_imd / beta), R(-0.25)));
    *hp = hp_imd - (*perc) * inv_cp;
}

/* md_gr_operator.f90:69-79 */
static void gr_exchange(oreal exc, oreal hft, oreal *l) { *l = exc * POW(hft, R(3.5)); }

/* md_gr_operator.f90:81-110 */
static void gr_transfer(oreal n, oreal prcp, oreal pr, oreal ct, oreal *ht, oreal *q) {
    oreal nm1 = n - R(1.0);
    oreal d1pnm1 = R(1.0) / nm1;
    oreal pr_imd;
    if (prcp < R(0.0))
        pr_imd = POW(POW((*ht) * ct, -nm1) - POW(ct, -nm1), -d1pnm1) - ((*ht) * ct);
    else
        pr_imd = pr;
    oreal ht_imd = fmax(R(1.e-6), (*ht) + pr_imd / ct);
    *ht = POW(POW(ht_imd * ct, -nm1) + POW(ct, -nm1), -d1pnm1) / ct;
    *q = (ht_imd - (*ht)) * ct;
}

void OSYM(oracle_gr_production)(oreal pn, oreal en, oreal cp, oreal beta, oreal *hp, oreal *pr, oreal *perc) {
    gr_production(pn, en, cp, beta, hp, pr, perc);
}
void OSYM(oracle_gr_transfer)(oreal n, oreal prcp, oreal pr, oreal ct, oreal *ht, oreal *q) {
    gr_transfer(n, prcp, pr, ct, ht, q);
}

/* ============================================================================================
 * Routing operators -- operator/md_routing_operator.f90
 * ========================================================================================== */
static const int DCOL[8] = {0, -1, -1, -1, 0, 1, 1, 1}; /* md_routing_operator.f90:29 */
static const int DROW[8] = {1, 1, 0, -1, -1, -1, 0, 1}; /* md_routing_operator.f90:30 */

/* md_routing_operator.f90:17-60 */
static void upstream_discharge(const Prob *P, int row, int col, const oreal *q, oreal *qup) {
    *qup = R(0.0);
    int fa = P->flwacc[IDX(P, row, col)];
    if (fa > 1) {
        for (int i = 0; i < 8; i++) {
            int col_imd = col + DCOL[i], row_imd = row + DROW[i];
            if (col_imd > 0 && col_imd <= P->ncol && row_imd > 0 && row_imd <= P->nrow) {
                if (P->flwdir[IDX(P, row_imd, col_imd)] == i + 1) *qup = *qup + q[IDX(P, row_imd, col_imd)];
            }
        }
        *qup = (*qup * P->dt) / (R(0.001) * P->dx * P->dx * (oreal)(fa - 1));
    }
}

/* md_routing_operator.f90:62-79 */
static void linear_routing(oreal dt, oreal qup, oreal lr, oreal *hr, oreal *qrout) {
    oreal hr_imd = *hr + qup;
    *hr = hr_imd * EXP(-dt / (lr * R(60.0)));
    *qrout = hr_imd - *hr;
}

/* ============================================================================================
 * gr_a_forward -- forward/md_forward_structure.f90:30-214
 * Tape: what GR_A_FORWARD_B pushes per active cell-step (forward_db.f90:8016-8098).
 * ========================================================================================== */
typedef struct {
    oreal *pn, *en, *hp0, *prr, *hft0, *qup, *hlr0;
    unsigned char *flags; /* bit0: not a forcing gap, bit1: (prd + l) > 0 */
    int *slot;            /* path position -> tape column, -1 if inactive */
    int ncolumns;
} Tape;

static int tape_alloc(Tape *tp, const Prob *P) {
    size_t n = (size_t)P->nrow * P->ncol;
    tp->slot = (int *)malloc(n * sizeof(int));
    int k = 0;
    for (size_t i = 0; i < n; i++) {
        int row = P->path[2 * i], col = P->path[2 * i + 1];
        tp->slot[i] = -1;
        if (row > 0 && col > 0) {
            size_t c = IDX(P, row, col);
            if (P->active_cell[c] == 1 && (!P->local_active_cell || P->local_active_cell[c] == 1)) tp->slot[i] = k++;
        }
    }
    tp->ncolumns = k;
    size_t m = (size_t)k * P->ntime_step;
    if (m == 0) m = 1;
    tp->pn = (oreal *)malloc(m * sizeof(oreal));
    tp->en = (oreal *)malloc(m * sizeof(oreal));
    tp->hp0 = (oreal *)malloc(m * sizeof(oreal));
    tp->prr = (oreal *)malloc(m * sizeof(oreal));
    tp->hft0 = (oreal *)malloc(m * sizeof(oreal));
    tp->qup = (oreal *)malloc(m * sizeof(oreal));
    tp->hlr0 = (oreal *)malloc(m * sizeof(oreal));
    tp->flags = (unsigned char *)malloc(m);
    return (tp->pn && tp->en && tp->hp0 && tp->prr && tp->hft0 && tp->qup && tp->hlr0 && tp->flags) ? 0 : -1;
}
static void tape_free(Tape *tp) {
    free(tp->pn); free(tp->en); free(tp->hp0); free(tp->prr); free(tp->hft0); free(tp->qup); free(tp->hlr0);
    free(tp->flags); free(tp->slot);
}

static inline void read_forcing(const Prob *P, int row, int col, int t, oreal *prcp, oreal *pet) {
    if (P->sparse_storage) { /* md_forward_structure.f90:94-97 */
        size_t k = (size_t)P->rowcol_to_ind_sparse[IDX(P, row, col)] - 1;
        *prcp = P->prcp[k + (size_t)P->nac * t];
        *pet = P->pet[k + (size_t)P->nac * t];
    } else { /* :101-102 */
        size_t o = IDX(P, row, col) + (size_t)P->nrow * P->ncol * t;
        *prcp = P->prcp[o];
        *pet = P->pet[o];
    }
}

/* grd != 0: the statements of gr_d_forward instead (md_forward_structure.f90:589-760: no exchange, prr = pr + perc :685,
 * qt = qr :689) -- the same loop, taped for GR_D_FORWARD_B */
static void gr_ad_forward(const Prob *P, const oreal *par, oreal *st, oreal *qsim, oreal *qdom, oreal *netp,
                          Tape *tape, int grd) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    const oreal *cp = par + OP_CP * ncell, *cft = par + OP_CFT * ncell, *exc = par + OP_EXC * ncell,
                *lr = par + OP_LR * ncell;
    oreal *hp = st + OS_HP * ncell, *hft = st + OS_HFT * ncell, *hlr = st + OS_HLR * ncell;
    oreal *q = (oreal *)calloc(ncell, sizeof(oreal));

    for (int t = 0; t < P->ntime_step; t++) {          /* :57 */
        for (size_t i = 0; i < ncell; i++) {           /* :59 */
            oreal ei, pn = 0, en = 0, pr = 0, perc = 0, l = 0, prr, prd, qr = 0, qd, qt, qup = 0, qrout = 0;
            int row = P->path[2 * i], col = P->path[2 * i + 1];
            if (!(row > 0 && col > 0)) continue;       /* :82 */
            size_t c = IDX(P, row, col);
            if (!(P->active_cell[c] == 1 && (!P->local_active_cell || P->local_active_cell[c] == 1))) continue; /* :92 */
            oreal prcp, pet;
            read_forcing(P, row, col, t, &prcp, &pet);
            int nogap = (prcp >= 0 && pet >= 0);       /* :106 */
            size_t tp = 0;
            if (tape) tp = (size_t)tape->slot[i] + (size_t)tape->ncolumns * t;
            if (tape) tape->hp0[tp] = hp[c];
            if (nogap) {
                ei = fmin(pet, prcp);                  /* :112 */
                pn = fmax(R(0.0), prcp - ei);          /* :114 */
                en = pet - ei;                         /* :116 */
                gr_production(pn, en, cp[c], R(1000.0), &hp[c], &pr, &perc); /* :122 */
                if (!grd) gr_exchange(exc[c], hft[c], &l); /* :129 */
            }
            if (grd) {
                prr = pr + perc;                       /* :685 */
                prd = R(0.0);
            } else {
                prr = R(0.9) * (pr + perc) + l;        /* :137 */
                prd = R(0.1) * (pr + perc);            /* :138 */
            }
            if (tape) { tape->pn[tp] = pn; tape->en[tp] = en; tape->prr[tp] = prr; tape->hft0[tp] = hft[c]; }
            gr_transfer(R(5.0), prcp, prr, cft[c], &hft[c], &qr); /* :140, :687 */
            if (grd) {
                qd = R(0.0);
                qt = qr;                               /* :689 */
            } else {
                qd = fmax(R(0.0), prd + l);            /* :142 */
                qt = (qr + qd);                        /* :144 */
            }
            upstream_discharge(P, row, col, q, &qup);  /* :150 */
            if (tape) {
                tape->qup[tp] = qup; tape->hlr0[tp] = hlr[c];
                tape->flags[tp] = (unsigned char)((nogap ? 1 : 0) | ((!grd && R(0.0) < prd + l) ? 2 : 0));
            }
            linear_routing(P->dt, qup, lr[c], &hlr[c], &qrout); /* :153 */
            q[c] = (qt + qrout * (oreal)(P->flwacc[c] - 1)) * P->dx * P->dx * R(0.001) / P->dt; /* :155 */
            if (netp) { /* :164-176 */
                if (P->sparse_storage) netp[(size_t)P->rowcol_to_ind_sparse[c] - 1 + (size_t)P->nac * t] = qt;
                else netp[c + ncell * t] = qt;
            }
            if (qdom) { /* :182-194 */
                if (P->sparse_storage) qdom[(size_t)P->rowcol_to_ind_sparse[c] - 1 + (size_t)P->nac * t] = q[c];
                else qdom[c + ncell * t] = q[c];
            }
        }
        for (int g = 0; g < P->ng; g++)                /* :206-210 */
            qsim[g + (size_t)P->ng * t] = q[IDX(P, P->gauge_pos[g], P->gauge_pos[g + P->ng])];
    }
    free(q);
}

static void gr_a_forward(const Prob *P, const oreal *par, oreal *st, oreal *qsim, oreal *qdom, oreal *netp, Tape *tape) {
    gr_ad_forward(P, par, st, qsim, qdom, netp, tape, 0);
}

/* ============================================================================================
 * The other structures -- forward/md_forward_structure.f90:216-931, operator/md_gr_operator.f90:20-34,
 * operator/md_vic_operator.f90.  Forward only (the reference's tests and golden file hold no vector of
 * these structures: parity of this part is UNPINNED, see the header).
 * ========================================================================================== */

/* md_gr_operator.f90:20-34 */
static void gr_interception(oreal prcp, oreal pet, oreal ci, oreal *hi, oreal *pn, oreal *ei) {
    *ei = fmin(pet, prcp + (*hi) * ci);
    *pn = fmax(R(0.0), prcp - ci * (R(1.0) - (*hi)) - (*ei));
    *hi = (*hi) + (prcp - (*ei) - (*pn)) / ci;
}

/* md_vic_operator.f90:21-77 */
static void vic_infiltration(oreal prcp, oreal cusl1, oreal cusl2, oreal b, oreal *husl1, oreal *husl2, oreal *runoff) {
    oreal bp1 = b + R(1.0), ifl;
    if (prcp <= R(0.0)) {
        ifl = R(0.0);
    } else {
        oreal cusl = cusl1 + cusl2;
        oreal wusl = (*husl1) * cusl1 + (*husl2) * cusl2;
        wusl = fmax(R(1.e-6), wusl);
        wusl = fmin(cusl - R(1e-6), wusl);
        oreal iflm = cusl * bp1;
        oreal iflc = iflm * (R(1.0) - POW(R(1.0) - (wusl / cusl), R(1.0) / bp1));
        if (iflc + prcp >= iflm) ifl = cusl - wusl;
        else ifl = (cusl - wusl) - cusl * POW(R(1.0) - ((iflc + prcp) / iflm), bp1);
        ifl = fmin(prcp, ifl);
    }
    oreal ifl_usl1 = fmin((R(1.0) - (*husl1)) * cusl1, ifl);
    ifl = ifl - ifl_usl1;
    oreal ifl_usl2 = fmin((R(1.0) - (*husl2)) * cusl2, ifl);
    ifl = ifl - ifl_usl2;
    *husl1 = (*husl1) + ifl_usl1 / cusl1;
    *husl2 = (*husl2) + ifl_usl2 / cusl2;
    *runoff = prcp - (ifl_usl1 + ifl_usl2);
}

/* md_vic_operator.f90:165-183, called with residual = 0, porosity = 1, lambda = 1 (:88, :92) */
static oreal brooks_and_corey_flow(oreal ks, oreal residual, oreal porosity, oreal lambda, oreal c_upper, oreal c_lower,
                                   oreal h_upper, oreal h_lower) {
    oreal flow = ks * POW((h_upper - residual) / (porosity - residual), lambda);
    oreal w_upper = h_upper * c_upper * porosity;
    oreal w_lower = h_lower * c_lower * porosity;
    oreal max_flow = fmin(w_upper, c_lower - w_lower);
    return fmin(max_flow, flow);
}

/* md_vic_operator.f90:185-200 */
static oreal linear_evapotranspiration(oreal e, oreal c, oreal h) { return fmin(c * h, e * h); }

/* md_vic_operator.f90:79-114 */
static void vic_vertical_transfer(oreal pet, oreal cusl1, oreal cusl2, oreal clsl, oreal ks, oreal *husl1, oreal *husl2,
                                  oreal *hlsl) {
    oreal fbc = brooks_and_corey_flow(ks, R(0.0), R(1.0), R(1.0), cusl1, cusl2, *husl1, *husl2);
    *husl1 = (*husl1) - fbc / cusl1;
    *husl2 = (*husl2) + fbc / cusl2;
    fbc = brooks_and_corey_flow(ks, R(0.0), R(1.0), R(1.0), cusl2, clsl, *husl2, *hlsl);
    *husl2 = (*husl2) - fbc / cusl2;
    *hlsl = (*hlsl) + fbc / clsl;
    oreal fe = linear_evapotranspiration(pet, cusl1, *husl1);
    *husl1 = (*husl1) - fe / cusl1;
    oreal pet_remain = fmax(R(0.0), pet - fe);
    fe = linear_evapotranspiration(pet_remain, cusl2, *husl2);
    *husl2 = (*husl2) - fe / cusl2;
    pet_remain = fmax(R(0.0), pet_remain - fe);
    fe = linear_evapotranspiration(pet_remain, clsl, *hlsl);
    *hlsl = (*hlsl) - fe / clsl;
}

/* md_vic_operator.f90:116-135 */
static void vic_interflow(oreal n, oreal cusl2, oreal *husl2, oreal *qi) {
    oreal nm1 = n - R(1.0), d1pnm1 = R(1.0) / nm1, h0 = *husl2;
    *husl2 = POW(POW(h0 * cusl2, -nm1) + POW(cusl2, -nm1), -d1pnm1) / cusl2;
    *qi = (h0 - (*husl2)) * cusl2;
}

/* md_vic_operator.f90:137-163 */
static void vic_baseflow(oreal clsl, oreal ds, oreal dsm, oreal ws, oreal *hlsl, oreal *qb) {
    if ((*hlsl) <= ws) *qb = (ds * dsm) / ws * (*hlsl);
    else *qb = dsm * (R(1.0) - ds / ws) * ((*hlsl) - ws) / (R(1.0) - ws);
    *qb = fmin(clsl * (*hlsl), *qb);
    *hlsl = (*hlsl) - (*qb) / clsl;
}

/* gr_b_forward :216-398, gr_c_forward :400-587, gr_d_forward :589-760, vic_a_forward :762-931: the loop skeleton of
 * gr_a_forward with another runoff-production part (the lines cited per branch); routing and stores are the same */
static void structure_forward(const Prob *P, const oreal *par, oreal *st, oreal *qsim, oreal *qdom, oreal *netp) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    const oreal *ci = par + OP_CI * ncell, *cp = par + OP_CP * ncell, *cft = par + OP_CFT * ncell, *cst = par + OP_CST * ncell,
                *exc = par + OP_EXC * ncell, *lr = par + OP_LR * ncell, *b = par + OP_B * ncell, *cusl1 = par + OP_CUSL1 * ncell,
                *cusl2 = par + OP_CUSL2 * ncell, *clsl = par + OP_CLSL * ncell, *ks = par + OP_KS * ncell, *ds = par + OP_DS * ncell,
                *dsm = par + OP_DSM * ncell, *ws = par + OP_WS * ncell;
    oreal *hi = st + OS_HI * ncell, *hp = st + OS_HP * ncell, *hft = st + OS_HFT * ncell, *hst = st + OS_HST * ncell,
          *husl1 = st + OS_HUSL1 * ncell, *husl2 = st + OS_HUSL2 * ncell, *hlsl = st + OS_HLSL * ncell, *hlr = st + OS_HLR * ncell;
    oreal *q = (oreal *)calloc(ncell, sizeof(oreal));
    const int S = P->structure;

    for (int t = 0; t < P->ntime_step; t++) {
        for (size_t i = 0; i < ncell; i++) {
            int row = P->path[2 * i], col = P->path[2 * i + 1];
            if (!(row > 0 && col > 0)) continue;
            size_t c = IDX(P, row, col);
            if (!(P->active_cell[c] == 1 && (!P->local_active_cell || P->local_active_cell[c] == 1))) continue;
            oreal prcp, pet, qt, qup = 0, qrout = 0;
            read_forcing(P, row, col, t, &prcp, &pet);
            const int nogap = (prcp >= 0 && pet >= 0);
            if (S == OST_VIC_A) {
                oreal runoff = 0, qi = 0, qb = 0;                                    /* :797-804 */
                if (nogap) {
                    vic_infiltration(prcp, cusl1[c], cusl2[c], b[c], &husl1[c], &husl2[c], &runoff);          /* :843 */
                    vic_vertical_transfer(pet, cusl1[c], cusl2[c], clsl[c], ks[c], &husl1[c], &husl2[c], &hlsl[c]); /* :851 */
                }
                vic_interflow(R(5.0), cusl2[c], &husl2[c], &qi);                    /* :861 */
                vic_baseflow(clsl[c], ds[c], dsm[c], ws[c], &hlsl[c], &qb);         /* :863 */
                qt = (runoff + qi + qb);                                            /* :866 */
            } else {
                oreal ei, pn = 0, en = 0, pr = 0, perc = 0, l = 0, prr, prl, prd, qr = 0, ql = 0, qd;
                if (nogap) {
                    if (S == OST_GR_D) {
                        ei = fmin(pet, prcp);                                       /* :666 */
                        pn = fmax(R(0.0), prcp - ei);                               /* :668 */
                    } else {
                        gr_interception(prcp, pet, ci[c], &hi[c], &pn, &ei);        /* :298, :482 */
                    }
                    en = pet - ei;
                    gr_production(pn, en, cp[c], R(1000.0), &hp[c], &pr, &perc);    /* :306, :490, :676 */
                    if (S != OST_GR_D) gr_exchange(exc[c], hft[c], &l);             /* :313, :497 */
                }
                if (S == OST_GR_B) {
                    prr = R(0.9) * (pr + perc) + l;                                 /* :321 */
                    prd = R(0.1) * (pr + perc);                                     /* :322 */
                    gr_transfer(R(5.0), prcp, prr, cft[c], &hft[c], &qr);           /* :324 */
                    qd = fmax(R(0.0), prd + l);                                     /* :326 */
                    qt = (qr + qd);                                                 /* :328 */
                } else if (S == OST_GR_C) {
                    prr = R(0.9) * R(0.6) * (pr + perc) + l;                        /* :505 */
                    prl = R(0.9) * R(0.4) * (pr + perc);                            /* :506 */
                    prd = R(0.1) * (pr + perc);                                     /* :507 */
                    gr_transfer(R(5.0), prcp, prr, cft[c], &hft[c], &qr);           /* :509 */
                    gr_transfer(R(5.0), prcp, prl, cst[c], &hst[c], &ql);           /* :511 */
                    qd = fmax(R(0.0), prd + l);                                     /* :513 */
                    qt = (qr + ql + qd);                                            /* :515 */
                } else {
                    prr = pr + perc;                                                /* :685 */
                    gr_transfer(R(5.0), prcp, prr, cft[c], &hft[c], &qr);           /* :687 */
                    qt = qr;                                                        /* :689 */
                }
            }
            upstream_discharge(P, row, col, q, &qup);
            linear_routing(P->dt, qup, lr[c], &hlr[c], &qrout);
            q[c] = (qt + qrout * (oreal)(P->flwacc[c] - 1)) * P->dx * P->dx * R(0.001) / P->dt;
            if (netp) {
                if (P->sparse_storage) netp[(size_t)P->rowcol_to_ind_sparse[c] - 1 + (size_t)P->nac * t] = qt;
                else netp[c + ncell * t] = qt;
            }
            if (qdom) {
                if (P->sparse_storage) qdom[(size_t)P->rowcol_to_ind_sparse[c] - 1 + (size_t)P->nac * t] = q[c];
                else qdom[c + ncell * t] = q[c];
            }
        }
        for (int g = 0; g < P->ng; g++)
            qsim[g + (size_t)P->ng * t] = q[IDX(P, P->gauge_pos[g], P->gauge_pos[g + P->ng])];
    }
    free(q);
}

/* ============================================================================================
 * Operator adjoints -- forward/forward_db.f90
 * ========================================================================================== */

/* GR_PRODUCTION_B forward_db.f90:6012-6103 (pn_b, en_b are dead: forcing is not a control) */
static void gr_production_b(oreal pn, oreal en, oreal cp, oreal *cp_b, oreal beta, oreal hp, oreal *hp_b,
                            oreal pr_b, oreal perc_b) {
    oreal inv_cp = R(1.0) / cp;
    oreal ps = cp * (R(1.0) - hp * hp) * TANH(pn * inv_cp) / (R(1.0) + hp * TANH(pn * inv_cp));
    oreal es = hp * cp * (R(2.0) - hp) * TANH(en * inv_cp) / (R(1.0) + (R(1.0) - hp) * TANH(en * inv_cp));
    oreal hp_imd = hp + (ps - es) * inv_cp;
    int branch = (pn > 0) ? 0 : 1;
    oreal pwx1 = R(1.0) + pow4i(hp_imd / beta);
    oreal pwr1 = POW(pwx1, R(-0.25));
    oreal perc = hp_imd * cp * (R(1.0) - pwr1);
    perc_b = perc_b - inv_cp * (*hp_b);
    oreal inv_cp_b = -(perc * (*hp_b));
    *cp_b = *cp_b + hp_imd * (R(1.0) - pwr1) * perc_b;
    oreal pwr1_b = -(hp_imd * cp * perc_b);
    oreal pwx1_b = -(R(0.25) * POW(pwx1, R(-1.25)) * pwr1_b);
    oreal hp_imd_b = (*hp_b) + cp * (R(1.0) - pwr1) * perc_b +
                     4 * (hp_imd * hp_imd * hp_imd) * pwx1_b / pow4i(beta);
    if (branch == 0) {
        hp_imd_b = hp_imd_b - cp * pr_b;
        *hp_b = cp * pr_b;
        *cp_b = *cp_b - (hp_imd - hp) * pr_b;
    } else {
        *hp_b = R(0.0);
    }
    oreal es_b = -(inv_cp * hp_imd_b);
    oreal temp4 = TANH(en * inv_cp);
    oreal temp3 = (-hp + R(1.0)) * temp4 + R(1.0);
    oreal temp1 = TANH(en * inv_cp);
    oreal temp0 = hp * cp * (-hp + R(2.0));
    oreal temp_b3 = es_b / temp3;
    oreal temp_b = (R(2.0) - hp) * temp1 * temp_b3;
    oreal temp_b0 = -(temp0 * temp1 * temp_b3 / temp3);
    *hp_b = *hp_b + hp_imd_b + cp * temp_b - hp * cp * temp1 * temp_b3 - temp4 * temp_b0;
    oreal ps_b = inv_cp * hp_imd_b;
    oreal temp_b4 = (R(1.0) - TANH(en * inv_cp) * TANH(en * inv_cp)) * temp0 * temp_b3;
    oreal temp_b5 = (R(1.0) - TANH(en * inv_cp) * TANH(en * inv_cp)) * (R(1.0) - hp) * temp_b0;
    *cp_b = *cp_b + hp * temp_b;
    oreal temp = TANH(pn * inv_cp);
    temp0 = hp * temp + R(1.0);
    temp1 = TANH(pn * inv_cp);
    oreal temp2 = cp * (-(hp * hp) + R(1.0));
    temp_b = ps_b / temp0;
    temp_b0 = (R(1.0) - TANH(pn * inv_cp) * TANH(pn * inv_cp)) * temp2 * temp_b;
    oreal temp_b1 = -(temp2 * temp1 * temp_b / temp0);
    *hp_b = *hp_b + temp * temp_b1 - 2 * hp * cp * temp1 * temp_b;
    oreal temp_b2 = (R(1.0) - TANH(pn * inv_cp) * TANH(pn * inv_cp)) * hp * temp_b1;
    inv_cp_b = inv_cp_b + (ps - es) * hp_imd_b + en * temp_b5 + en * temp_b4 + pn * temp_b2 + pn * temp_b0;
    *cp_b = *cp_b + (R(1.0) - hp * hp) * temp1 * temp_b - inv_cp_b / (cp * cp);
}

/* GR_EXCHANGE_B forward_db.f90:6147-6157 */
static void gr_exchange_b(oreal exc, oreal *exc_b, oreal hft, oreal *hft_b, oreal l_b) {
    *exc_b = *exc_b + POW(hft, R(3.5)) * l_b;
    *hft_b = *hft_b + R(3.5) * POW(hft, R(2.5)) * exc * l_b;
}

/* derivative guard Tapenade emits for x**y with real y (forward_db.f90:6344-6349) */
static inline oreal dpow_b(oreal x, oreal y, oreal r_b) {
    if (x <= R(0.0) && (y == R(0.0) || y != (oreal)(int)y)) return R(0.0);
    return y * POW(x, y - 1) * r_b;
}

/* GR_TRANSFER_B forward_db.f90:6275-6412 ; ht is the PRE-update value, *ht_b in/out */
static void gr_transfer_b(oreal n, oreal prcp, oreal pr, oreal *pr_b, oreal ct, oreal *ct_b, oreal ht,
                          oreal *ht_b, oreal q_b) {
    oreal nm1 = n - R(1.0), d1pnm1 = R(1.0) / nm1;
    oreal pr_imd, ht_imd;
    oreal g_pwx1 = 0, g_pwx3 = 0; /* values of the gap branch, restored by the POPREAL4s (:6377-6382) */
    int br_gap, br_max;
    if (prcp < R(0.0)) {
        g_pwx1 = ht * ct;
        oreal pwr1 = POW(g_pwx1, -nm1);
        oreal pwr2 = POW(ct, -nm1);
        g_pwx3 = pwr1 - pwr2;
        oreal pwr3 = POW(g_pwx3, -d1pnm1);
        pr_imd = pwr3 - ht * ct;
        br_gap = 1;
    } else {
        pr_imd = pr;
        br_gap = 0;
    }
    if (R(1.e-6) < ht + pr_imd / ct) { ht_imd = ht + pr_imd / ct; br_max = 0; }
    else { ht_imd = R(1.e-6); br_max = 1; }
    oreal pwx1 = ht_imd * ct, pwy1 = -nm1;
    oreal pwr1 = POW(pwx1, pwy1);
    oreal pwy2 = -nm1;
    oreal pwr2 = POW(ct, pwy2);
    oreal pwx3 = pwr1 + pwr2, pwy3 = -d1pnm1;
    oreal pwr3 = POW(pwx3, pwy3);
    oreal ht_new = pwr3 / ct;
    (void)pwr2;
    oreal htb = *ht_b - ct * q_b;
    oreal pwr3_b = htb / ct;
    oreal pwx3_b = dpow_b(pwx3, pwy3, pwr3_b);
    oreal pwr1_b = pwx3_b, pwr2_b = pwx3_b;
    oreal pwx1_b = dpow_b(pwx1, pwy1, pwr1_b);
    oreal ht_imd_b = ct * q_b + ct * pwx1_b;
    if (ct <= R(0.0) && (pwy2 == R(0.0) || pwy2 != (oreal)(int)pwy2))
        *ct_b = *ct_b + (ht_imd - ht_new) * q_b + ht_imd * pwx1_b - pwr3 * htb / (ct * ct);
    else
        *ct_b = *ct_b + (ht_imd - ht_new) * q_b + pwy2 * POW(ct, pwy2 - 1) * pwr2_b - pwr3 * htb / (ct * ct) +
                ht_imd * pwx1_b;
    oreal pr_imd_b;
    if (br_max == 0) {
        htb = ht_imd_b;
        pr_imd_b = ht_imd_b / ct;
        *ct_b = *ct_b - pr_imd * ht_imd_b / (ct * ct);
    } else {
        htb = R(0.0);
        pr_imd_b = R(0.0);
    }
    if (br_gap == 0) {
        *pr_b = pr_imd_b;
    } else {
        pwr3_b = pr_imd_b;
        pwx3_b = dpow_b(g_pwx3, pwy3, pwr3_b);
        pwr1_b = pwx3_b;
        pwr2_b = -pwx3_b;
        pwx1_b = dpow_b(g_pwx1, pwy1, pwr1_b);
        htb = htb + ct * pwx1_b - ct * pr_imd_b;
        if (ct <= R(0.0) && (pwy2 == R(0.0) || pwy2 != (oreal)(int)pwy2))
            *ct_b = *ct_b + ht * pwx1_b - ht * pr_imd_b;
        else
            *ct_b = *ct_b + pwy2 * POW(ct, pwy2 - 1) * pwr2_b - ht * pr_imd_b + ht * pwx1_b;
        *pr_b = R(0.0);
    }
    *ht_b = htb;
}

/* LINEAR_ROUTING_B forward_db.f90:6628-6652 ; hr is the PRE-update value */
static void linear_routing_b(oreal dt, oreal qup, oreal *qup_b, oreal lr, oreal *lr_b, oreal hr, oreal *hr_b,
                             oreal qrout_b) {
    oreal hr_imd = hr + qup;
    oreal arg1 = -(dt / (lr * R(60.0)));
    *hr_b = *hr_b - qrout_b;
    oreal hr_imd_b = qrout_b + EXP(arg1) * (*hr_b);
    oreal arg1_b = EXP(arg1) * hr_imd * (*hr_b);
    *lr_b = *lr_b + dt * arg1_b / ((lr * lr) * R(60.0));
    *hr_b = hr_imd_b;
    *qup_b = hr_imd_b;
}

/* UPSTREAM_DISCHARGE_B forward_db.f90:6520-6564 */
static void upstream_discharge_b(const Prob *P, int row, int col, oreal *q_b, oreal qup_b) {
    int fa = P->flwacc[IDX(P, row, col)];
    if (fa > 1) {
        qup_b = P->dt * qup_b / (R(0.001) * (P->dx * P->dx) * (oreal)(fa - 1));
        for (int i = 7; i >= 0; i--) {
            int col_imd = col + DCOL[i], row_imd = row + DROW[i];
            if (col_imd > 0 && col_imd <= P->ncol && row_imd > 0 && row_imd <= P->nrow)
                if (P->flwdir[IDX(P, row_imd, col_imd)] == i + 1) q_b[IDX(P, row_imd, col_imd)] += qup_b;
        }
    }
}

/* GR_A_FORWARD_B forward_db.f90:7954-8175: taped forward sweep then reverse sweep.
 * qsim_b is consumed (zeroed) like output_b%qsim (:8107). */
/* GR_A_FORWARD_B forward_db.f90:7954-8175; with grd != 0 GR_D_FORWARD_B :9604-9797 (qr_b = qt_b :9763, pr_b = perc_b = prr_b
 * :9770-9771, no exchange, parameters_b%exc untouched) */
static void gr_a_forward_b(const Prob *P, const oreal *par, oreal *par_b, oreal *st, oreal *st_b, oreal *qsim,
                           oreal *qsim_b) {
    const int grd = P->structure == OST_GR_D;
    const size_t ncell = (size_t)P->nrow * P->ncol;
    Tape tape;
    tape_alloc(&tape, P);
    oreal *st0 = (oreal *)malloc(O_GNS * ncell * sizeof(oreal));
    memcpy(st0, st, O_GNS * ncell * sizeof(oreal));
    gr_ad_forward(P, par, st, qsim, NULL, NULL, &tape, grd);
    memcpy(st, st0, O_GNS * ncell * sizeof(oreal)); /* the POPREAL4s leave the initial states */
    free(st0);

    const oreal *cp = par + OP_CP * ncell, *cft = par + OP_CFT * ncell, *exc = par + OP_EXC * ncell,
                *lr = par + OP_LR * ncell;
    oreal *cp_b = par_b + OP_CP * ncell, *cft_b = par_b + OP_CFT * ncell, *exc_b = par_b + OP_EXC * ncell,
          *lr_b = par_b + OP_LR * ncell;
    oreal *hp_b = st_b + OS_HP * ncell, *hft_b = st_b + OS_HFT * ncell, *hlr_b = st_b + OS_HLR * ncell;
    oreal *q_b = (oreal *)calloc(ncell, sizeof(oreal));

    for (int t = P->ntime_step - 1; t >= 0; t--) {          /* :8103 */
        for (int g = P->ng - 1; g >= 0; g--) {              /* :8104-8108 */
            size_t c = IDX(P, P->gauge_pos[g], P->gauge_pos[g + P->ng]);
            q_b[c] = q_b[c] + qsim_b[g + (size_t)P->ng * t];
            qsim_b[g + (size_t)P->ng * t] = R(0.0);
        }
        for (size_t ii = ncell; ii-- > 0;) {                /* :8109 */
            if (tape.slot[ii] < 0) continue;
            int row = P->path[2 * ii], col = P->path[2 * ii + 1];
            size_t c = IDX(P, row, col);
            size_t tp = (size_t)tape.slot[ii] + (size_t)tape.ncolumns * t;
            oreal prcp, pet;
            read_forcing(P, row, col, t, &prcp, &pet);
            oreal temp_b = P->dx * P->dx * R(0.001) * q_b[c] / P->dt; /* :8114 */
            q_b[c] = R(0.0);
            oreal qt_b = temp_b;
            oreal qrout_b = (oreal)(P->flwacc[c] - 1) * temp_b;
            oreal qup_b = R(0.0);
            linear_routing_b(P->dt, tape.qup[tp], &qup_b, lr[c], &lr_b[c], tape.hlr0[tp], &hlr_b[c], qrout_b);
            upstream_discharge_b(P, row, col, q_b, qup_b);
            oreal qr_b = qt_b, qd_b = qt_b, prd_b, l_b;
            if (tape.flags[tp] & 2) { prd_b = qd_b; l_b = qd_b; }  /* :8128-8137 */
            else { l_b = R(0.0); prd_b = R(0.0); }
            oreal prr_b = R(0.0);
            gr_transfer_b(R(5.0), prcp, tape.prr[tp], &prr_b, cft[c], &cft_b[c], tape.hft0[tp], &hft_b[c], qr_b);
            oreal pr_b = grd ? prr_b : R(0.1) * prd_b + R(0.9) * prr_b;    /* :8143, :9770 */
            oreal perc_b = grd ? prr_b : R(0.1) * prd_b + R(0.9) * prr_b;  /* :8144, :9771 */
            l_b = l_b + prr_b;
            if (tape.flags[tp] & 1) {                        /* :8148-8165, :9773-9787 */
                if (!grd) gr_exchange_b(exc[c], &exc_b[c], tape.hft0[tp], &hft_b[c], l_b);
                gr_production_b(tape.pn[tp], tape.en[tp], cp[c], &cp_b[c], R(1000.0), tape.hp0[tp], &hp_b[c], pr_b,
                                perc_b);
            }
        }
    }
    free(q_b);
    tape_free(&tape);
}

/* ============================================================================================
 * Cost -- optimize/mwd_cost.f90 and its adjoint in forward_db.f90
 * ========================================================================================== */

/* mwd_cost.f90:350-401 */
static oreal nse(const oreal *x, const oreal *y, int size) {
    int n = 0;
    oreal sum_x = 0, sum_xx = 0, sum_yy = 0, sum_xy = 0;
    for (int i = 0; i < size; i++)
        if (x[i] >= R(0.0)) {
            n++;
            sum_x = sum_x + x[i];
            sum_xx = sum_xx + (x[i] * x[i]);
            sum_yy = sum_yy + (y[i] * y[i]);
            sum_xy = sum_xy + (x[i] * y[i]);
        }
    oreal mean_x = sum_x / n;
    oreal num = sum_xx - 2 * sum_xy + sum_yy;
    oreal den = sum_xx - n * mean_x * mean_x;
    return num / den;
}

/* NSE_B forward_db.f90:3505-3545 */
static void nse_b(const oreal *x, const oreal *y, oreal *y_b, int size, oreal res_b) {
    int n = 0;
    oreal sum_x = 0, sum_xx = 0;
    for (int i = 0; i < size; i++)
        if (x[i] >= R(0.0)) { n++; sum_x = sum_x + x[i]; sum_xx = sum_xx + x[i] * x[i]; }
    oreal mean_x = sum_x / n;
    oreal den = sum_xx - n * mean_x * mean_x;
    oreal num_b = res_b / den;
    oreal sum_yy_b = num_b, sum_xy_b = -(2 * num_b);
    for (int i = size - 1; i >= 0; i--)
        if (x[i] >= R(0.0)) y_b[i] = y_b[i] + x[i] * sum_xy_b + 2 * y[i] * sum_yy_b;
}

typedef struct { oreal mean_x, mean_y, var_x, var_y, cov; int n; } KgeMoments;

/* mwd_cost.f90:403-459 */
static KgeMoments kge_moments(const oreal *x, const oreal *y, int size) {
    KgeMoments m;
    int n = 0;
    oreal sum_x = 0, sum_y = 0, sum_xx = 0, sum_yy = 0, sum_xy = 0;
    for (int i = 0; i < size; i++)
        if (x[i] >= R(0.0)) {
            n++;
            sum_x = sum_x + x[i];
            sum_y = sum_y + y[i];
            sum_xx = sum_xx + (x[i] * x[i]);
            sum_yy = sum_yy + (y[i] * y[i]);
            sum_xy = sum_xy + (x[i] * y[i]);
        }
    m.n = n;
    m.mean_x = sum_x / n;
    m.mean_y = sum_y / n;
    m.var_x = (sum_xx / n) - (m.mean_x * m.mean_x);
    m.var_y = (sum_yy / n) - (m.mean_y * m.mean_y);
    m.cov = (sum_xy / n) - (m.mean_x * m.mean_y);
    return m;
}
static void kge_components(const oreal *x, const oreal *y, int size, oreal *r, oreal *a, oreal *b) {
    KgeMoments m = kge_moments(x, y, size);
    *r = (m.cov / SQRT(m.var_x)) / SQRT(m.var_y);
    *a = SQRT(m.var_y) / SQRT(m.var_x);
    *b = m.mean_y / m.mean_x;
}
/* mwd_cost.f90:461-490 */
static oreal kge(const oreal *x, const oreal *y, int size) {
    oreal r, a, b;
    kge_components(x, y, size, &r, &a, &b);
    return SQRT((r - 1) * (r - 1) + (b - 1) * (b - 1) + (a - 1) * (a - 1));
}
/* KGE_B forward_db.f90:3805-3829 + KGE_COMPONENTS_B :3657-3729 */
static void kge_b(const oreal *x, const oreal *y, oreal *y_b, int size, oreal res_b) {
    oreal r, a, b;
    kge_components(x, y, size, &r, &a, &b);
    oreal arg1 = (r - 1) * (r - 1) + (b - 1) * (b - 1) + (a - 1) * (a - 1);
    oreal arg1_b = (arg1 == R(0.0)) ? R(0.0) : res_b / (R(2.0) * SQRT(arg1));
    oreal r_b = 2 * (r - 1) * arg1_b, b_b = 2 * (b - 1) * arg1_b, a_b = 2 * (a - 1) * arg1_b;
    KgeMoments m = kge_moments(x, y, size);
    oreal result1 = SQRT(m.var_x), result2 = SQRT(m.var_y);
    oreal result1_b = a_b / SQRT(m.var_x);
    oreal var_y_b = (m.var_y == R(0.0)) ? R(0.0) : result1_b / (R(2.0) * SQRT(m.var_y));
    oreal temp_b = r_b / (result1 * result2);
    oreal cov_b = temp_b;
    oreal result2_b = -(m.cov * temp_b / result2);
    if (!(m.var_y == R(0.0))) var_y_b = var_y_b + result2_b / (R(2.0) * SQRT(m.var_y));
    oreal mean_y_b = b_b / m.mean_x - m.mean_x * cov_b - 2 * m.mean_y * var_y_b;
    oreal sum_xy_b = cov_b / m.n, sum_yy_b = var_y_b / m.n, sum_y_b = mean_y_b / m.n;
    for (int i = size - 1; i >= 0; i--)
        if (x[i] >= R(0.0)) y_b[i] = y_b[i] + x[i] * sum_xy_b + 2 * y[i] * sum_yy_b + sum_y_b;
}
/* mwd_cost.f90:492-522 */
static oreal se(const oreal *x, const oreal *y, int size) {
    oreal res = 0;
    for (int i = 0; i < size; i++)
        if (x[i] >= R(0.0)) res = res + (x[i] - y[i]) * (x[i] - y[i]);
    return res;
}
/* mwd_cost.f90:524-560 */
static oreal rmse(const oreal *x, const oreal *y, int size) {
    int n = 0;
    for (int i = 0; i < size; i++)
        if (x[i] >= R(0.0)) n++;
    return SQRT(se(x, y, size) / n);
}
/* mwd_cost.f90:562-592 */
static oreal logarithmic(const oreal *x, const oreal *y, int size) {
    oreal res = 0;
    for (int i = 0; i < size; i++)
        if (x[i] > R(0.0) && y[i] > R(0.0)) res = res + x[i] * LOG(y[i] / x[i]) * LOG(y[i] / x[i]);
    return res;
}
/* SE_B / RMSE_B / LOGARITHMIC_B: plain reverse of the three functions above (forward_db.f90:3880-4200) */
static void se_b(const oreal *x, const oreal *y, oreal *y_b, int size, oreal res_b) {
    for (int i = size - 1; i >= 0; i--)
        if (x[i] >= R(0.0)) y_b[i] = y_b[i] - 2 * (x[i] - y[i]) * res_b;
}
static void rmse_b(const oreal *x, const oreal *y, oreal *y_b, int size, oreal res_b) {
    int n = 0;
    for (int i = 0; i < size; i++)
        if (x[i] >= R(0.0)) n++;
    oreal s = se(x, y, size) / n;
    oreal s_b = (s == R(0.0)) ? R(0.0) : res_b / (R(2.0) * SQRT(s));
    se_b(x, y, y_b, size, s_b / n);
}
static void logarithmic_b(const oreal *x, const oreal *y, oreal *y_b, int size, oreal res_b) {
    for (int i = size - 1; i >= 0; i--)
        if (x[i] > R(0.0) && y[i] > R(0.0)) y_b[i] = y_b[i] + 2 * x[i] * LOG(y[i] / x[i]) * res_b / y[i];
}

static oreal compute_jobs(const Prob *P, const oreal *qsim, oreal *qsim_b, oreal jobs_b);
/* compute_jobs on a given qsim(ng,T) (lets tests check the cost kernel independently of the simulation) */
oreal OSYM(oracle_compute_jobs)(const Prob *P, const oreal *qsim, oreal *qsim_b) {
    return compute_jobs(P, qsim, qsim_b, R(1.0));
}
oreal OSYM(oracle_nse)(const oreal *x, const oreal *y, int n) { return nse(x, y, n); }
oreal OSYM(oracle_kge)(const oreal *x, const oreal *y, int n) { return kge(x, y, n); }

/* quantile (mwd_cost.f90:675-720) with p = 0.5 over <= ng gauge costs; heap_sort (:594-673) is replaced by
 * qsort (same sorted result).  lo/hi/frac_out describe the two order statistics for QUANTILE_B
 * (forward_db.f90:4327): res = q1 + (q2 - q1)*frac. */
static int cmp_oreal(const void *a, const void *b) {
    oreal x = *(const oreal *)a, y = *(const oreal *)b;
    return (x > y) - (x < y);
}
static oreal quantile_half(const oreal *dat, int n, int *lo, int *hi, oreal *frac_out) {
    oreal res = dat[0];
    *lo = *hi = 0;
    *frac_out = 0;
    if (n > 1) {
        oreal *s = (oreal *)malloc(n * sizeof(oreal));
        memcpy(s, dat, n * sizeof(oreal));
        qsort(s, n, sizeof(oreal), cmp_oreal);
        oreal frac = (n - 1) * R(0.5) + 1; /* 1-based fractional rank */
        int i1, i2;
        oreal w;
        if (frac <= 1) { i1 = i2 = 0; w = 0; res = s[0]; }
        else if (frac >= n) { i1 = i2 = n - 1; w = 0; res = s[n - 1]; }
        else {
            i1 = (int)frac - 1;
            i2 = i1 + 1;
            w = frac - (int)frac;
            res = s[i1] + (s[i2] - s[i1]) * w;
        }
        *frac_out = w;
        *lo = *hi = -1;
        for (int i = 0; i < n; i++) if (*lo < 0 && dat[i] == s[i1]) *lo = i;
        for (int i = n - 1; i >= 0; i--) if (*hi < 0 && dat[i] == s[i2]) *hi = i;
        free(s);
    }
    return res;
}


/* quantile (mwd_cost.f90:675-720) of dat(1:n): heap_sort (:594-673) is replaced by qsort (same sorted result) */
static oreal quantile_p(const oreal *dat, int n, oreal p) {
    oreal res;
    if (n <= 0) return 0;
    if (n == 1) return dat[0];
    oreal *s = (oreal *)malloc(n * sizeof(oreal));
    memcpy(s, dat, n * sizeof(oreal));
    qsort(s, n, sizeof(oreal), cmp_oreal);
    oreal frac = (n - 1) * p + 1;
    if (frac <= 1) res = s[0];
    else if (frac >= n) res = s[n - 1];
    else { oreal q1 = s[(int)frac - 1], q2 = s[(int)frac]; res = q1 + (q2 - q1) * (frac - (int)frac); }
    free(s);
    return res;
}
/* flow_percentile (mwd_cost.f90:722-768): quantiles over the steps where both series are non-negative */
static void flow_percentile(const oreal *qo, const oreal *qs, int n, oreal p, oreal *num, oreal *den) {
    oreal *a = (oreal *)malloc((size_t)(n > 0 ? n : 1) * 2 * sizeof(oreal)), *b = a + n;
    int j = 0;
    for (int i = 0; i < n; i++) if (qo[i] >= 0 && qs[i] >= 0) { a[j] = qo[i]; b[j] = qs[i]; j++; }
    *num = quantile_p(b, j, p);
    *den = quantile_p(a, j, p);
    free(a);
}
/* signature (mwd_cost.f90:770-970): |s(qs) / s(qo) - 1| for one signature; event-based ones are averaged over the events
 * of mask_event.  num / den keep their previous values when a case does not set them, as in the Fortran function where
 * they are uninitialised locals: they start at 0 here (den = 0 adds nothing). */
static oreal signature(const oreal *po, const oreal *qo, const oreal *qs, const int *mask, int mstride, int n, int stype) {
    oreal res = 0, num = 0, den = 0;
    if (stype == OJ_ERC || stype == OJ_ELT || stype == OJ_EPF) {
        int n_event = 0;
        for (int i = n - 1; i >= 0; i--) if (mask[(size_t)i * mstride] > 0) { n_event = mask[(size_t)i * mstride]; break; }
        for (int ev = 1; ev <= n_event; ev++) {
            int start = -1, cnt = 0;
            for (int j = 0; j < n; j++) if (mask[(size_t)j * mstride] == ev) { if (start < 0) start = j; cnt++; }
            if (start < 0) start = 0;          /* Fortran: start_event keeps its previous value; an event number always occurs */
            oreal sum_qo = 0, sum_qs = 0, sum_po = 0, max_qo = 0, max_qs = 0, max_po = 0;
            int imax_qo = 0, imax_qs = 0, imax_po = 0;
            for (int j = start; j < start + cnt && j < n; j++) {
                if (qo[j] >= 0 && po[j] >= 0) {
                    sum_qo += qo[j]; sum_qs += qs[j]; sum_po += po[j];
                    if (qo[j] > max_qo) { max_qo = qo[j]; imax_qo = j + 1; }
                    if (qs[j] > max_qs) { max_qs = qs[j]; imax_qs = j + 1; }
                    if (po[j] > max_po) { max_po = po[j]; imax_po = j + 1; }
                }
            }
            if (stype == OJ_EPF) { num = max_qs; den = max_qo; }
            else if (stype == OJ_ELT) { num = (oreal)(imax_qs - imax_po); den = (oreal)(imax_qo - imax_po); }
            else if (sum_po > 0) { num = sum_qs / sum_po; den = sum_qo / sum_po; }
            if (den > 0) res = res + FABS(num / den - 1);
        }
        if (n_event > 0) res = res / n_event;
    } else {
        if (stype == OJ_CRC) {
            oreal sum_qo = 0, sum_qs = 0, sum_po = 0;
            for (int i = 0; i < n; i++) if (qo[i] >= 0 && po[i] >= 0) { sum_qo += qo[i]; sum_qs += qs[i]; sum_po += po[i]; }
            if (sum_po > 0) { num = sum_qs / sum_po; den = sum_qo / sum_po; }
        } else {
            const oreal p = stype == OJ_CFP2 ? R(0.02) : stype == OJ_CFP10 ? R(0.1) : stype == OJ_CFP50 ? R(0.5) : R(0.9);
            flow_percentile(qo, qs, n, p, &num, &den);
        }
        if (den > 0) res = FABS(num / den - 1);
    }
    return res;
}

/* compute_jobs (mwd_cost.f90:37-156) and COMPUTE_JOBS_B (forward_db.f90:2553-2715).
 * If qsim_b != NULL the adjoint seed jobs_b is propagated into qsim_b (which is first zeroed, :2660). */
static oreal compute_jobs(const Prob *P, const oreal *qsim, oreal *qsim_b, oreal jobs_b) {
    const int s0 = P->optimize_start_step - 1;
    const int n = P->ntime_step - s0;
    const int ng = P->ng;
    oreal jobs = 0;
    oreal *qo = (oreal *)malloc((size_t)(n > 0 ? n : 1) * sizeof(oreal) * 4);
    oreal *qs = qo + n, *qs_b = qs + n, *po = qs_b + n;
    oreal *arr = (oreal *)calloc(ng > 0 ? ng : 1, sizeof(oreal));
    oreal *gjobs_b = (oreal *)calloc(ng > 0 ? ng : 1, sizeof(oreal));
    int *arr_g = (int *)malloc((ng > 0 ? ng : 1) * sizeof(int));
    int arr_size = 0;
    oreal j_imd = 0;
    for (int g = 0; g < ng; g++) {
        oreal gauge_jobs = 0;
        oreal wg = P->wgauge[g];
        if (wg > 0 || wg < 0) {
            size_t c = IDX(P, P->gauge_pos[g], P->gauge_pos[g + ng]);
            int any = 0;
            for (int i = 0; i < n; i++) {
                qs[i] = qsim[g + (size_t)ng * (s0 + i)] * P->dt / P->area[g] * R(1e3);
                qo[i] = P->qobs[g + (size_t)ng * (s0 + i)] * P->dt / ((oreal)P->flwacc[c] * P->dx * P->dx) * R(1e3);
                if (qo[i] >= 0) any = 1;
            }
            for (int j = 0; j < P->njf; j++) {
                if (any) {
                    switch (P->jobs_fun[j]) {
                        case OJ_NSE: j_imd = nse(qo, qs, n); break;
                        case OJ_KGE: j_imd = kge(qo, qs, n); break;
                        case OJ_KGE2: { oreal imd = kge(qo, qs, n); j_imd = imd * imd; } break;
                        case OJ_SE: j_imd = se(qo, qs, n); break;
                        case OJ_RMSE: j_imd = rmse(qo, qs, n); break;
                        case OJ_LOGARITHMIC: j_imd = logarithmic(qo, qs, n); break;
                        default:
                            if (P->jobs_fun[j] >= OJ_CRC && P->jobs_fun[j] <= OJ_EPF && P->mean_prcp && P->mask_event) {
                                for (int i = 0; i < n; i++) po[i] = P->mean_prcp[g + (size_t)ng * (s0 + i)];
                                j_imd = signature(po, qo, qs, P->mask_event + g + (size_t)ng * s0, ng, n, P->jobs_fun[j]);
                            }
                            break;
                    }
                }
                gauge_jobs = gauge_jobs + P->wjobs_fun[j] * j_imd;
            }
            if (wg > 0) jobs = jobs + wg * gauge_jobs;
            else { arr_g[arr_size] = g; arr[arr_size++] = gauge_jobs; }
        }
    }
    int qlo = 0, qhi = 0;
    oreal qfrac = 0;
    if (arr_size > 0) jobs = quantile_half(arr, arr_size, &qlo, &qhi, &qfrac);

    if (qsim_b) {
        memset(qsim_b, 0, sizeof(oreal) * (size_t)ng * P->ntime_step);
        if (arr_size > 0) { /* QUANTILE_B: weights of the two order statistics */
            if (qhi != qlo) { gjobs_b[arr_g[qlo]] += (1 - qfrac) * jobs_b; gjobs_b[arr_g[qhi]] += qfrac * jobs_b; }
            else gjobs_b[arr_g[qlo]] += jobs_b;
            jobs_b = 0;
        }
        oreal j_imd_b = 0;
        for (int g = ng - 1; g >= 0; g--) {
            oreal wg = P->wgauge[g];
            if (!(wg > 0 || wg < 0)) continue;
            oreal gauge_jobs_b = (wg > 0) ? wg * jobs_b : gjobs_b[g];
            size_t c = IDX(P, P->gauge_pos[g], P->gauge_pos[g + ng]);
            int any = 0;
            for (int i = 0; i < n; i++) {
                qs[i] = qsim[g + (size_t)ng * (s0 + i)] * P->dt / P->area[g] * R(1e3);
                qo[i] = P->qobs[g + (size_t)ng * (s0 + i)] * P->dt / ((oreal)P->flwacc[c] * P->dx * P->dx) * R(1e3);
                if (qo[i] >= 0) any = 1;
                qs_b[i] = 0;
            }
            for (int j = P->njf - 1; j >= 0; j--) {
                j_imd_b = j_imd_b + P->wjobs_fun[j] * gauge_jobs_b;
                if (!any) continue;
                switch (P->jobs_fun[j]) {
                    case OJ_NSE: nse_b(qo, qs, qs_b, n, j_imd_b); j_imd_b = 0; break;
                    case OJ_KGE: kge_b(qo, qs, qs_b, n, j_imd_b); j_imd_b = 0; break;
                    case OJ_KGE2: { oreal imd = kge(qo, qs, n); kge_b(qo, qs, qs_b, n, 2 * imd * j_imd_b); j_imd_b = 0; } break;
                    case OJ_SE: se_b(qo, qs, qs_b, n, j_imd_b); j_imd_b = 0; break;
                    case OJ_RMSE: rmse_b(qo, qs, qs_b, n, j_imd_b); j_imd_b = 0; break;
                    case OJ_LOGARITHMIC: logarithmic_b(qo, qs, qs_b, n, j_imd_b); j_imd_b = 0; break;
                    default: break;
                }
            }
            for (int i = 0; i < n; i++)
                qsim_b[g + (size_t)ng * (s0 + i)] = qsim_b[g + (size_t)ng * (s0 + i)] + P->dt * R(1e3) * qs_b[i] / P->area[g];
        }
    }
    free(qo); free(arr); free(gjobs_b); free(arr_g);
    return jobs;
}

/* reg_prior mwd_cost.f90:1180-1221 / REG_PRIOR_B forward_db.f90:5756-5799 */
static oreal reg_prior(const Prob *P, const int *optim, int nplanes, const oreal *mat, const oreal *bgd,
                       oreal *mat_b, oreal res_b) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    oreal res = 0;
    for (int i = 0; i < nplanes; i++)
        if (optim[i] > 0)
            for (size_t c = 0; c < ncell; c++) {
                oreal d = mat[c + i * ncell] - bgd[c + i * ncell];
                res = res + d * d;
            }
    if (mat_b)
        for (int i = nplanes - 1; i >= 0; i--)
            if (optim[i] > 0)
                for (size_t c = ncell; c-- > 0;)
                    mat_b[c + i * ncell] = mat_b[c + i * ncell] + R(2.0) * (mat[c + i * ncell] - bgd[c + i * ncell]) * res_b;
    return res;
}

/* reg_smoothing mwd_cost.f90:1100-1178 / REG_SMOOTHING_B forward_db.f90:5504-5657 */
static oreal reg_smoothing(const Prob *P, const int *optim, int nplanes, const oreal *matrix, const oreal *bgd,
                           int rel_to_bgd, oreal *matrix_b, oreal res_b) {
    const int nrow = P->nrow, ncol = P->ncol;
    const size_t ncell = (size_t)nrow * ncol;
    oreal *mat = (oreal *)malloc(ncell * nplanes * sizeof(oreal));
    for (size_t k = 0; k < ncell * nplanes; k++) mat[k] = rel_to_bgd ? matrix[k] - bgd[k] : matrix[k];
    oreal res = 0;
#define M(r, c, i) mat[(size_t)((r)-1) + (size_t)((c)-1) * nrow + (size_t)(i) * ncell]
#define MB(r, c, i) matrix_b[(size_t)((r)-1) + (size_t)((c)-1) * nrow + (size_t)(i) * ncell]
#define ACT(r, c) P->active_cell[(size_t)((r)-1) + (size_t)((c)-1) * nrow]
    for (int pass = 0; pass < (matrix_b ? 2 : 1); pass++)
        for (int ii = 0; ii < nplanes; ii++) {
            int i = pass ? nplanes - 1 - ii : ii;
            if (!(optim[i] > 0)) continue;
            for (int cc = 1; cc <= ncol; cc++)
                for (int rr = 1; rr <= nrow; rr++) {
                    int col = pass ? ncol + 1 - cc : cc, row = pass ? nrow + 1 - rr : rr;
                    if (ACT(row, col) != 1) continue;
                    int min_col = col - 1 > 1 ? col - 1 : 1;
                    int max_col = col + 1 < ncol ? col + 1 : ncol;
                    int min_row = row - 1 > 1 ? row - 1 : 1;
                    int max_row = row + 1 < nrow ? row + 1 : nrow;
                    if (ACT(row, min_col) == 0) min_col = col;
                    if (ACT(row, max_col) == 0) max_col = col;
                    if (ACT(min_row, col) == 0) min_row = row;
                    if (ACT(max_row, col) == 0) max_row = row;
                    oreal dr = M(max_row, col, i) - R(2.0) * M(row, col, i) + M(min_row, col, i);
                    oreal dc = M(row, max_col, i) - R(2.0) * M(row, col, i) + M(row, min_col, i);
                    if (!pass) {
                        res = res + (dr * dr + dc * dc);
                    } else {
                        oreal temp_b = R(2.0) * dr * res_b, temp_b0 = R(2.0) * dc * res_b;
                        MB(row, max_col, i) += temp_b0;
                        MB(row, col, i) -= R(2.0) * temp_b0;
                        MB(row, min_col, i) += temp_b0;
                        MB(max_row, col, i) += temp_b;
                        MB(row, col, i) -= R(2.0) * temp_b;
                        MB(min_row, col, i) += temp_b;
                    }
                }
        }
#undef M
#undef MB
#undef ACT
    free(mat);
    return res;
}

/* compute_jreg mwd_cost.f90:159-245 / COMPUTE_JREG_B forward_db.f90:2927-3092.
 * par_b / st_b (if given) are OVERWRITTEN by the regularisation adjoint (:3056-3090). */
static oreal compute_jreg(const Prob *P, const oreal *par, const oreal *par_bgd, const oreal *st, const oreal *st_bgd,
                          oreal *par_b, oreal *st_b, oreal jreg_b) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    oreal parameters_jreg = 0, states_jreg = 0;
    for (int i = 0; i < P->njr; i++) {
        oreal w = P->wjreg_fun[i];
        switch (P->jreg_fun[i]) {
            case OR_PRIOR:
                parameters_jreg = parameters_jreg + w * reg_prior(P, P->optim_parameters, O_GNP, par, par_bgd, NULL, 0);
                states_jreg = states_jreg + w * reg_prior(P, P->optim_states, O_GNS, st, st_bgd, NULL, 0);
                break;
            case OR_SMOOTHING:
                parameters_jreg = parameters_jreg + POW(w, R(2.0)) * reg_smoothing(P, P->optim_parameters, O_GNP, par, par_bgd, 1, NULL, 0);
                states_jreg = states_jreg + POW(w, R(2.0)) * reg_smoothing(P, P->optim_states, O_GNS, st, st_bgd, 1, NULL, 0);
                break;
            case OR_HARD_SMOOTHING:
                parameters_jreg = parameters_jreg + POW(w, R(2.0)) * reg_smoothing(P, P->optim_parameters, O_GNP, par, par_bgd, 0, NULL, 0);
                states_jreg = states_jreg + POW(w, R(2.0)) * reg_smoothing(P, P->optim_states, O_GNS, st, st_bgd, 0, NULL, 0);
                break;
            default: break;
        }
    }
    if (par_b && st_b) {
        memset(par_b, 0, O_GNP * ncell * sizeof(oreal));
        memset(st_b, 0, O_GNS * ncell * sizeof(oreal));
        for (int i = P->njr - 1; i >= 0; i--) {
            oreal w = P->wjreg_fun[i];
            switch (P->jreg_fun[i]) {
                case OR_PRIOR:
                    reg_prior(P, P->optim_states, O_GNS, st, st_bgd, st_b, w * jreg_b);
                    reg_prior(P, P->optim_parameters, O_GNP, par, par_bgd, par_b, w * jreg_b);
                    break;
                case OR_SMOOTHING:
                    reg_smoothing(P, P->optim_states, O_GNS, st, st_bgd, 1, st_b, POW(w, R(2.0)) * jreg_b);
                    reg_smoothing(P, P->optim_parameters, O_GNP, par, par_bgd, 1, par_b, POW(w, R(2.0)) * jreg_b);
                    break;
                case OR_HARD_SMOOTHING:
                    reg_smoothing(P, P->optim_states, O_GNS, st, st_bgd, 0, st_b, POW(w, R(2.0)) * jreg_b);
                    reg_smoothing(P, P->optim_parameters, O_GNP, par, par_bgd, 0, par_b, POW(w, R(2.0)) * jreg_b);
                    break;
                default: break;
            }
        }
    }
    return parameters_jreg + states_jreg;
}

/* (de)normalisation -- routine/mwd_parameters_manipulation.f90:154-206, mwd_states_manipulation.f90:137-189 */
static void normalize(const Prob *P, oreal *a, int nplanes, const oreal *lb, const oreal *ub) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    for (int i = 0; i < nplanes; i++)
        for (size_t c = 0; c < ncell; c++) a[c + i * ncell] = (a[c + i * ncell] - lb[i]) / (ub[i] - lb[i]);
}
static void denormalize(const Prob *P, oreal *a, int nplanes, const oreal *lb, const oreal *ub) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    for (int i = 0; i < nplanes; i++)
        for (size_t c = 0; c < ncell; c++) a[c + i * ncell] = a[c + i * ncell] * (ub[i] - lb[i]) + lb[i];
}

/* compute_cost mwd_cost.f90:247-306 */
static oreal compute_cost(const Prob *P, oreal *par, const oreal *par_bgd, oreal *st, const oreal *st_bgd,
                          const oreal *qsim, oreal *out_cost) {
    oreal jobs = compute_jobs(P, qsim, NULL, 0);
    if (P->denormalize_forward) {
        normalize(P, par, O_GNP, P->lb_parameters, P->ub_parameters);
        normalize(P, st, O_GNS, P->lb_states, P->ub_states);
    }
    oreal jreg = compute_jreg(P, par, par_bgd, st, st_bgd, NULL, NULL, 0);
    if (P->denormalize_forward) {
        denormalize(P, par, O_GNP, P->lb_parameters, P->ub_parameters);
        denormalize(P, st, O_GNS, P->lb_states, P->ub_states);
    }
    oreal cost = jobs + P->wjreg * jreg;
    if (out_cost) { out_cost[0] = cost; out_cost[1] = jobs; out_cost[2] = jreg; }
    return cost;
}

/* ============================================================================================
 * adjust_interception_store -- routine/mw_interception_store.f90:19-160 (UNPINNED like the structures that use ci).
 * ci of every computed cell := the capacity among 0.1, 0.2 .. 4.9 mm (arange_r, m_array_creation.f90:41-54) whose
 * cumulated sub-daily interception evaporation is closest to the cumulated daily one.  The reference loops candidate ->
 * time -> cell; cells are independent, so cell -> candidate -> time below forms the same sums in the same order.
 * ========================================================================================== */
int OSYM(oracle_adjust_interception_store)(const Prob *P, int nday, const int *day_index, oreal *ci) {
    const oreal stt = R(0.1), stp = R(5.0), step = R(0.1);
    const int ncand = (int)ceil((double)((stp - stt) / step));           /* :33 */
    if (ncand <= 0 || ncand > 64 || nday <= 0) return 1;
    oreal cmax[64];
    for (int i = 0; i < ncand; i++) cmax[i] = stt + (oreal)i * step;      /* m_array_creation.f90:50 */
    oreal *daily_p = (oreal *)malloc((size_t)nday * sizeof(oreal)), *daily_e = (oreal *)malloc((size_t)nday * sizeof(oreal));
    for (int col = 1; col <= P->ncol; col++) {
        for (int row = 1; row <= P->nrow; row++) {
            size_t c = IDX(P, row, col);
            if (!(P->active_cell[c] == 1 && (!P->local_active_cell || P->local_active_cell[c] == 1))) continue; /* :100, :138 */
            for (int d = 0; d < nday; d++) daily_p[d] = daily_e[d] = 0;
            int n = 0;
            for (int t = 0; t < P->ntime_step; t++) {                     /* :44-75 */
                if (t > 0 && day_index[t] != day_index[t - 1]) n++;
                if (n >= nday) { free(daily_p); free(daily_e); return 1; }
                oreal prcp, pet;
                read_forcing(P, row, col, t, &prcp, &pet);
                daily_p[n] = daily_p[n] + prcp;
                daily_e[n] = daily_e[n] + pet;
            }
            oreal daily_cumulated = 0;
            for (int d = 0; d < nday; d++) daily_cumulated = daily_cumulated + fmin(daily_p[d], daily_e[d]); /* :79-91 */
            int best = 0;
            oreal best_diff = 0;
            for (int i = 0; i < ncand; i++) {                              /* :95-131 */
                oreal h = 0, sub = 0, pth, ec;
                for (int t = 0; t < P->ntime_step; t++) {
                    oreal prcp, pet;
                    read_forcing(P, row, col, t, &prcp, &pet);
                    gr_interception(prcp, pet, cmax[i], &h, &pth, &ec);   /* :117 */
                    sub = sub + ec;
                }
                oreal diff = FABS(sub - daily_cumulated);
                if (i == 0 || diff < best_diff) { best = i; best_diff = diff; }   /* minloc: first smallest, :142 */
            }
            ci[c] = cmax[best];                                           /* :143 */
        }
    }
    free(daily_p); free(daily_e);
    return 0;
}

/* select case (trim(setup%structure)) forward.f90:43-65, :120-142 */
static void run_structure(const Prob *P, const oreal *par, oreal *st, oreal *qsim, oreal *qdom, oreal *netp) {
    if (P->structure <= OST_GR_A) gr_a_forward(P, par, st, qsim, qdom, netp, NULL);
    else structure_forward(P, par, st, qsim, qdom, netp);
}

/* ============================================================================================
 * base_forward -- forward/forward.f90:1-80
 * ========================================================================================== */
int OSYM(oracle_forward)(const Prob *P, oreal *parameters, const oreal *parameters_bgd, oreal *states,
                         const oreal *states_bgd, oreal *qsim, oreal *fstates, oreal *out_cost,
                         oreal *qsim_domain, oreal *net_prcp_domain) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    if (P->denormalize_forward) {                       /* :33-38 */
        denormalize(P, parameters, O_GNP, P->lb_parameters, P->ub_parameters);
        denormalize(P, states, O_GNS, P->lb_states, P->ub_states);
    }
    oreal *states_imd = (oreal *)malloc(O_GNS * ncell * sizeof(oreal));
    memcpy(states_imd, states, O_GNS * ncell * sizeof(oreal)); /* :41 */
    run_structure(P, parameters, states, qsim, P->save_qsim_domain ? qsim_domain : NULL,
                  P->save_net_prcp_domain ? net_prcp_domain : NULL);   /* :43-65 */
    if (fstates) memcpy(fstates, states, O_GNS * ncell * sizeof(oreal));     /* :71 */
    memcpy(states, states_imd, O_GNS * ncell * sizeof(oreal));              /* :72 */
    free(states_imd);
    compute_cost(P, parameters, parameters_bgd, states, states_bgd, qsim, out_cost); /* :78 */
    return 0;
}

/* ============================================================================================
 * BASE_FORWARD_B -- forward/forward_db.f90:10648-10936, COMPUTE_COST_B :3252-3353
 * ========================================================================================== */
int OSYM(oracle_forward_b)(const Prob *P, oreal *parameters, oreal *parameters_b, const oreal *parameters_bgd,
                           oreal *states, oreal *states_b, const oreal *states_bgd, oreal *qsim, oreal *out_cost) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    const oreal cost_b = R(1.0);
    if (P->structure > OST_GR_A && P->structure != OST_GR_D) return 2;   /* GR_A_FORWARD_B and GR_D_FORWARD_B are restated */
    if (P->denormalize_forward) {                       /* :10697-10703 */
        denormalize(P, parameters, O_GNP, P->lb_parameters, P->ub_parameters);
        denormalize(P, states, O_GNS, P->lb_states, P->ub_states);
    }
    oreal *par0 = (oreal *)malloc(O_GNP * ncell * sizeof(oreal));
    oreal *st0 = (oreal *)malloc(O_GNS * ncell * sizeof(oreal));
    memcpy(par0, parameters, O_GNP * ncell * sizeof(oreal));
    memcpy(st0, states, O_GNS * ncell * sizeof(oreal));
    run_structure(P, parameters, states, qsim, NULL, NULL);            /* :10713 */
    memcpy(states, st0, O_GNS * ncell * sizeof(oreal));                /* :10770 */
    compute_cost(P, parameters, parameters_bgd, states, states_bgd, qsim, out_cost); /* :10821 */
    memcpy(parameters, par0, O_GNP * ncell * sizeof(oreal));           /* POPREAL4ARRAYs :10823-10868 */
    memcpy(states, st0, O_GNS * ncell * sizeof(oreal));
    memset(parameters_b, 0, O_GNP * ncell * sizeof(oreal));            /* :10869 */
    memset(states_b, 0, O_GNS * ncell * sizeof(oreal));                /* :10870 */

    /* COMPUTE_COST_B :3252-3353 */
    oreal jobs_b = cost_b, jreg_b = P->wjreg * cost_b;
    if (P->denormalize_forward) {
        normalize(P, parameters, O_GNP, P->lb_parameters, P->ub_parameters);
        normalize(P, states, O_GNS, P->lb_states, P->ub_states);
    }
    compute_jreg(P, parameters, parameters_bgd, states, states_bgd, parameters_b, states_b, jreg_b);
    if (P->denormalize_forward) {
        for (int i = 0; i < O_GNS; i++)                                /* NORMALIZE_STATES_B :1877-1900 */
            for (size_t c = 0; c < ncell; c++) states_b[c + i * ncell] = states_b[c + i * ncell] / (P->ub_states[i] - P->lb_states[i]);
        for (int i = 0; i < O_GNP; i++)                                /* NORMALIZE_PARAMETERS_B :809-889 */
            for (size_t c = 0; c < ncell; c++) parameters_b[c + i * ncell] = parameters_b[c + i * ncell] / (P->ub_parameters[i] - P->lb_parameters[i]);
    }
    memcpy(parameters, par0, O_GNP * ncell * sizeof(oreal));
    memcpy(states, st0, O_GNS * ncell * sizeof(oreal));                /* :10877-10884 */
    oreal *qsim_b = (oreal *)calloc((size_t)(P->ng > 0 ? P->ng : 1) * P->ntime_step, sizeof(oreal));
    compute_jobs(P, qsim, qsim_b, jobs_b);

    gr_a_forward_b(P, parameters, parameters_b, states, states_b, qsim, qsim_b); /* :10885 */

    if (P->denormalize_forward) {                                      /* :10931-10935 */
        for (int i = 0; i < O_GNS; i++)
            for (size_t c = 0; c < ncell; c++) states_b[c + i * ncell] = (P->ub_states[i] - P->lb_states[i]) * states_b[c + i * ncell];
        for (int i = 0; i < O_GNP; i++)
            for (size_t c = 0; c < ncell; c++) parameters_b[c + i * ncell] = (P->ub_parameters[i] - P->lb_parameters[i]) * parameters_b[c + i * ncell];
    }
    free(qsim_b); free(par0); free(st0);
    return 0;
}

/* ============================================================================================
 * hyper mapping -- mwd_parameters_manipulation.f90:304-362, mwd_states_manipulation.f90:271-329
 * and HYPER_*_TO_*_B forward_db.f90:1434-1537, 2272-2369
 * ========================================================================================== */
static void hyper_to_field(const Prob *P, const oreal *hyper, oreal *field, int nplanes, const oreal *lb, const oreal *ub) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    const int nh = P->nhyper;
    for (int i = 0; i < nplanes; i++) {
        oreal *f = field + i * ncell;
        for (size_t c = 0; c < ncell; c++) f[c] = hyper[0 + nh * i];
        for (int j = 1; j <= P->nd; j++) {
            const oreal *d = P->descriptor + (size_t)(j - 1) * ncell;
            oreal a, b;
            if (P->mapping == OM_HYPER_LINEAR) { a = hyper[j + nh * i]; b = R(1.0); }
            else { a = hyper[2 * j - 1 + nh * i]; b = hyper[2 * j + nh * i]; }
            for (size_t c = 0; c < ncell; c++) f[c] = f[c] + a * POW(d[c], b);
        }
        for (size_t c = 0; c < ncell; c++) f[c] = (ub[i] - lb[i]) * (R(1.0) / (R(1.0) + EXP(-f[c]))) + lb[i];
    }
}

static void hyper_to_field_b(const Prob *P, const oreal *hyper, oreal *hyper_b, oreal *field_b, int nplanes,
                             const oreal *lb, const oreal *ub, const int *zero_planes, int nzero) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    const int nh = P->nhyper;
    oreal *z = (oreal *)malloc(ncell * sizeof(oreal));
    for (int k = 0; k < nzero; k++) memset(field_b + zero_planes[k] * ncell, 0, ncell * sizeof(oreal));
    for (int k = 0; k < nh * nplanes; k++) hyper_b[k] = 0;
    for (int i = nplanes - 1; i >= 0; i--) {
        /* recompute the pre-sigmoid field */
        for (size_t c = 0; c < ncell; c++) z[c] = hyper[0 + nh * i];
        for (int j = 1; j <= P->nd; j++) {
            const oreal *d = P->descriptor + (size_t)(j - 1) * ncell;
            oreal a, b;
            if (P->mapping == OM_HYPER_LINEAR) { a = hyper[j + nh * i]; b = R(1.0); }
            else { a = hyper[2 * j - 1 + nh * i]; b = hyper[2 * j + nh * i]; }
            for (size_t c = 0; c < ncell; c++) z[c] = z[c] + a * POW(d[c], b);
        }
        oreal *fb = field_b + i * ncell;
        for (size_t c = 0; c < ncell; c++) {
            oreal temp = EXP(-z[c]) + R(1.0);
            fb[c] = EXP(-z[c]) * (ub[i] - lb[i]) * fb[c] / (temp * temp);
        }
        for (int j = P->nd; j >= 1; j--) {
            const oreal *d = P->descriptor + (size_t)(j - 1) * ncell;
            oreal a, b;
            if (P->mapping == OM_HYPER_LINEAR) { a = hyper[j + nh * i]; b = R(1.0); }
            else { a = hyper[2 * j - 1 + nh * i]; b = hyper[2 * j + nh * i]; }
            oreal a_b = 0, b_b = 0;
            for (size_t c = 0; c < ncell; c++) a_b = a_b + POW(d[c], b) * fb[c];
            for (size_t c = 0; c < ncell; c++)
                if (!(d[c] <= R(0.0))) b_b = b_b + POW(d[c], b) * LOG(d[c]) * (a * fb[c]);
            if (P->mapping == OM_HYPER_LINEAR) hyper_b[j + nh * i] += a_b;
            else { hyper_b[2 * j + nh * i] += b_b; hyper_b[2 * j - 1 + nh * i] += a_b; }
        }
        oreal s = 0;
        for (size_t c = 0; c < ncell; c++) s = s + fb[c];
        hyper_b[0 + nh * i] += s;
        memset(fb, 0, ncell * sizeof(oreal));
    }
    free(z);
}

/* base_hyper_forward forward.f90:82-157 */
int OSYM(oracle_hyper_forward)(const Prob *P, oreal *parameters, const oreal *hyper_parameters, oreal *states,
                               const oreal *hyper_states, oreal *qsim, oreal *fstates, oreal *out_cost) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    hyper_to_field(P, hyper_parameters, parameters, O_GNP, P->lb_parameters, P->ub_parameters);
    hyper_to_field(P, hyper_states, states, O_GNS, P->lb_states, P->ub_states);
    run_structure(P, parameters, states, qsim, NULL, NULL);
    if (fstates) memcpy(fstates, states, O_GNS * ncell * sizeof(oreal));
    oreal jobs = compute_jobs(P, qsim, NULL, 0); /* hyper_compute_cost mwd_cost.f90:309-348 */
    if (out_cost) { out_cost[0] = jobs + P->wjreg * R(0.0); out_cost[1] = jobs; out_cost[2] = 0; }
    return 0;
}

/* BASE_HYPER_FORWARD_B forward_db.f90:11231-11554 */
int OSYM(oracle_hyper_forward_b)(const Prob *P, oreal *parameters, const oreal *hyper_parameters,
                                 oreal *hyper_parameters_b, oreal *states, const oreal *hyper_states,
                                 oreal *hyper_states_b, oreal *qsim, oreal *out_cost) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    if (P->structure > OST_GR_A && P->structure != OST_GR_D) return 2;
    hyper_to_field(P, hyper_parameters, parameters, O_GNP, P->lb_parameters, P->ub_parameters);
    hyper_to_field(P, hyper_states, states, O_GNS, P->lb_states, P->ub_states);
    oreal *st0 = (oreal *)malloc(O_GNS * ncell * sizeof(oreal));
    memcpy(st0, states, O_GNS * ncell * sizeof(oreal));
    run_structure(P, parameters, states, qsim, NULL, NULL);
    oreal jobs = compute_jobs(P, qsim, NULL, 0);
    if (out_cost) { out_cost[0] = jobs; out_cost[1] = jobs; out_cost[2] = 0; }
    oreal *qsim_b = (oreal *)calloc((size_t)(P->ng > 0 ? P->ng : 1) * P->ntime_step, sizeof(oreal));
    compute_jobs(P, qsim, qsim_b, R(1.0));
    oreal *par_b = (oreal *)calloc(O_GNP * ncell, sizeof(oreal));
    oreal *st_b = (oreal *)calloc(O_GNS * ncell, sizeof(oreal));
    memcpy(states, st0, O_GNS * ncell * sizeof(oreal));
    gr_a_forward_b(P, parameters, par_b, states, st_b, qsim, qsim_b);
    hyper_to_field_b(P, hyper_states, hyper_states_b, st_b, O_GNS, P->lb_states, P->ub_states, NULL, 0);
    const int zp[2] = {OP_BETA, OP_ALPHA}; /* forward_db.f90:1489-1490 */
    hyper_to_field_b(P, hyper_parameters, hyper_parameters_b, par_b, O_GNP, P->lb_parameters, P->ub_parameters, zp, 2);
    free(qsim_b); free(par_b); free(st_b); free(st0);
    return 0;
}

/* ============================================================================================
 * compute_multiple_run -- routine/mw_multiple_run.f90:40-119
 * ========================================================================================== */
int OSYM(oracle_multiple_run)(const Prob *P, const oreal *parameters, const oreal *states, const oreal *sample,
                              const int *ind, int nvar, int ns, oreal *res_cost, oreal *res_qsim, int nthreads) {
    const size_t ncell = (size_t)P->nrow * P->ncol;
    const size_t nq = (size_t)P->ng * P->ntime_step;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 1)
#endif
    for (int i = 0; i < ns; i++) {
        oreal *par = (oreal *)malloc(O_GNP * ncell * sizeof(oreal));
        oreal *st = (oreal *)malloc(O_GNS * ncell * sizeof(oreal));
        oreal *qsim = (oreal *)malloc((nq > 0 ? nq : 1) * sizeof(oreal));
        memcpy(par, parameters, O_GNP * ncell * sizeof(oreal));
        memcpy(st, states, O_GNS * ncell * sizeof(oreal));
        for (int v = 0; v < nvar; v++) { /* set_sample_to_parameters_states :40-65: whole plane */
            int k = ind[v] - 1;
            oreal *plane = (k < O_GNP) ? par + (size_t)k * ncell : st + (size_t)(k - O_GNP) * ncell;
            for (size_t c = 0; c < ncell; c++) plane[c] = sample[v + (size_t)nvar * i];
        }
        oreal cost3[3];
        OSYM(oracle_forward)(P, par, parameters, st, states, qsim, NULL, cost3, NULL, NULL);
        res_cost[i] = cost3[0];
        if (res_qsim) memcpy(res_qsim + nq * i, qsim, nq * sizeof(oreal));
        free(par); free(st); free(qsim);
    }
    return 0;
}
