/*
 * smash_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the reference's forward / adjoint solver hot path
 * (DassHydro-dev/smash v0.5.0, Fortran 90 + Tapenade 3.16).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (smash_b200/) never does.
 *
 * The file is compiled twice: -DORACLE_DOUBLE=0 (float32, the reference's real kind,
 * md_constant.f90:29) and -DORACLE_DOUBLE=1 (float64 referee for finite-difference /
 * Taylor tests).  Exported symbols carry the suffix _f32 / _f64.
 *
 * Parity status: PINNED.  The reference cannot be compiled here (no Fortran compiler,
 * SURVEY.md 8c); the float32 build is checked against the reference's own golden file
 * smash/tests/baseline.hdf5 (run.cost, multiple_run.{cost,qsim}, mutiple_run.slc_*,
 * optimize.* after one L-BFGS-B iteration) through tests/golden/cance_golden.npz.
 * Exception -- parity UNPINNED: the forward restatement of the structures gr-b, gr-c, gr-d and
 * vic-a (structure_forward in smash_oracle.c).  The reference's tests and golden file run gr-a
 * only, so there is no vector to pin them to; what is checked instead (tests/test_oracle_structures.py):
 * gr-b with ci -> 0 reproduces the pinned gr-a run, float32 against float64 builds, forward.f90's bookkeeping.
 *
 * All arrays use the reference's memory layout: Fortran order, 1-based index VALUES.
 *   (nrow,ncol)      -> a[row-1 + (col-1)*nrow]
 *   (nrow,ncol,n)    -> a[row-1 + (col-1)*nrow + k*nrow*ncol]
 *   path(2,nrow*ncol)-> path[2*i] = row, path[2*i+1] = col
 *   gauge_pos(ng,2)  -> gauge_pos[g] = row, gauge_pos[g+ng] = col
 *   qsim(ng,T)       -> qsim[g + ng*t]
 */
#ifndef SMASH_ORACLE_H
#define SMASH_ORACLE_H

#ifndef ORACLE_DOUBLE
#define ORACLE_DOUBLE 0
#endif
#if ORACLE_DOUBLE
typedef double oreal;
#define OSYM(name) name##_f64
#else
typedef float oreal;
#define OSYM(name) name##_f32
#endif

#define O_GNP 16 /* md_constant.f90:32 */
#define O_GNS 8  /* md_constant.f90:33 */

/* parameter planes (md_constant.f90:35-55) */
enum { OP_CI = 0, OP_CP, OP_BETA, OP_CFT, OP_CST, OP_ALPHA, OP_EXC, OP_B, OP_CUSL1, OP_CUSL2, OP_CLSL,
       OP_KS, OP_DS, OP_DSM, OP_WS, OP_LR };
/* state planes (md_constant.f90:57-69) */
enum { OS_HI = 0, OS_HP, OS_HFT, OS_HST, OS_HUSL1, OS_HUSL2, OS_HLSL, OS_HLR };

/* setup%structure (mwd_setup.f90:113, _constant.py STRUCTURE_PARAMETERS) */
enum { OST_GR_A = 1, OST_GR_B = 2, OST_GR_C = 3, OST_GR_D = 4, OST_VIC_A = 5 };

/* jobs_fun codes (mwd_cost.f90:98-131) */
enum { OJ_NSE = 1, OJ_KGE = 2, OJ_KGE2 = 3, OJ_SE = 4, OJ_RMSE = 5, OJ_LOGARITHMIC = 6,
       /* signatures (mwd_cost.f90:117-122, 772-970): continuous Crc, Cfp2/10/50/90; event-based Erc, Elt, Epf */
       OJ_CRC = 7, OJ_CFP2 = 8, OJ_CFP10 = 9, OJ_CFP50 = 10, OJ_CFP90 = 11, OJ_ERC = 12, OJ_ELT = 13, OJ_EPF = 14 };
/* jreg_fun codes (mwd_cost.f90:200-240) */
enum { OR_PRIOR = 1, OR_SMOOTHING = 2, OR_HARD_SMOOTHING = 3 };
/* mapping codes (mwd_parameters_manipulation.f90:330-342) */
enum { OM_NONE = 0, OM_HYPER_LINEAR = 1, OM_HYPER_POLYNOMIAL = 2 };

typedef struct {
    /* SetupDT (mwd_setup.f90:107-157) */
    int ntime_step, nd;
    oreal dt;
    int sparse_storage, save_qsim_domain, save_net_prcp_domain;
    /* MeshDT (mwd_mesh.f90:45-72) */
    int nrow, ncol, ng, nac;
    oreal dx;
    const int *flwdir, *flwacc, *active_cell, *local_active_cell; /* (nrow,ncol) */
    const int *path;                 /* (2,nrow*ncol) 1-based, <=0 = unused slot */
    const int *gauge_pos;            /* (ng,2) 1-based */
    const int *rowcol_to_ind_sparse; /* (nrow,ncol) 1-based, only if sparse_storage */
    const oreal *area;               /* (ng) */
    /* Input_DataDT (mwd_input_data.f90:32-50) */
    const oreal *prcp, *pet;  /* (nrow,ncol,T) or sparse (nac,T) */
    const oreal *qobs;        /* (ng,T) */
    const oreal *descriptor;  /* (nrow,ncol,nd) */
    /* Optimize_SetupDT (mwd_setup.f90:57-105) */
    int njf; const int *jobs_fun; const oreal *wjobs_fun;
    int njr; const int *jreg_fun; const oreal *wjreg_fun;
    oreal wjreg;
    int denormalize_forward, optimize_start_step /* 1-based */, mapping, nhyper;
    const int *optim_parameters; /* (16) */
    const int *optim_states;     /* (8) */
    const oreal *lb_parameters, *ub_parameters; /* (16) */
    const oreal *lb_states, *ub_states;         /* (8) */
    const oreal *wgauge;                        /* (ng) */
    /* signature objectives only (may be NULL otherwise) */
    const oreal *mean_prcp;                     /* Input_DataDT%mean_prcp (ng,T) */
    const int *mask_event;                      /* Optimize_SetupDT%mask_event (ng,T): event number of every step, 0 = none */
    int structure;                              /* OST_* (0 = gr-a); the adjoint and multiple_run entry points take gr-a only */
} OSYM(OProblem);

#ifdef __cplusplus
extern "C" {
#endif

/* base_forward (forward.f90:1-80).  parameters (nrow,ncol,16) / states (nrow,ncol,8) are updated in
 * place exactly like the reference (denormalised on exit if denormalize_forward; states restored).
 * out_cost = {cost, cost_jobs, cost_jreg}.  qsim_domain / net_prcp_domain may be NULL. */
int OSYM(oracle_forward)(const OSYM(OProblem) *P, oreal *parameters, const oreal *parameters_bgd,
                         oreal *states, const oreal *states_bgd, oreal *qsim, oreal *fstates,
                         oreal *out_cost, oreal *qsim_domain, oreal *net_prcp_domain);

/* BASE_FORWARD_B (forward_db.f90:10648-10936) with cost_b = 1: store-all tape like Tapenade. */
int OSYM(oracle_forward_b)(const OSYM(OProblem) *P, oreal *parameters, oreal *parameters_b,
                           const oreal *parameters_bgd, oreal *states, oreal *states_b,
                           const oreal *states_bgd, oreal *qsim, oreal *out_cost);

/* base_hyper_forward (forward.f90:82-157); hyper_* are (nhyper,1,16|8). */
int OSYM(oracle_hyper_forward)(const OSYM(OProblem) *P, oreal *parameters, const oreal *hyper_parameters,
                               oreal *states, const oreal *hyper_states, oreal *qsim, oreal *fstates,
                               oreal *out_cost);

/* BASE_HYPER_FORWARD_B (forward_db.f90:11231-11554) with cost_b = 1. */
int OSYM(oracle_hyper_forward_b)(const OSYM(OProblem) *P, oreal *parameters, const oreal *hyper_parameters,
                                 oreal *hyper_parameters_b, oreal *states, const oreal *hyper_states,
                                 oreal *hyper_states_b, oreal *qsim, oreal *out_cost);

/* compute_multiple_run (mw_multiple_run.f90:68-119).  sample is F(nvar,ns); ind is 1-based into the
 * 24 stacked planes; res_qsim F(ng,T,ns) may be NULL.  nthreads mirrors setup%ncpu (OpenMP). */
int OSYM(oracle_multiple_run)(const OSYM(OProblem) *P, const oreal *parameters, const oreal *states,
                              const oreal *sample, const int *ind, int nvar, int ns, oreal *res_cost,
                              oreal *res_qsim, int nthreads);

/* Pieces exported for unit tests. */
/* adjust_interception_store (mw_interception_store.f90:19-160): ci (nrow,ncol) is updated on the computed cells only */
int OSYM(oracle_adjust_interception_store)(const OSYM(OProblem) *P, int nday, const int *day_index, oreal *ci);
void OSYM(oracle_gr_production)(oreal pn, oreal en, oreal cp, oreal beta, oreal *hp, oreal *pr, oreal *perc);
void OSYM(oracle_gr_transfer)(oreal n, oreal prcp, oreal pr, oreal ct, oreal *ht, oreal *q);
oreal OSYM(oracle_nse)(const oreal *x, const oreal *y, int n);
/* compute_jobs (mwd_cost.f90:37-156) on a given qsim(ng,T); qsim_b (may be NULL) receives COMPUTE_JOBS_B with jobs_b = 1 */
oreal OSYM(oracle_compute_jobs)(const OSYM(OProblem) *P, const oreal *qsim, oreal *qsim_b);
oreal OSYM(oracle_kge)(const oreal *x, const oreal *y, int n);

#ifdef __cplusplus
}
#endif
#endif
