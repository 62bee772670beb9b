/*
 * smash_b200.h -- C ABI of libsmash_b200.so, the B200 (sm_100a) replacement for the
 * f90wrap-exposed solver entry points of DassHydro-dev/smash v0.5.0.
 *
 * Every entry point below names the reference interface it replaces (paths relative to
 * /root/reference/smash/solver/).  All arrays are caller-owned HOST memory in the reference's
 * own layout: Fortran order, real(sp) = float, integer(4) = int32_t, and 1-BASED index values
 * in `path`, `gauge_pos`, `rowcol_to_ind_sparse` and `ind_parameters_states` (what the Fortran
 * side holds in memory; the f90wrap Python layer shows them 0-based, _f90wrap_decorator.py:72-106).
 *   (nrow,ncol)       a[(row-1) + (col-1)*nrow]
 *   (nrow,ncol,T)     a[(row-1) + (col-1)*nrow + t*nrow*ncol]
 *   (nac,T) sparse    a[k + nac*t]        k = rowcol_to_ind_sparse(row,col) - 1
 *   path(2,nrow*ncol) path[2*i] = row, path[2*i+1] = col
 *   gauge_pos(ng,2)   gauge_pos[g] = row, gauge_pos[g+ng] = col
 *   qsim(ng,T)        qsim[g + ng*t]
 *
 * Return value: 0 on success, non-zero on error (message via smash_b200_last_error()).  There is
 * no CPU fallback: without a CUDA device every compute entry point fails with SMASH_B200_ENODEV.
 * Structures: gr-a (the structure named by BASELINE.json's configs) behind every entry point; gr-b, gr-c, gr-d and vic-a
 * (forward/md_forward_structure.f90:216-931) behind smash_b200_forward, smash_b200_compute_multiple_run and the plan API's forward
 * run; gr-d also behind smash_b200_forward_b (GR_D_FORWARD_B, forward_db.f90:9604-9797).  The adjoint of gr-b / gr-c / vic-a and the
 * descriptor mappings of the four answer SMASH_B200_EUNSUPPORTED.
 */
#ifndef SMASH_B200_H
#define SMASH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMASH_B200_GNP 16 /* global/md_constant.f90:32 */
#define SMASH_B200_GNS 8  /* global/md_constant.f90:33 */

enum { SMASH_B200_OK = 0, SMASH_B200_EINVAL = 1, SMASH_B200_ENODEV = 2, SMASH_B200_ECUDA = 3,
       SMASH_B200_EUNSUPPORTED = 4, SMASH_B200_ENOMEM = 5 };

/* setup%structure (derived_type/mwd_setup.f90:113) */
enum { SMASH_STRUCTURE_GR_A = 1, SMASH_STRUCTURE_GR_B = 2, SMASH_STRUCTURE_GR_C = 3, SMASH_STRUCTURE_GR_D = 4,
       SMASH_STRUCTURE_VIC_A = 5 };
/* setup%optimize%jobs_fun(:) (optimize/mwd_cost.f90:98-131) */
enum { SMASH_JOBS_NSE = 1, SMASH_JOBS_KGE = 2, SMASH_JOBS_KGE2 = 3, SMASH_JOBS_SE = 4, SMASH_JOBS_RMSE = 5,
       SMASH_JOBS_LOGARITHMIC = 6,
       /* signature-based objectives (optimize/mwd_cost.f90:117-122, 770-970); forward only: forward_b rejects them */
       SMASH_JOBS_CRC = 7, SMASH_JOBS_CFP2 = 8, SMASH_JOBS_CFP10 = 9, SMASH_JOBS_CFP50 = 10, SMASH_JOBS_CFP90 = 11,
       SMASH_JOBS_ERC = 12, SMASH_JOBS_ELT = 13, SMASH_JOBS_EPF = 14 };
/* setup%optimize%jreg_fun(:) (optimize/mwd_cost.f90:200-240) */
enum { SMASH_JREG_PRIOR = 1, SMASH_JREG_SMOOTHING = 2, SMASH_JREG_HARD_SMOOTHING = 3 };
/* setup%optimize%mapping (routine/mwd_parameters_manipulation.f90:330-342) */
enum { SMASH_MAPPING_NONE = 0, SMASH_MAPPING_HYPER_LINEAR = 1, SMASH_MAPPING_HYPER_POLYNOMIAL = 2 };

/* Index of each field inside SmashParameters.v / SmashStates.v = position in GPARAMETERS_NAME /
 * GSTATES_NAME (global/md_constant.f90:35-69). */
enum { SMASH_P_CI = 0, SMASH_P_CP, SMASH_P_BETA, SMASH_P_CFT, SMASH_P_CST, SMASH_P_ALPHA, SMASH_P_EXC, SMASH_P_B,
       SMASH_P_CUSL1, SMASH_P_CUSL2, SMASH_P_CLSL, SMASH_P_KS, SMASH_P_DS, SMASH_P_DSM, SMASH_P_WS, SMASH_P_LR };
enum { SMASH_S_HI = 0, SMASH_S_HP, SMASH_S_HFT, SMASH_S_HST, SMASH_S_HUSL1, SMASH_S_HUSL2, SMASH_S_HLSL, SMASH_S_HLR };

/* SetupDT + Optimize_SetupDT (derived_type/mwd_setup.f90:57-157): the fields the solver reads. */
typedef struct SmashSetup {
    int32_t structure;
    float dt;
    int32_t ntime_step;
    int32_t nd;
    int32_t ncpu; /* accepted for signature fidelity; members run on the GPU */
    int32_t sparse_storage, save_qsim_domain, save_net_prcp_domain;
    /* setup%optimize */
    int32_t njf;
    const int32_t *jobs_fun; /* (njf) SMASH_JOBS_* */
    const float *wjobs_fun;  /* (njf) */
    int32_t njr;
    const int32_t *jreg_fun; /* (njr) SMASH_JREG_* */
    const float *wjreg_fun;  /* (njr) */
    float wjreg;
    int32_t mapping, denormalize_forward, nhyper;
    int32_t optimize_start_step; /* 1-based */
    int32_t optim_parameters[SMASH_B200_GNP];
    int32_t optim_states[SMASH_B200_GNS];
    float lb_parameters[SMASH_B200_GNP], ub_parameters[SMASH_B200_GNP];
    float lb_states[SMASH_B200_GNS], ub_states[SMASH_B200_GNS];
    const float *wgauge; /* (ng) */
    const int32_t *mask_event; /* (ng,T) setup%optimize%mask_event: event number of every step, 0 = none; signature objectives only */
} SmashSetup;

/* MeshDT (derived_type/mwd_mesh.f90:45-72) */
typedef struct SmashMesh {
    float dx;
    int32_t nrow, ncol, ng, nac;
    const int32_t *flwdir, *flwacc, *active_cell; /* (nrow,ncol) */
    const int32_t *local_active_cell;             /* (nrow,ncol) or NULL = all 1 */
    const int32_t *path;                          /* (2,nrow*ncol) */
    const int32_t *gauge_pos;                     /* (ng,2) */
    const int32_t *rowcol_to_ind_sparse;          /* (nrow,ncol), used when sparse_storage */
    const float *area;                            /* (ng) */
} SmashMesh;

/* Input_DataDT (derived_type/mwd_input_data.f90:32-50) */
typedef struct SmashInputData {
    const float *qobs;                    /* (ng,T) */
    const float *prcp, *pet;              /* (nrow,ncol,T)  when !sparse_storage */
    const float *sparse_prcp, *sparse_pet; /* (nac,T)       when  sparse_storage */
    const float *descriptor;              /* (nrow,ncol,nd) */
    /* 0 = forcing is re-uploaded on every call (reference semantics: arrays may change between
     * calls).  Non-zero = caller's promise that (pointer, version) identifies immutable content,
     * so the device-resident [block][tick][cell] copy is reused. */
    uint64_t forcing_version;
    const float *mean_prcp;               /* (ng,T) catchment-mean precipitation (routine/mw_forcing_statistic.f90:18-75); signature
                                           * objectives only */
} SmashInputData;

/* ParametersDT / StatesDT (derived_type/mwd_parameters.f90:58-78, mwd_states.f90:49-60): one
 * (nrow,ncol) plane per field; Hyper_* variants hold (nhyper,1) planes. */
typedef struct SmashParameters { float *v[SMASH_B200_GNP]; } SmashParameters;
typedef struct SmashStates { float *v[SMASH_B200_GNS]; } SmashStates;

/* OutputDT (derived_type/mwd_output.f90:36-57); any pointer may be NULL = not wanted. */
typedef struct SmashOutput {
    float *qsim;                                         /* (ng,T) */
    float *qsim_domain, *sparse_qsim_domain;             /* (nrow,ncol,T) / (nac,T) */
    float *net_prcp_domain, *sparse_net_prcp_domain;     /* (nrow,ncol,T) / (nac,T) */
    float cost, cost_jobs, cost_jreg;
    SmashStates fstates;
} SmashOutput;

/* ---- drop-in entry points --------------------------------------------------------------------- */

/* replaces mw_forward::forward (forward/mw_forward.f90:18-39 -> base_forward forward/forward.f90:1-80) */
int smash_b200_forward(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *input_data,
                       SmashParameters *parameters, const SmashParameters *parameters_bgd, SmashStates *states,
                       const SmashStates *states_bgd, SmashOutput *output, float *cost);

/* replaces mw_forward::forward_b (forward/mw_forward.f90:41-68 -> BASE_FORWARD_B forward/forward_db.f90:10648-10936).
 * parameters_b / states_b are overwritten (zero, then accumulate); *cost_b is the seed (callers pass 1).
 * parameters_bgd_b, states_bgd_b and output_b of the Fortran signature carry no information and are omitted. */
int smash_b200_forward_b(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *input_data,
                         SmashParameters *parameters, SmashParameters *parameters_b,
                         const SmashParameters *parameters_bgd, SmashStates *states, SmashStates *states_b,
                         const SmashStates *states_bgd, SmashOutput *output, float *cost, float *cost_b);

/* replaces mw_forward::hyper_forward (forward/mw_forward.f90:99-123 -> base_hyper_forward forward/forward.f90:82-157) */
int smash_b200_hyper_forward(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *input_data,
                             SmashParameters *parameters, const SmashParameters *hyper_parameters,
                             const SmashParameters *hyper_parameters_bgd, SmashStates *states,
                             const SmashStates *hyper_states, const SmashStates *hyper_states_bgd,
                             SmashOutput *output, float *cost);

/* replaces mw_forward::hyper_forward_b (forward/mw_forward.f90:125-152 -> BASE_HYPER_FORWARD_B forward_db.f90:11231-11554) */
int smash_b200_hyper_forward_b(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *input_data,
                               SmashParameters *parameters, const SmashParameters *hyper_parameters,
                               SmashParameters *hyper_parameters_b, SmashStates *states,
                               const SmashStates *hyper_states, SmashStates *hyper_states_b, SmashOutput *output,
                               float *cost, float *cost_b);

/* replaces mw_multiple_run::compute_multiple_run (routine/mw_multiple_run.f90:68-119).
 * sample F(nvar,ns); ind_parameters_states (nvar) 1-based into the 16+8 stacked fields;
 * res_cost (ns); res_qsim F(ng,T,ns) or NULL (the reference's size-0 array, :113). */
int smash_b200_compute_multiple_run(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *input_data,
                                    const SmashParameters *parameters, const SmashStates *states, SmashOutput *output,
                                    const float *sample, const int32_t *ind_parameters_states, int32_t nvar, int32_t ns,
                                    float *res_cost, float *res_qsim);

/* ---- library services ------------------------------------------------------------------------- */
const char *smash_b200_last_error(void);
const char *smash_b200_version(void);
int smash_b200_device_count(void);           /* number of CUDA devices, 0 if none / no driver */
int smash_b200_set_device(int device);       /* device used by subsequent calls of this thread */
void smash_b200_clear_cache(void);           /* drop cached mesh plans and device-resident forcing */
/* options (also readable from the environment as SMASH_B200_<NAME>):
 *   "math"  1 (default) = MUFU reciprocal / rsqrt with the cancellation-free transfer formula (closer to the float64 solution
 *           than the reference's own float32 arithmetic); 0 = IEEE division / sqrt + libm tanhf in the
 *           reference's statement order (runs on the fused engine)
 *   "engine" / "ensemble_engine"  1 = split engine (default with math = 1), 0 = fused tick wavefront; the second name is the
 *           same choice for compute_multiple_run
 *   "member_budget_mb"  device memory per ensemble launch (default 16384)
 *   "pin_host"  1 = page-lock large caller-owned host arrays in place on first use (PCIe-speed copies); the caller must then
 *           call smash_b200_clear_cache() before freeing them; off by default
 *   "stream" (default 1), "stream_min_mb" (256)  forward runs with at least that much sparse forcing are cut into 256-step
 *           windows so that upload, kernels and the download of sparse_qsim_domain overlap on three streams
 *   "adjoint_checkpoint"  -1 (default) = gradient runs whose store-all tape would exceed "tape_budget_mb" (16384) run their
 *           reverse sweep window by window from checkpointed states (one forward replay per 256-step window); 1 = always,
 *           0 = never
 *   "window_pass" (default 0), "shallow_acc" (32)  experimental forward path for large domains: reservoirs and the routing of
 *           the cells with flwacc <= shallow_acc in one pass over 8-step windows (window_kernels.cu), chain scans over the rest
 *   "fuse_export" (default 4)  warps per routing CTA that also write the routed cells' series to the [t][cell] layout
 * Build-time options are part of the plan cache key: changing one makes a new plan. */
int smash_b200_set_option(const char *name, long long value);

/* ---- device-resident plan API (bench / advanced callers) --------------------------------------
 * A plan owns the level-ordered topology of one mesh and every device buffer; run_* launch the
 * kernels only (inputs already in HBM) and return the device time measured with CUDA events on the
 * plan's stream. */
typedef struct SmashPlan SmashPlan;

int smash_b200_plan_create(const SmashSetup *setup, const SmashMesh *mesh, int32_t nmember, SmashPlan **plan);
void smash_b200_plan_destroy(SmashPlan *plan);
/* forcing to the device (sparse [t][k] arrays whose order is the cell order are used in place, anything else is packed to
 * [t][cell] once); qobs too if present */
int smash_b200_plan_set_forcing(SmashPlan *plan, const SmashSetup *setup, const SmashInputData *input_data);
/* member-major parameter / state values: uniform_sample F(nvar,nmember) applied on top of the planes */
int smash_b200_plan_set_fields(SmashPlan *plan, const SmashParameters *parameters, const SmashStates *states,
                               const float *sample, const int32_t *ind_parameters_states, int32_t nvar);
int smash_b200_plan_run_forward(SmashPlan *plan, float *elapsed_ms);
int smash_b200_plan_run_gradient(SmashPlan *plan, float *elapsed_ms_forward, float *elapsed_ms_reverse);
/* regionalisation step with everything resident on the device: hyper-parameters (Hyper_ParametersDT / Hyper_StatesDT,
 * (nhyper,1) planes) -> field planes by the mapping kernel (routine/mwd_parameters_manipulation.f90:304-362), forward and
 * reverse sweeps, hyper adjoint reductions (forward/forward_db.f90:1434-1537).  hyper_b: [7][nhyper] gradient of cp, cft,
 * exc, lr, hp, hft, hlr; ms: device time of the mapping, forward sweep, reverse sweep, reductions */
int smash_b200_plan_run_hyper_gradient(SmashPlan *plan, const SmashSetup *setup, const SmashInputData *input_data,
                                       const SmashParameters *hyper_parameters, const SmashStates *hyper_states, float *hyper_b,
                                       float ms[4]);
int smash_b200_plan_get_qsim(SmashPlan *plan, float *qsim /* F(ng,T,nmember) */, float *cost /* (nmember) */);
int smash_b200_plan_get_gradient(SmashPlan *plan, SmashParameters *parameters_b, SmashStates *states_b);
/* sum of the device-resident q of the last run over all active cell-steps (size-independent checksum) */
int smash_b200_plan_checksum(SmashPlan *plan, double *sum_q);
/* device time (ms, CUDA events on the plan's stream) of the kernels of the last run: [0] vertical_forward,
 * [1] route_forward, [2] rows_to_domain, [3] route_adjoint, [4] vertical_adjoint; -1 where a kernel did not run */
int smash_b200_plan_kernel_times(SmashPlan *plan, float ms[5]);
/* named facts about a plan: "engine" (0 fused, 1 split), "routed_cells", "source_cells", "inflow_edges", "checkpoint" (1: the
 * adjoint runs window by window from checkpoints), "route_window" (steps), "route_windows", "tape_bytes" (what a gradient
 * run keeps between its sweeps, after the first gradient run), "window_pass", "deep_cells", "shallow_cells"; -1 if unknown */
double smash_b200_plan_stat(const SmashPlan *plan, const char *name);
/* topology facts: [0]=ncell_active [1]=nblocks [2]=block_size [3]=max in-block skew [4]=total ticks over blocks
 * [5]=cross-block edges [6]=pit pairs [7]=kernel launches of the last run_* call [8]=critical path over blocks in
 * ticks [9]=longest chain of dependent blocks [10..11]=reserved */
int smash_b200_plan_info(const SmashPlan *plan, int64_t info[12]);

/* level ordering exposed for the bit-exact mesh tests: order[k] = 0-based flat (row + col*nrow) index of the
 * k-th cell in device order, block_of/offset_of its block and in-block skew (n = number of active cells) */
int smash_b200_plan_order(const SmashPlan *plan, int32_t *order, int32_t *block_of, int32_t *offset_of);
/* same ordering computed on the host only (no CUDA device needed): info as smash_b200_plan_info, block = cells per
 * CTA (0 = automatic) */
int smash_b200_mesh_order(const SmashSetup *setup, const SmashMesh *mesh, int32_t block, int64_t info[12],
                          int32_t *order, int32_t *block_of, int32_t *offset_of);

/* Heavy-path decomposition used by the split engine (reservoir pass per cell + routing pass per chain), computed on
 * the host only.  Cells are numbered j = 0..n-1 in `path` order (md_forward_structure.f90:82-92).
 *   info[0] n computed cells, [1] chains, [2] pit pairs, [3] largest dependency height of a chain, [4] longest chain,
 *   [5] longest dependency path in cells, [6] source cells (flwacc == 1), [7] bit 0: sparse arrays are usable in place, bits 1..: number of longest
 *   chains that run on dedicated warps (they are the last chain tasks and may feed tasks with a smaller index)
 *   cell[j]      flat rect index row + col*nrow of cell j                                    (n entries)
 *   task_of[j]   task (chain or pair, execution order) that routes cell j, -1 for lone source cells
 *   pos_of[j]    position of the cell inside its task, upstream -> downstream
 *   down_of[j]   cell that gathers j in upstream_discharge (md_routing_operator.f90:37-53), -1 if none
 * Returns SMASH_B200_EUNSUPPORTED when the mesh needs the fused engine. */
int smash_b200_mesh_chains(const SmashMesh *mesh, int64_t info[8], int32_t *cell, int32_t *task_of, int32_t *pos_of,
                           int32_t *down_of);
/* ---- preprocessing on the device (meshing, input pipeline) ----------------- */

/* replaces mw_meshing::flow_accumulation (smash/mesh/mw_meshing.f90:204-233 with fill_nipd :111-152 and
 * downstream_cell_flwacc :154-202).  flwdir / flwacc: (nrow,ncol) Fortran order; mask: (nrow,ncol) or NULL -- cells with
 * mask == 0 neither give nor receive (a catchment window).  Integer-exact. */
int smash_b200_flow_accumulation(int32_t nrow, int32_t ncol, const int32_t *flwdir, const int32_t *mask, int32_t *flwacc);

/* replaces mw_mask::mask_upstream_cells (smash/solver/routine/mw_mask.f90:11-55) for every gauge of the mesh:
 * mask (nrow,ncol,ng) bytes, 1 = the cell drains through the gauge cell. */
int smash_b200_gauge_masks(const SmashMesh *mesh, uint8_t *mask);

/* replaces mw_forcing_statistic::compute_mean_forcing (smash/solver/routine/mw_forcing_statistic.f90:18-75):
 * mean_prcp / mean_pet (ng,T) = mean over the cells upstream of each gauge whose value is >= 0 (either may be NULL). */
int smash_b200_compute_mean_forcing(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *input_data, float *mean_prcp,
                                    float *mean_pet);

/* replaces mw_interception_store::adjust_interception_store (smash/solver/routine/mw_interception_store.f90:19-160), the
 * calibration of the interception capacity of gr-b / gr-c on sub-daily runs: ci (nrow,ncol) of every computed cell := the value
 * among 0.1, 0.2 .. 4.9 mm whose cumulated sub-daily interception evaporation is closest to the cumulated daily one.
 * day_index (ntime_step): day number of every step; nday >= the number of days it holds.  Other cells of ci are left untouched.
 * kernel_ms (may be NULL): device time of the search. */
int smash_b200_adjust_interception_store(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *input_data, int32_t nday,
                                         const int32_t *day_index, float *ci, float *kernel_ms);

/* ---- the ANN mapping's Dense layers on the tensor cores --------------------- */

/* replaces Net._forward_pass (smash/core/net.py:281-299) for a chain of Dense (+ Activation) layers (net.py:579-688, 458-498):
 * y = act_L(... act_1(x W_1 + b_1) ...) with every contraction a TF32 tcgen05 GEMM (float32 accumulation) and bias + activation
 * fused into its epilogue; the whole chain stays on the device.
 * x (nrows, sizes[0]) row-major host array; weight[l] (sizes[l], sizes[l+1]) row-major; bias[l] (sizes[l+1]);
 * activation[l]: 0 none, 1 relu, 2 sigmoid, 3 tanh, 4 leaky_relu(0.2), 5 elu(0.1), 6 selu, 7 softplus; y (nrows, sizes[nlayer]).
 * ms / flops (may be NULL): device time of the layers and 2 x multiply-adds. */
int smash_b200_mlp_forward(int64_t nrows, int32_t nlayer, const int32_t *sizes, const float *x, const float *const *weight,
                           const float *const *bias, const int32_t *activation, float *y, float *ms, double *flops);

/* The same chain kept on the device between the forward and the backward pass of a training epoch (Net._fit_d2p,
 * smash/core/net.py:353-415).  run_forward: x may be NULL after the first call (the rows stay on the device); y may be NULL.
 * run_backward replaces Net._backward_pass (net.py:301-303) for the chain: from grad_y = d loss / d y (nrows, sizes[nlayer]) it
 * returns grad_weight[l] (sizes[l], sizes[l+1]) = a_l^T g and grad_bias[l] (sizes[l+1]) = column sums of g
 * (Dense._backward_pass net.py:672-685; the optimiser update stays with the caller), g being pushed back through the
 * activations and W^T layer by layer; both contractions are TF32 tensor-core GEMMs. */
typedef struct SmashMlp SmashMlp;
int smash_b200_mlp_create(int64_t nrows, int32_t nlayer, const int32_t *sizes, const int32_t *activation, SmashMlp **mlp);
void smash_b200_mlp_destroy(SmashMlp *mlp);
int smash_b200_mlp_run_forward(SmashMlp *mlp, const float *x, const float *const *weight, const float *const *bias, float *y, float *ms,
                               double *flops);
int smash_b200_mlp_run_backward(SmashMlp *mlp, const float *grad_y, float *const *grad_weight, float *const *grad_bias, float *ms);

/* Host-only: the ticket schedule of the tick pass (tick_kernels.cu) for this mesh, dealt to nwarp warps and replayed on the
 * host the way the device walks it.  info: [0] cells [1] tiles [2] reaches [3] largest stage [4] shallow routed cells
 * [5] deep cells [6] pit cells [7] units per warp [8] longest deep chain [9] tickets replayed [10] 1 = the schedule is
 * consistent (every ticket only reads smaller keys, the replay completes) [11] cells that publish an exchange block.
 * Replaces nothing in the reference (the time x space loop of md_forward_structure.f90:82-214 is sequential); test aid. */
int smash_b200_mesh_tick_schedule(const SmashMesh *mesh, int32_t shallow_acc, int32_t nwarp, int32_t nwin, int64_t info[12],
                                  int32_t *unit_of, int32_t *sigma_of);

/* ---- the one collective of the path ----------------------------------------------------
 * One process per GPU.  The reference has no distributed layer; the regionalised multi-catchment calibration (shared
 * hyper-parameters, catchments spread over the GPUs) sums (cost, gradient) -- a few hundred values -- over the ranks once per
 * evaluation, and a sharded ensemble gathers its per-member costs.  NCCL (libnccl.so.2, opened at run time) over NVLink /
 * NVSwitch; the 128-byte unique id is created by rank 0 and handed to the other ranks by the caller. */
typedef struct SmashComm SmashComm;
int smash_b200_comm_unique_id(char id[128]);
int smash_b200_comm_create(const char id[128], int32_t rank, int32_t world, SmashComm **comm);
void smash_b200_comm_destroy(SmashComm *comm);
/* in-place on a host vector; kind: 0 float32, 1 float64, 2 int32; op: 0 sum, 2 max, 3 min */
int smash_b200_comm_allreduce(SmashComm *comm, void *host, int64_t count, int32_t kind, int32_t op);
/* recv holds world * count values in rank order */
int smash_b200_comm_allgather(SmashComm *comm, const void *send, void *recv, int64_t count, int32_t kind);
int smash_b200_comm_rank(const SmashComm *comm);
int smash_b200_comm_world(const SmashComm *comm);
const char *smash_b200_comm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SMASH_B200_H */
