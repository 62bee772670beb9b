// split_kernels.cuh -- argument blocks and launch wrappers of the split engine (sm_100a only).
//
// Kernels:
//   vertical_forward   one thread per cell, whole time loop, reservoirs in registers, forcing tiles [4 steps][32 cells]
//                      by 2-D TMA (cp.async.bulk.tensor, SASS UTMALDG) into a per-warp mbarrier ring; writes the
//                      cell's runoff series qt as a row [cell][time]
//   route_forward      one warp per heavy-path chain, time axis across the lanes: upstream_discharge + linear_routing
//                      as a scan; turns qt rows into q rows in place
//   rows_to_domain     [cell][time] -> [time][cell] for the routed cells (qsim_domain layout)
//   route_adjoint      reverse of route_forward (LINEAR_ROUTING_B, UPSTREAM_DISCHARGE_B): w rows, lr_b, hlr_b
//   vertical_adjoint   reverse time loop per cell (GR_TRANSFER_B, GR_EXCHANGE_B, GR_PRODUCTION_B)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.cuh"
#include "route_graph.hpp"

namespace smash {

struct SplitTopo {
    int n, npad, ng, ntask, nchain;
    int nded;                    // the last nded chain tasks (longest chains) run on warps of their own in route_forward
    const int32_t *flwacc;       // [npad] (1 on padding)
    const int32_t *up_begin;     // [n + 1]
    const RouteUp *up;
    const int32_t *down;         // [npad] consumer cell or -1
    const int32_t *down_task;    // [npad]
    const int32_t *down_need;    // [npad] reverse sweep: cells the consumer's task must have finished (0xffff: all)
    const uint8_t *down_lag;     // [npad] 1: the consumer reads this cell's previous time step (late cell of a pit pair)
    const int32_t *task_begin, *task_cells;
    const int4 *tcell;           // TaskCell records, parallel to task_cells
    const int2 *tup;             // RouteUp entries of the task cells
    int nrouted;                 // cells with flwacc > 1
    const int32_t *rlist;        // [nrouted] routed cells in path order (a topological order)
    const int32_t *rindex;       // [npad] position in rlist or -1
    const int32_t *cell_task;    // [npad] task that routes the cell, -1 for lone source cells and padding
    const int32_t *gauge_first;  // [npad]
    const int32_t *gauge_next;   // [ng]
    const uint8_t *deep;         // [npad] or nullptr.  Set for the graph of the deep cells that follows the window pass: 1 = the
                                 // cell is routed by the chain scans; every other cell is final when they start (nullptr: the
                                 // cells with flwacc > 1 are routed)
};

struct SplitArgs {
    SplitTopo tp;
    int T, Tp, W, nwin;          // time steps, row pitch (= nwin * W), steps per window (= 32 * S), windows
    int t_begin, t_end;          // time range of this launch of the per-cell passes / the export (multiples of 8; whole run: 0, T)
    int tape_t0;                 // time step held by row 0 of tape_hp / tape_hft (0: whole-run tape; checkpointed runs: the window start)
    const float *qprev;          // checkpointed runs: q of every cell at the last step of the PREVIOUS window (pit pairs read it at the
                                 // first step of a window); nullptr: rows holds every window
    const float *wnext;          // checkpointed runs: rows_w of every cell at the first step of the NEXT window (the window buffers
                                 // are reused), read across a window boundary by the pit pairs; nullptr: rows_w holds every window
    int first_routed;            // smallest cell index with flwacc > 1
    int nmember;
    float dt, dx;
    int save_q, save_netp;
    int grd;                     // 1: the gr-a kernels run gr-d's statements (shares 1 / 0, no exchange; cell_math.cuh CellConst::kr, kd)
    int fuse_export;             // 1: the routing warps also write the routed cells' series to qdom ([t][cell]) once their chains are done
    unsigned long long *dbg_prof;  // diagnostics: per dedicated chain [cells, cycles, cycles waiting for tributaries, end time ns], or nullptr
    const float *fields;         // [m][NFIELD][npad]
    float *fstates;              // [m][3][npad]
    const float *sfields;        // [m][16 + 8][npad] every parameter and state plane (structures other than gr-a), or nullptr
    float *sfstates;             // [m][8][npad]      final stores of those structures (hlr: fstates)
    float *rows;                 // [m][npad][Tp]   qt of every cell after vertical_forward, q after route_forward
    float *qdom;                 // [m][T][qpitch]  domain discharge, cell order j
    float *netp;                 // [m][T][qpitch]  qt (save_net_prcp_domain)
    int64_t qpitch;
    float *qsim;                 // [m][T][ng]
    float *tape_hp, *tape_hft;   // [m][T][npad]    pre-step reservoir states (gradient runs)
    float *rows_hr;              // [m][npad][Tp]   hr_imd = hlr0 + qup of the routed cells (gradient runs)
    float *hcar;                 // [m][npad]       routing state carried across windows (starts as the hlr field)
    int *done;                   // [m][ntask]      forward: windows finished by each task
    unsigned int *ticket;        // [2] chain tickets, export-tile tickets
    // dynamic scheduling of the ticketed chains (option route_dynamic): a chain is pushed to the ready queue of its basin
    // class when its last tributary chain has finished; warps pop from the queues in priority order
    int dyn, dyn_nq;             // on / number of queues (<= 16)
    unsigned int *qctl;          // [32] heads [0..16), tails [16..32): one 128-byte line
    int *queue;                  // [nticket] ready entries of all queues, queue q owns [qoff[q], qoff[q + 1]); -1 = not pushed yet
    int *ndep;                   // [nticket] tributary chains still running
    const int *cons;             // [nchain] ticketed chain that gathers the task's last cell, or -1
    const int *qid;              // [nticket] queue of the chain
    const int *qoff;             // [dyn_nq + 1]
    const unsigned int *qctl0;   // initial images, copied before every window
    const int *queue0, *ndep0;
    // adjoint
    const float *qsim_b;         // [m][T][ng]
    float *rows_w;               // [m][npad][Tp]   s * hr_imd_b of the routed cells (UPSTREAM_DISCHARGE_B)
    float *gcar;                 // [m][npad]       hlr_b carried across windows
    float *grad;                 // [m][NFIELD][npad]
    int *rdone;                  // [m][ntask]
};

// ---- tick pass (tick_kernels.cu): reservoirs + routing of the whole domain, 8 steps at a time ------------------------------
constexpr int TK_W = 8;          // time steps per window = one TMA box = one 32-byte sector per cell

struct TkTopo {
    int n, npad, ntile, nreach, ng;
    const int32_t *meta;         // [npad] TickTopoHost::meta (route_graph.hpp)
    const int32_t *upoff;        // [npad] first inflow entry of an R or D cell
    const int32_t *ups;          // producer cells in the reference's summation order (md_routing_operator.f90:37-53); bit 30: the
                                 // block was written many ticks before it is read
    const int32_t *cons1, *cons2;  // [npad] TickTopoHost::cons1 / cons2; bit 30: the reader runs many ticks later
    const int32_t *need;         // [ntile + nreach] blocks of other units a ticket reads
    const uint8_t *tile_rounds;  // [ntile] dependency rounds inside the tile (0: no R cell)
    const int32_t *reach_cells;  // [nreach][32]
    const int32_t *gauge_first;  // [npad]
    const int32_t *gauge_next;   // [ng]
};

struct TkArgs {
    TkTopo tp;
    int T, Tp, nwin;             // time steps; pitch of the pit rows; windows = ceil(T / 8)
    int nb;                      // windows per visit: a ticket is (unit, visit), key = key of the stage + visit
    int nwarp, maxu;             // warps of the resident grid; columns of wunits
    const int32_t *wunits;       // [nwarp][maxu] (unit, ticket key of its stage), sorted by stage, -1 = none
    float dt, dx;
    int save_q, save_netp;
    float4 *cc;                  // [npad] cp, cft, exc, exp(-dt / (60 lr))
    float *fstates;              // [3][npad] reservoir and routing states carried from window to window; final states
    float *X;                    // [nwin][npad][TK_W] discharge blocks handed from producer to consumer
    float *rows;                 // [npad][Tp] runoff of the pit cells, discharge of the cells that flow into a pit cell
    float *qdom, *netp;          // [T][qpitch]
    int64_t qpitch;
    float *qsim;                 // [T][ng]
    int *cnt;                    // [visits][ntile + nreach] cells whose blocks have arrived for ticket (unit, visit)
    int *err;                    // set to 1 + unit when a wait did not end (the results are then invalid)
    unsigned long long *dbg;     // diagnostics (option tick_dbg) or nullptr: [k] time the last warp left tick k, [1024] start
};

size_t tick_smem_bytes();
int tick_max_units();
// warps of the fully resident grid (variant: 8 / 6 / 4 CTAs per SM aimed at; ctas_per_sm > 0 caps it)
cudaError_t tick_grid_warps(int variant, int ctas_per_sm, int *nwarp);
// the whole run: per-cell constants and carried states from `fields`, then every ticket
cudaError_t launch_tick_forward(const TkArgs &a, const float *fields, const CUtensorMap &prcp, const CUtensorMap &pet,
                                cudaStream_t s, int variant = 8);

// ---- subtree engine (sub_kernels.cu): one pass over the forcing, the engine owns the cell order --------------------------
constexpr int SB_W = 8;          // time steps per forcing box / exchange block

struct SbTopo {
    int ntile, ng, nslot, dmax;
    const int32_t *cell;         // [ntile * 32] SubTopoHost::cell (route_graph.hpp)
    const int32_t *rec;          // [ntile * 32]
    const uint2 *child;          // [ntile * 32]
    const int32_t *xout;         // [ntile * 32]
    const int32_t *extoff;       // [ntile * 32]
    const int32_t *extlist;
    const uint8_t *tile_kmax, *tile_ext;   // [ntile]
};

struct SbArgs {
    SbTopo tp;
    int T, Tp, nwin, npad;       // time steps; pitch of the pit rows; windows = ceil(T / 8); columns of fields / fstates (cell order j)
    float dt, dx;
    int save_q, save_netp;
    const float *fields;         // [NFIELD][npad] cell order j
    const int32_t *flwacc;       // [npad]
    const int32_t *gauge_first;  // [npad]
    const int32_t *gauge_next;   // [ng]
    float *fstates;              // [3][npad] final states, cell order j
    float *X;                    // [nslot][nwin][8] exchange blocks (NaN = not written yet)
    float *rows;                 // [npad][Tp] runoff of the pit cells, discharge of the cells that flow into a pit cell
    float *qdom, *netp;          // [T][qpitch] ENGINE order (column j' = tile * 32 + lane), qpitch = ntile * 32
    int64_t qpitch;
    float *qsim;                 // [T][ng]
    int *err;                    // set to 1 + tile when a wait did not end
    int nowait;                  // diagnostics: do not wait for the exchange blocks (wrong results; shows the dependency-free time)
};
size_t sub_smem_bytes();
cudaError_t launch_sub_forward(const SbArgs &a, const CUtensorMap &prcp, const CUtensorMap &pet, cudaStream_t s);

// 2-D tensor map over a [rows][pitch] float array, box = 8 rows x 32 columns.  cols = valid columns (the rest reads 0).
int make_tensor_map_2d(CUtensorMap *tm, const float *base, uint64_t cols, uint64_t rows, uint64_t pitch_elems, const char **err);

// structures other than gr-a (struct_kernels.cu): the reservoir pass of gr-b / gr-c / gr-d / vic-a, whole run, forward only
cudaError_t launch_vertical_struct(const SplitArgs &a, const CUtensorMap &prcp, const CUtensorMap &pet, int structure, int math_mode,
                                   cudaStream_t s);
cudaError_t launch_gather_struct_fields(const int32_t *cell, int npad, int nmember, const float *planes, int64_t ncell, const float *sample,
                                        const int32_t *sample_plane, int nvar, float *sfields, cudaStream_t s);
cudaError_t launch_vertical_forward(const SplitArgs &a, const CUtensorMap &prcp, const CUtensorMap &pet, int math_mode, bool tape,
                                    cudaStream_t s);
cudaError_t launch_route_forward(const SplitArgs &a, bool tape, cudaStream_t s);      // all windows
// one window (streamed runs: window w is routed as soon as its forcing has arrived); w = 0 resets the flags
// reset: clear the done flags first (checkpointed runs replay the windows in reverse order)
cudaError_t launch_route_forward_window(const SplitArgs &a, int w, bool tape, cudaStream_t s, bool reset = false);
// ensembles on small meshes without pit pairs: lane = member, strictly sequential arithmetic (nrouted <= 12000)
cudaError_t launch_route_members(const SplitArgs &a, bool tape, cudaStream_t s);     // all windows
cudaError_t launch_rows_to_domain(const SplitArgs &a, cudaStream_t s);
cudaError_t launch_route_adjoint(const SplitArgs &a, cudaStream_t s);                // all windows, reverse order
// one window of the reverse routing sweep (checkpointed runs); w = nwin - 1 resets the flags and the carried adjoint state
cudaError_t launch_route_adjoint_window(const SplitArgs &a, int w, cudaStream_t s);
// out[j] = rows[j * pitch] (the first step of a window buffer), j < npad
cudaError_t launch_first_step(const float *rows, int pitch, int npad, float *out, cudaStream_t s);
cudaError_t launch_vertical_adjoint(const SplitArgs &a, const CUtensorMap &prcp, const CUtensorMap &pet, const CUtensorMap &hp,
                                    const CUtensorMap &hft, int math_mode, cudaStream_t s);

// out[t * pitch + j] = raw[t * stride + idx[j]]  (j < n; padding columns get 0)
cudaError_t launch_pack_columns(const float *raw, int64_t stride, const int32_t *idx, int n, int npad, int T, float *out,
                                cudaStream_t s);
// out[t * stride + idx[j]] = src[t * pitch + j]
cudaError_t launch_scatter_columns(const float *src, int64_t pitch, const int32_t *idx, int n, int T, int64_t stride, float *out,
                                   cudaStream_t s);
// out[t * stride + didx[i]] = src[t * pitch + sidx[i]], i < n
cudaError_t launch_copy_columns(const float *src, int64_t pitch, const int32_t *sidx, const int32_t *didx, int n, int T, int64_t stride,
                                float *out, cudaStream_t s);
// sum over t < T, j < n of a [T][pitch] array (double accumulation)
cudaError_t launch_sum_domain(const float *src, int64_t pitch, int n, int T, double *out, cudaStream_t s);

int split_pick_window(int T, int *S, int *nwin, bool small = false);   // returns W = 32 * S; small: 256-step windows

}  // namespace smash
