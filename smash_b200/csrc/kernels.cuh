// kernels.cuh -- device-side argument blocks and launch wrappers (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "topology.hpp"

namespace smash {

constexpr int NFIELD = 7;  // device field order: cp, cft, exc, lr, hp, hft, hlr
enum { F_CP = 0, F_CFT, F_EXC, F_LR, F_HP, F_HFT, F_HLR };
constexpr int RING_STAGES = 8;  // forcing / tape rows in flight per CTA (TMA bulk copies)

struct DeviceTopology {
    int T, B, nblocks, nslots, ng;
    int64_t total_ticks;
    const int32_t *cell, *off, *flwacc;
    const uint8_t *late, *early;
    const int32_t *up_begin;
    const UpEntry *up;
    const ExtRef *ext, *rext;
    const int32_t *down_kind, *down_lane;
    const int32_t *gauge_first, *gauge_next;
    const int32_t *hmax;
    const int64_t *tick_base;
    const uint8_t *bflags;
};

struct SolverArgs {
    DeviceTopology tp;
    int nmember;
    float dt, dx;
    int save_q, save_netp, tape_on;
    int debug_nowait;       // timing experiments only: skip the cross-block progress waits (results are then wrong)
    const float *forcing;   // [tick_base[b] + d][2][B]   (shared by all members)
    const float *fields;    // [m][NFIELD][nslots]
    float *fstates;         // [m][3][nslots]
    float *qsim;            // [m][T][ng]
    float *qdom;            // [m][total_ticks][B]   skewed: row = tick_base[b] + d, time step t = d - off
    float *netp;            // same layout, qt (save_net_prcp_domain)
    float *tape;            // [m][total_ticks][4][B]  hp0, hft0, hlr0, qup
    int *prog;              // [m][nblocks] forward progress (ticks published)
    unsigned int *ticket;   // dynamic CTA numbering (deadlock-free look-back)
    // reverse sweep
    const float *qsim_b;    // [m][T][ng]
    float *wdom;            // [m][total_ticks][B]  s * qup_b published for cross-block upstream cells
    float *grad;            // [m][NFIELD][nslots]  cp_b, cft_b, exc_b, lr_b, hp_b, hft_b, hlr_b
    int *rprog;             // [m][nblocks] reverse progress (lowest tick published)
};

struct CostArgs {
    int T, ng, nmember, start;          // start = optimize_start_step - 1
    float dt, dx;
    const float *qsim;                  // [m][T][ng]
    const float *qobs;                  // [T][ng]  (= F(ng,T))
    const float *area, *wgauge;         // [ng]
    const int32_t *gauge_flwacc;        // [ng] flwacc at the gauge cell
    int njf;
    int jobs_fun[32];
    float wjobs_fun[32];
    float jobs_b;                       // adjoint seed (cost_b); qsim_b written iff qsim_b != nullptr
    float *cost_jobs;                   // [m]
    float *qsim_b;                      // [m][T][ng] or nullptr
    // signature objectives (mwd_cost.f90:770-970), forward only
    const float *mean_prcp;             // [T][ng] (= F(ng,T)) or nullptr
    const int32_t *mask_event;          // [T][ng] or nullptr
    float *scratch;                     // [m][ng][2][T] sort space of the flow percentiles, or nullptr
};

// math_mode: 0 = IEEE division / sqrt + libm tanhf ; 1 = reciprocal / rsqrt approximations
cudaError_t launch_forward(const SolverArgs &a, int math_mode, cudaStream_t s);
cudaError_t launch_reverse(const SolverArgs &a, int math_mode, cudaStream_t s);
cudaError_t launch_cost(const CostArgs &a, cudaStream_t s);

// forcing re-layout: raw[t*stride + src[slot]] -> forcing[(tick_base[b] + d)*2*B + {0,B} + lane]
cudaError_t launch_relayout_forcing(const DeviceTopology &tp, const int32_t *src_index, const float *prcp_raw,
                                    const float *pet_raw, int64_t raw_stride, float *forcing, cudaStream_t s);
// fields[m][f][slot] = plane_f[cell[slot]] overridden by uniform sample values
cudaError_t launch_gather_fields(const DeviceTopology &tp, int nmember, const float *planes /*[NFIELD][ncell]*/,
                                 int64_t ncell, const float *sample /*[m][nvar] or null*/, const int32_t *sample_field,
                                 int nvar, float *fields, cudaStream_t s);
// un-skew: out[t*out_stride + dst[slot]] = skewed[(tick_base[b] + t + off)*B + lane]
cudaError_t launch_unskew(const DeviceTopology &tp, const int32_t *dst_index, const float *skewed, int64_t out_stride,
                          float fill, float *out, cudaStream_t s);
// sum over active cell-steps of a skewed array (double accumulation)
cudaError_t launch_checksum(const DeviceTopology &tp, const float *skewed, double *out, cudaStream_t s);

}  // namespace smash
