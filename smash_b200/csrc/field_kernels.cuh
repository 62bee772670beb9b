// field_kernels.cuh -- device kernels for the O(nrow * ncol) parts of base_forward / BASE_FORWARD_B / base_hyper_forward(_b)
// (forward/forward.f90:1-157, forward/forward_db.f90:10648-10936, 11231-11554): hyper mapping + adjoint reductions,
// Jreg + adjoint, gradient planes back to the rectangle.  sm_100a only.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/smash_b200.h"
#include "kernels.cuh"

namespace smash {

constexpr int HYPER_MAX_ND = 8;       // descriptors (setup%nd); Cance has 2, Lez 6
constexpr int HYPER_NPLANE = SMASH_B200_GNP + SMASH_B200_GNS;   // 16 parameter + 8 state planes, stacked

struct HyperArgs {
    int n, npad, ncell, nd, nh, poly;  // computed cells, padded, nrow * ncol, descriptors, nhyper, 0 linear / 1 polynomial
    const int32_t *cell;               // [npad] rectangle index of cell j or -1
    const float *desc;                 // [nd][ncell] descriptors, normalised by the caller (mw_optimize.f90:960-980)
    const float *hyper;                // [24][nh] hyper-parameters of the stacked planes
    int live[NFIELD];                  // stacked plane index of the device fields cp, cft, exc, lr, hp, hft, hlr
    float lb[HYPER_NPLANE], ub[HYPER_NPLANE];
    float *fields;                     // [NFIELD][npad]
};
struct GradScale { int on; float span[NFIELD]; };   // denormalize_forward: ub - lb of the live fields

struct JregArgs {
    int nrow, ncol, ncell, nplanes;    // planes = the optimised fields only
    int normalize;                     // the planes are denormalised: normalise before use
    const int32_t *active;             // [ncell]
    const float *mat, *bgd;            // [nplanes][ncell]
    float *mat_b;                      // [nplanes][ncell] adjoint (normalised space), accumulated
    float lb[HYPER_NPLANE], ub[HYPER_NPLANE];   // per plane of the list
};

cudaError_t launch_hyper_fields(const HyperArgs &a, cudaStream_t s);
cudaError_t launch_hyper_rect(const HyperArgs &a, float *rect, cudaStream_t s);
int hyper_reduce_blocks(int n);   // rows of the partial buffer: [blocks][NFIELD][1 + 2 * HYPER_MAX_ND] doubles
cudaError_t launch_hyper_reduce(const HyperArgs &a, const float *grad, double *partial, float *hyper_b, cudaStream_t s);
cudaError_t launch_scatter_grad(const float *grad, const int32_t *cell, int n, int npad, int ncell, const GradScale &sc, float *rect,
                                cudaStream_t s);
int jreg_blocks(int ncell);       // partial buffer: [nplanes][blocks] doubles
// *jreg += weight * term; res_b != 0: mat_b += d term / d theta * res_b
cudaError_t launch_jreg_term(const JregArgs &a, int kind, float weight, float res_b, double *partial, float *jreg, cudaStream_t s);

}  // namespace smash
