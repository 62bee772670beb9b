// field_kernels.cu -- the O(nrow * ncol) halves of base_forward / BASE_FORWARD_B / base_hyper_forward(_b) on the device:
// hyper mapping and its adjoint reductions, the regularisation term Jreg and its adjoint, (de)normalisation fused into
// the kernels that need it, gradient planes scattered back to the rectangle.
//
// Reference statements are cited as file:line under /root/reference/smash/solver/.
#include "field_kernels.cuh"

#include <algorithm>

namespace smash {

namespace {

constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(FULLM, v, d);   // fixed order: deterministic
    return v;
}

// z = h(1) + sum_j a_j * d_j ** b_j  (routine/mwd_parameters_manipulation.f90:326-350; linear: a = h(j+1), b = 1;
// polynomial: a = h(2j), b = h(2j+1)).  A term with a == 0 contributes a signed zero in the reference (descriptors are
// normalised to [0,1] and b is bounded to [0.5, 2], mw_optimize.f90:960-1010, so d ** b is finite): it is skipped.
__device__ __forceinline__ float hyper_z(const float *h, const float *d, int nd, int poly) {
    float z = h[0];
    for (int j = 1; j <= nd; j++) {
        const float a = poly ? h[2 * j - 1] : h[j];
        if (a == 0.0f) continue;
        const float p = poly ? powf(d[j - 1], h[2 * j]) : d[j - 1];
        z = __fadd_rn(z, __fmul_rn(a, p));                               // no FMA contraction: the reference has none
    }
    return z;
}
__device__ __forceinline__ float sigmoid_map(float z, float lb, float ub) {
    return __fadd_rn(__fmul_rn(ub - lb, 1.0f / (1.0f + expf(-z))), lb);   // :352-356
}

// ---- hyper_parameters_to_parameters + hyper_states_to_states straight into the plan's [field][npad] planes -----------
__global__ void __launch_bounds__(256) hyper_fields_kernel(const HyperArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.npad) return;
    const int c = a.cell[j];
    float d[HYPER_MAX_ND];
    for (int k = 0; k < a.nd; k++) d[k] = (c >= 0) ? a.desc[(size_t)k * a.ncell + c] : 0.0f;
    for (int f = 0; f < NFIELD; f++) {
        const float v = (c >= 0) ? sigmoid_map(hyper_z(a.hyper + (size_t)a.live[f] * a.nh, d, a.nd, a.poly), a.lb[a.live[f]], a.ub[a.live[f]]) : 1.0f;
        a.fields[(size_t)f * a.npad + j] = v;
    }
}

// ---- the same mapping over the whole rectangle for all 16 + 8 planes: what the caller's arrays hold after the call --------
__global__ void __launch_bounds__(256) hyper_rect_kernel(const HyperArgs a, float *rect /*[24][ncell]*/) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.ncell) return;
    float d[HYPER_MAX_ND];
    for (int k = 0; k < a.nd; k++) d[k] = a.desc[(size_t)k * a.ncell + c];
    for (int i = 0; i < HYPER_NPLANE; i++)
        rect[(size_t)i * a.ncell + c] = sigmoid_map(hyper_z(a.hyper + (size_t)i * a.nh, d, a.nd, a.poly), a.lb[i], a.ub[i]);
}

// ---- HYPER_PARAMETERS_TO_PARAMETERS_B / HYPER_STATES_TO_STATES_B (forward/forward_db.f90:1434-1537, 2272-2369) ------------
// Per live field f (cp, cft, exc, lr, hp, hft, hlr; every other plane has a zero adjoint) and cell:
//   g = theta_b * (ub - lb) * e / (1 + e)^2, e = exp(-z);   h1_b = sum g;  a_j_b = sum d_j**b_j * g;
//   b_j_b = sum_{d_j > 0} d_j**b_j * ln d_j * a_j * g.
// Inactive cells have theta_b = 0, so the sums run over the computed cells.  Stage 1: per-block partial sums in double,
// fixed shuffle order; stage 2: one thread per output adds the partials in block order.  Deterministic.
__global__ void __launch_bounds__(256) hyper_reduce_kernel(const HyperArgs a, const float *grad /*[NFIELD][npad]*/, double *partial) {
    __shared__ double sh[8][1 + 2 * HYPER_MAX_ND];
    const int f = blockIdx.y;
    const int nout = 1 + 2 * a.nd;
    double acc[1 + 2 * HYPER_MAX_ND];
#pragma unroll
    for (int k = 0; k < 1 + 2 * HYPER_MAX_ND; k++) acc[k] = 0.0;
    const float *h = a.hyper + (size_t)a.live[f] * a.nh;
    const float span = a.ub[a.live[f]] - a.lb[a.live[f]];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += gridDim.x * blockDim.x) {
        const int c = a.cell[j];
        if (c < 0) continue;
        const float fb = grad[(size_t)f * a.npad + j];
        if (fb == 0.0f) continue;
        float d[HYPER_MAX_ND], p[HYPER_MAX_ND];
        float z = h[0];
        for (int k = 1; k <= a.nd; k++) {
            d[k - 1] = a.desc[(size_t)(k - 1) * a.ncell + c];
            p[k - 1] = a.poly ? powf(d[k - 1], h[2 * k]) : d[k - 1];
            const float ak = a.poly ? h[2 * k - 1] : h[k];
            if (ak != 0.0f) z = __fadd_rn(z, __fmul_rn(ak, p[k - 1]));
        }
        const float e = expf(-z);
        const float t = e + 1.0f;
        const float g = e * span * fb / (t * t);                          // :1489-1500
        acc[0] += (double)g;
        for (int k = 1; k <= a.nd; k++) {
            acc[2 * k - 1] += (double)(p[k - 1] * g);                     // a_b
            if (a.poly && !(d[k - 1] <= 0.0f)) acc[2 * k] += (double)(p[k - 1] * logf(d[k - 1]) * (h[2 * k - 1] * g));   // b_b
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < nout; k++) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) sh[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < nout) {
        double v = 0.0;
        for (int w = 0; w < 8; w++) v += sh[w][threadIdx.x];
        partial[((size_t)blockIdx.x * NFIELD + f) * (1 + 2 * HYPER_MAX_ND) + threadIdx.x] = v;
    }
}
__global__ void hyper_reduce_final_kernel(const HyperArgs a, const double *partial, int nblocks, float *hyper_b /*[NFIELD][nh]*/) {
    const int f = blockIdx.x, k = threadIdx.x;
    const int nout = 1 + 2 * a.nd;
    if (k >= nout) return;
    double v = 0.0;
    for (int b = 0; b < nblocks; b++) v += partial[((size_t)b * NFIELD + f) * (1 + 2 * HYPER_MAX_ND) + k];
    // partial index: 0 -> h(1); 2j-1 -> a_j; 2j -> b_j.  hyper layout: linear h(1+j) = a_j; polynomial h(2j) = a_j, h(2j+1) = b_j
    int dst;
    if (k == 0) dst = 0;
    else if (a.poly) dst = k;                                             // (2j-1) -> index 2j-1 (0-based of h(2j)), 2j -> 2j
    else { if ((k & 1) == 0) return; dst = (k + 1) / 2; }
    hyper_b[(size_t)f * a.nh + dst] = (float)v;
}

// ---- gradient planes [NFIELD][npad] (cell order) -> rectangle planes, optionally scaled by (ub - lb) --------------------
__global__ void __launch_bounds__(256) scatter_grad_kernel(const float *grad, const int32_t *cell, int n, int npad, int ncell,
                                                           GradScale sc, float *rect /*[NFIELD][ncell], zeroed or holding Jreg_b*/) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int c = cell[j];
    if (c < 0) return;
    for (int f = 0; f < NFIELD; f++) {
        float v = rect[(size_t)f * ncell + c];
        if (sc.on) v = v / sc.span[f];                                    // NORMALIZE_*_B forward_db.f90:809-889, 1877-1900
        v = v + grad[(size_t)f * npad + j];                               // GR_A_FORWARD_B accumulates on top (:10885)
        if (sc.on) v = sc.span[f] * v;                                    // DENORMALIZE_*_B :10931-10935
        rect[(size_t)f * ncell + c] = v;
    }
}

// ---- Jreg (optimize/mwd_cost.f90:159-245, 1100-1221) and its adjoint (forward_db.f90:5504-5800) ---------------------------
// theta(i, c): the plane as the caller holds it; when `normalize` is set it is denormalised and is normalised first
// (compute_cost mwd_cost.f90:284-298).  bgd is always in normalised space.
__device__ __forceinline__ float jreg_theta(const JregArgs &a, int i, int c) {
    const float v = a.mat[(size_t)i * a.ncell + c];
    return a.normalize ? (v - a.lb[i]) / (a.ub[i] - a.lb[i]) : v;         // mwd_parameters_manipulation.f90:154-179
}
// second differences of plane i around the active cell (row, col), 1-based, edges and inactive neighbours clamped to the
// cell itself (mwd_cost.f90:1131-1168)
struct Stencil { int min_row, max_row, min_col, max_col; float dr, dc; };
__device__ __forceinline__ float jreg_value(const JregArgs &a, int i, int row, int col, bool rel) {
    const int c = (row - 1) + (col - 1) * a.nrow;
    const float v = jreg_theta(a, i, c);
    return rel ? v - a.bgd[(size_t)i * a.ncell + c] : v;
}
__device__ __forceinline__ Stencil jreg_stencil(const JregArgs &a, int i, int row, int col, bool rel) {
    Stencil s;
    s.min_col = max(1, col - 1); s.max_col = min(a.ncol, col + 1);
    s.min_row = max(1, row - 1); s.max_row = min(a.nrow, row + 1);
    auto act = [&](int r, int c) { return a.active[(r - 1) + (c - 1) * a.nrow]; };
    if (act(row, s.min_col) == 0) s.min_col = col;
    if (act(row, s.max_col) == 0) s.max_col = col;
    if (act(s.min_row, col) == 0) s.min_row = row;
    if (act(s.max_row, col) == 0) s.max_row = row;
    const float m0 = jreg_value(a, i, row, col, rel);
    s.dr = jreg_value(a, i, s.max_row, col, rel) - 2.0f * m0 + jreg_value(a, i, s.min_row, col, rel);
    s.dc = jreg_value(a, i, row, s.max_col, rel) - 2.0f * m0 + jreg_value(a, i, row, s.min_col, rel);
    return s;
}

// one term of Jreg (prior / smoothing / hard_smoothing) over the optimised planes: partial sums per block (double);
// with res_b != 0 the adjoint is accumulated into mat_b in gather form (each cell collects what the stencils around it
// would scatter), so no atomics and a fixed summation order.
__global__ void __launch_bounds__(256) jreg_term_kernel(const JregArgs a, int kind, float res_b, double *partial) {
    __shared__ double sh[8];
    const int i = blockIdx.y;
    double acc = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.ncell; c += gridDim.x * blockDim.x) {
        if (kind == SMASH_JREG_PRIOR) {                                   // reg_prior mwd_cost.f90:1180-1221: whole rectangle
            const float d = jreg_theta(a, i, c) - a.bgd[(size_t)i * a.ncell + c];
            acc += (double)(d * d);
            if (res_b != 0.0f) a.mat_b[(size_t)i * a.ncell + c] += 2.0f * d * res_b;          // REG_PRIOR_B :5756-5799
        } else {                                                          // reg_smoothing :1100-1178: active cells
            if (a.active[c] != 1) continue;
            const bool rel = kind == SMASH_JREG_SMOOTHING;
            const int row = c % a.nrow + 1, col = c / a.nrow + 1;
            const Stencil s = jreg_stencil(a, i, row, col, rel);
            acc += (double)(s.dr * s.dr + s.dc * s.dc);
            if (res_b != 0.0f) {                                          // REG_SMOOTHING_B :5504-5657, gathered
                float g = 0.0f;
                // this cell's own stencil
                g += 2.0f * s.dr * res_b * (float)((s.max_row == row) + (s.min_row == row) - 2);
                g += 2.0f * s.dc * res_b * (float)((s.max_col == col) + (s.min_col == col) - 2);
                // stencils of the four neighbours that reach this cell
                if (row > 1 && a.active[c - 1] == 1) { const Stencil p = jreg_stencil(a, i, row - 1, col, rel); if (p.max_row == row) g += 2.0f * p.dr * res_b; }
                if (row < a.nrow && a.active[c + 1] == 1) { const Stencil p = jreg_stencil(a, i, row + 1, col, rel); if (p.min_row == row) g += 2.0f * p.dr * res_b; }
                if (col > 1 && a.active[c - a.nrow] == 1) { const Stencil p = jreg_stencil(a, i, row, col - 1, rel); if (p.max_col == col) g += 2.0f * p.dc * res_b; }
                if (col < a.ncol && a.active[c + a.nrow] == 1) { const Stencil p = jreg_stencil(a, i, row, col + 1, rel); if (p.min_col == col) g += 2.0f * p.dc * res_b; }
                a.mat_b[(size_t)i * a.ncell + c] += g;
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double v = warp_sum(acc);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sh[w];
        partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}
__global__ void jreg_final_kernel(const double *partial, int count, float weight, float *jreg /*accumulated*/) {
    double v = 0.0;
    for (int k = 0; k < count; k++) v += partial[k];
    *jreg = *jreg + weight * (float)v;                                    // compute_jreg :206-226
}

}  // namespace

cudaError_t launch_hyper_fields(const HyperArgs &a, cudaStream_t s) {
    hyper_fields_kernel<<<(a.npad + 255) / 256, 256, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_hyper_rect(const HyperArgs &a, float *rect, cudaStream_t s) {
    hyper_rect_kernel<<<(a.ncell + 255) / 256, 256, 0, s>>>(a, rect);
    return cudaGetLastError();
}
int hyper_reduce_blocks(int n) { return std::max(1, std::min(296, (n + 255) / 256)); }
cudaError_t launch_hyper_reduce(const HyperArgs &a, const float *grad, double *partial, float *hyper_b, cudaStream_t s) {
    const int nb = hyper_reduce_blocks(a.n);
    cudaError_t e = cudaMemsetAsync(hyper_b, 0, sizeof(float) * (size_t)NFIELD * a.nh, s);
    if (e != cudaSuccess) return e;
    hyper_reduce_kernel<<<dim3(nb, NFIELD), 256, 0, s>>>(a, grad, partial);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    hyper_reduce_final_kernel<<<NFIELD, 32, 0, s>>>(a, partial, nb, hyper_b);
    return cudaGetLastError();
}
cudaError_t launch_scatter_grad(const float *grad, const int32_t *cell, int n, int npad, int ncell, const GradScale &sc, float *rect,
                                cudaStream_t s) {
    scatter_grad_kernel<<<(n + 255) / 256, 256, 0, s>>>(grad, cell, n, npad, ncell, sc, rect);
    return cudaGetLastError();
}
int jreg_blocks(int ncell) { return std::max(1, std::min(296, (ncell + 255) / 256)); }
cudaError_t launch_jreg_term(const JregArgs &a, int kind, float weight, float res_b, double *partial, float *jreg, cudaStream_t s) {
    if (a.nplanes <= 0) return cudaSuccess;
    const int nb = jreg_blocks(a.ncell);
    jreg_term_kernel<<<dim3(nb, a.nplanes), 256, 0, s>>>(a, kind, res_b, partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    jreg_final_kernel<<<1, 1, 0, s>>>(partial, nb * a.nplanes, weight, jreg);
    return cudaGetLastError();
}

}  // namespace smash
