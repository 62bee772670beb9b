// pre_kernels.cu -- device versions of the per-model preprocessing that sits in front of the solver (SURVEY.md section 8f,
// "next-4": meshing and input pipeline): flow accumulation of a D8 raster, catchment masks of the gauges, catchment means of
// the forcing.  All integer work is bit-exact whatever the thread order (integer sums commute); the means are accumulated
// in float64 and rounded once.
//
// Reference statements are cited as file:line under /root/reference/smash/.
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "../../include/smash_b200.h"

namespace smash {

namespace {

// mesh/mw_meshing.f90:163-164 and solver/routine/mw_mask.f90:28-30: the cell a direction code fd = 1..8 points to
__constant__ int FD_DROW[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
__constant__ int FD_DCOL[8] = {0, 1, 1, 1, 0, -1, -1, -1};

// downstream cell of c (flat Fortran index row + col * nrow) or -1
__device__ __forceinline__ int down_cell(const int32_t *flwdir, int nrow, int ncol, int c) {
    const int fd = flwdir[c];
    if (fd < 1 || fd > 8) return -1;
    const int row = c % nrow + FD_DROW[fd - 1], col = c / nrow + FD_DCOL[fd - 1];
    if (row < 0 || row >= nrow || col < 0 || col >= ncol) return -1;
    return row + col * nrow;
}

// fill_nipd (mw_meshing.f90:111-152): number of neighbours that point at the cell.  start = 1 where nobody does.
__global__ void nipd_kernel(int nrow, int ncol, const int32_t *flwdir, const int32_t *mask, int *nipd, int32_t *acc, uint8_t *start) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nrow * ncol) return;
    acc[c] = 1;                                                           // :212
    const bool in = !mask || mask[c] != 0;
    int n = 0;
    if (in) {
        const int row = c % nrow, col = c / nrow;
        for (int i = 0; i < 8; i++) {
            // the neighbour the code i + 1 leads AWAY from: it points here iff its own code is i + 1
            const int r2 = row - FD_DROW[i], c2 = col - FD_DCOL[i];
            if (r2 < 0 || r2 >= nrow || c2 < 0 || c2 >= ncol) continue;
            const int nb = r2 + c2 * nrow;
            if (mask && mask[nb] == 0) continue;
            if (flwdir[nb] == i + 1) n++;
        }
    }
    nipd[c] = n;
    start[c] = in && n == 0;
}

// downstream_cell_flwacc (mw_meshing.f90:154-202) without recursion: every cell nobody points at walks downstream; it adds its
// count to the receiving cell and goes on only when it was the last neighbour to arrive.  Two cells that point at each other
// (|fd - fd'| = 4) neither exchange nor release each other (:182).
__global__ void flwacc_walk_kernel(int nrow, int ncol, const int32_t *flwdir, const int32_t *mask, int *nipd, int32_t *acc,
                                   const uint8_t *start) {
    int cur = blockIdx.x * blockDim.x + threadIdx.x;
    if (cur >= nrow * ncol || !start[cur]) return;
    for (;;) {
        const int nxt = down_cell(flwdir, nrow, ncol, cur);
        if (nxt < 0 || (mask && mask[nxt] == 0)) return;
        const int d = flwdir[cur] - flwdir[nxt];
        if (d == 4 || d == -4) return;                                    // pit pair
        const int v = atomicAdd(&acc[cur], 0);                            // final: every neighbour of cur has arrived
        atomicAdd(&acc[nxt], v);
        __threadfence();
        if (atomicSub(&nipd[nxt], 1) > 1) return;                         // others still to come: the last one carries on
        cur = nxt;
    }
}

// mask_upstream_cells (mw_mask.f90:11-55) for every gauge at once: a cell belongs to gauge g iff walking downstream from it
// passes through the gauge cell.  gauge_at[c] = first gauge on cell c or -1, gauge_next chains gauges on the same cell.
__global__ void gauge_mask_kernel(int nrow, int ncol, const int32_t *flwdir, const int32_t *gauge_at, const int32_t *gauge_next,
                                  uint8_t *mask) {
    const int c0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int ncell = nrow * ncol;
    if (c0 >= ncell) return;
    int cur = c0;
    for (int steps = 0; steps <= ncell; steps++) {
        for (int g = gauge_at[cur]; g >= 0; g = gauge_next[g]) mask[(size_t)g * ncell + c0] = 1;
        const int nxt = down_cell(flwdir, nrow, ncol, cur);
        if (nxt < 0) return;
        const int d = flwdir[cur] - flwdir[nxt];
        if (d == 4 || d == -4) {                                          // pit pair: the partner is the last cell of the walk
            for (int g = gauge_at[nxt]; g >= 0; g = gauge_next[g]) mask[(size_t)g * ncell + c0] = 1;
            return;
        }
        cur = nxt;
    }
}

// compute_mean_forcing (solver/routine/mw_forcing_statistic.f90:18-75): one block per (time step, gauge); mean over the cells
// of the gauge's mask whose value is >= 0.  cells: the n cells that hold a value (sparse: the active cells in sparse order;
// dense: every cell of the rectangle), cell_of[k] = flat rectangle index.
__global__ void __launch_bounds__(256) mean_forcing_kernel(int n, int ncell, int T, int ng, const int32_t *cell_of, const uint8_t *mask,
                                                           const float *values, int64_t stride, float *mean) {
    const int t = blockIdx.x, g = blockIdx.y;
    const float *row = values + (size_t)t * stride;
    const uint8_t *mg = mask + (size_t)g * ncell;
    double s = 0.0;
    int cnt = 0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int c = cell_of ? cell_of[k] : k;
        const float v = row[k];
        if (mg[c] && v >= 0.0f) { s += (double)v; cnt++; }
    }
    __shared__ double ss[256];
    __shared__ int sc[256];
    ss[threadIdx.x] = s; sc[threadIdx.x] = cnt;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
        if (threadIdx.x < h) { ss[threadIdx.x] += ss[threadIdx.x + h]; sc[threadIdx.x] += sc[threadIdx.x + h]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) mean[(size_t)g + (size_t)t * ng] = (float)ss[0] / (float)sc[0];   // 0 / 0 = NaN as in the Fortran (:71-72)
}

template <typename T> struct Dev {
    T *p = nullptr;
    ~Dev() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(1, n) * sizeof(T)); }
};

}  // namespace

cudaError_t pre_flow_accumulation(int nrow, int ncol, const int32_t *flwdir, const int32_t *mask, int32_t *flwacc) {
    const size_t n = (size_t)nrow * ncol;
    Dev<int32_t> d_fd, d_mask, d_acc;
    Dev<int> d_nipd;
    Dev<uint8_t> d_start;
    cudaError_t e;
    if ((e = d_fd.alloc(n)) || (e = d_acc.alloc(n)) || (e = d_nipd.alloc(n)) || (e = d_start.alloc(n))) return e;
    if ((e = cudaMemcpy(d_fd.p, flwdir, n * sizeof(int32_t), cudaMemcpyHostToDevice))) return e;
    if (mask) {
        if ((e = d_mask.alloc(n))) return e;
        if ((e = cudaMemcpy(d_mask.p, mask, n * sizeof(int32_t), cudaMemcpyHostToDevice))) return e;
    }
    const int blocks = (int)((n + 255) / 256);
    nipd_kernel<<<blocks, 256>>>(nrow, ncol, d_fd.p, d_mask.p, d_nipd.p, d_acc.p, d_start.p);
    flwacc_walk_kernel<<<blocks, 256>>>(nrow, ncol, d_fd.p, d_mask.p, d_nipd.p, d_acc.p, d_start.p);
    if ((e = cudaGetLastError())) return e;
    return cudaMemcpy(flwacc, d_acc.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost);
}

// mask: (nrow, ncol, ng) bytes, Fortran order, written in full
cudaError_t pre_gauge_masks(int nrow, int ncol, int ng, const int32_t *flwdir, const int32_t *gauge_pos, uint8_t *mask, uint8_t *d_mask_out) {
    const size_t n = (size_t)nrow * ncol;
    std::vector<int32_t> at(n, -1), next(std::max(1, ng), -1);
    for (int g = ng - 1; g >= 0; g--) {
        const int row = gauge_pos[g] - 1, col = gauge_pos[g + ng] - 1;
        if (row < 0 || row >= nrow || col < 0 || col >= ncol) return cudaErrorInvalidValue;
        next[g] = at[row + (size_t)col * nrow]; at[row + (size_t)col * nrow] = g;
    }
    Dev<int32_t> d_fd, d_at, d_next;
    Dev<uint8_t> d_own;
    cudaError_t e;
    if ((e = d_fd.alloc(n)) || (e = d_at.alloc(n)) || (e = d_next.alloc(next.size()))) return e;
    uint8_t *dm = d_mask_out;
    if (!dm) { if ((e = d_own.alloc(n * std::max(1, ng)))) return e; dm = d_own.p; }
    if ((e = cudaMemcpy(d_fd.p, flwdir, n * sizeof(int32_t), cudaMemcpyHostToDevice))) return e;
    if ((e = cudaMemcpy(d_at.p, at.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice))) return e;
    if ((e = cudaMemcpy(d_next.p, next.data(), next.size() * sizeof(int32_t), cudaMemcpyHostToDevice))) return e;
    if ((e = cudaMemset(dm, 0, n * std::max(1, ng)))) return e;
    gauge_mask_kernel<<<(int)((n + 255) / 256), 256>>>(nrow, ncol, d_fd.p, d_at.p, d_next.p, dm);
    if ((e = cudaGetLastError())) return e;
    if (mask) return cudaMemcpy(mask, dm, n * std::max(1, ng), cudaMemcpyDeviceToHost);
    return cudaDeviceSynchronize();
}

// values: host array [T][stride] (sparse: stride = nac, cell_of = flat index of every sparse cell; dense: stride = nrow * ncol,
// cell_of = nullptr); mean: host (ng, T)
cudaError_t pre_mean_forcing(int nrow, int ncol, int ng, int T, const int32_t *flwdir, const int32_t *gauge_pos, int n, const int32_t *cell_of,
                             const float *prcp, const float *pet, float *mean_prcp, float *mean_pet) {
    if (ng <= 0 || T <= 0) return cudaSuccess;
    const size_t ncell = (size_t)nrow * ncol;
    Dev<uint8_t> d_mask;
    Dev<int32_t> d_cell;
    Dev<float> d_val, d_mean;
    cudaError_t e;
    if ((e = d_mask.alloc(ncell * ng))) return e;
    if ((e = pre_gauge_masks(nrow, ncol, ng, flwdir, gauge_pos, nullptr, d_mask.p))) return e;
    if (cell_of) {
        if ((e = d_cell.alloc(n))) return e;
        if ((e = cudaMemcpy(d_cell.p, cell_of, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice))) return e;
    }
    if ((e = d_val.alloc((size_t)n * T)) || (e = d_mean.alloc((size_t)ng * T))) return e;
    const float *src[2] = {prcp, pet};
    float *dst[2] = {mean_prcp, mean_pet};
    for (int a = 0; a < 2; a++) {
        if (!src[a] || !dst[a]) continue;
        if ((e = cudaMemcpy(d_val.p, src[a], (size_t)n * T * sizeof(float), cudaMemcpyHostToDevice))) return e;
        mean_forcing_kernel<<<dim3(T, ng), 256>>>(n, (int)ncell, T, ng, d_cell.p, d_mask.p, d_val.p, n, d_mean.p);
        if ((e = cudaGetLastError())) return e;
        if ((e = cudaMemcpy(dst[a], d_mean.p, (size_t)ng * T * sizeof(float), cudaMemcpyDeviceToHost))) return e;
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// adjust_interception_store (solver/routine/mw_interception_store.f90:19-160): the capacity ci of gr-b / gr-c, cell by cell.
// Thread (cell, group): one pass over the cell's forcing series forms the daily sums (:44-91) and, for the group's seven
// capacities of cmax = 0.1, 0.2 .. 4.9 (arange_r, m_array_creation.f90:41-54), the cumulated interception evaporation of
// gr_interception (operator/md_gr_operator.f90:20-34; :95-131).  |sub-daily - daily| and the candidate index are packed into
// one 64-bit key (a non-negative float orders like its bit pattern), atomicMin over the seven groups = minloc's first smallest
// (:142).  Every statement is a single rounded operation (no contraction), like the scalar Fortran.
// ------------------------------------------------------------------------------------------------
constexpr int CI_NCAND = 49, CI_GROUP = 7;

__global__ void __launch_bounds__(128) interception_kernel(int n, int T, const int64_t *src_index, int64_t slab, const float *prcp, const float *pet,
                                                           const uint8_t *newday, unsigned long long *best) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c0 = blockIdx.y * CI_GROUP;
    float ci[CI_GROUP], h[CI_GROUP], sub[CI_GROUP];
#pragma unroll
    for (int k = 0; k < CI_GROUP; k++) {
        ci[k] = __fadd_rn(0.1f, __fmul_rn((float)(c0 + k), 0.1f));       // res(i) = stt + (i - 1)*step
        h[k] = 0.0f; sub[k] = 0.0f;
    }
    const float *pp = prcp + src_index[i], *pe = pet + src_index[i];
    float dp = 0.0f, de = 0.0f, cum = 0.0f;
    for (int t = 0; t < T; t++) {
        const float p = pp[(int64_t)t * slab], e = pe[(int64_t)t * slab];
        if (newday[t]) {                                                  // :46 a new day starts: close the previous one (:79-91)
            cum = __fadd_rn(cum, fminf(dp, de));
            dp = 0.0f; de = 0.0f;
        }
        dp = __fadd_rn(dp, p); de = __fadd_rn(de, e);
#pragma unroll
        for (int k = 0; k < CI_GROUP; k++) {
            const float ei = fminf(e, __fadd_rn(p, __fmul_rn(h[k], ci[k])));
            const float pn = fmaxf(0.0f, __fsub_rn(__fsub_rn(p, __fmul_rn(ci[k], __fsub_rn(1.0f, h[k]))), ei));
            h[k] = __fadd_rn(h[k], __fdiv_rn(__fsub_rn(__fsub_rn(p, ei), pn), ci[k]));
            sub[k] = __fadd_rn(sub[k], ei);
        }
    }
    cum = __fadd_rn(cum, fminf(dp, de));
    unsigned long long key = ~0ull;
#pragma unroll
    for (int k = 0; k < CI_GROUP; k++) {
        const float d = fabsf(__fsub_rn(sub[k], cum));
        const unsigned long long kk = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(c0 + k);
        key = kk < key ? kk : key;
    }
    atomicMin(&best[i], key);
}

// n computed cells; src_index[i] = offset of cell i inside one time slab of prcp / pet (host arrays of slab * T floats);
// ci_out[i] = the chosen capacity
cudaError_t pre_interception(int n, int T, const int64_t *src_index, int64_t slab, const float *prcp, const float *pet, const int32_t *day_index,
                             float *ci_out, float *ms) {
    if (n <= 0 || T <= 0) return cudaSuccess;
    Dev<int64_t> d_idx;
    Dev<float> d_p, d_e;
    Dev<uint8_t> d_new;
    Dev<unsigned long long> d_best;
    cudaError_t e;
    std::vector<uint8_t> newday(T, 0);
    for (int t = 1; t < T; t++) newday[t] = day_index[t] != day_index[t - 1];
    if ((e = d_idx.alloc(n)) || (e = d_p.alloc((size_t)slab * T)) || (e = d_e.alloc((size_t)slab * T)) || (e = d_new.alloc(T)) ||
        (e = d_best.alloc(n)))
        return e;
    if ((e = cudaMemcpy(d_idx.p, src_index, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice))) return e;
    if ((e = cudaMemcpy(d_p.p, prcp, (size_t)slab * T * sizeof(float), cudaMemcpyHostToDevice))) return e;
    if ((e = cudaMemcpy(d_e.p, pet, (size_t)slab * T * sizeof(float), cudaMemcpyHostToDevice))) return e;
    if ((e = cudaMemcpy(d_new.p, newday.data(), (size_t)T, cudaMemcpyHostToDevice))) return e;
    if ((e = cudaMemset(d_best.p, 0xff, (size_t)n * sizeof(unsigned long long)))) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    interception_kernel<<<dim3((unsigned)((n + 127) / 128), CI_NCAND / CI_GROUP), 128>>>(n, T, d_idx.p, slab, d_p.p, d_e.p, d_new.p, d_best.p);
    cudaEventRecord(e1);
    if ((e = cudaGetLastError())) return e;
    std::vector<unsigned long long> best(n);
    if ((e = cudaMemcpy(best.data(), d_best.p, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost))) return e;
    if (ms) cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (int i = 0; i < n; i++) {
        const int c = (int)(best[i] & 0xffffffffu);
        ci_out[i] = 0.1f + (float)c * 0.1f;
    }
    return cudaSuccess;
}

}  // namespace smash
