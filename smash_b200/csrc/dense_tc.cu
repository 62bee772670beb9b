// dense_tc.cu -- the one dense contraction on the path: the Dense layers of the
// descriptor-to-parameter network of Model.ann_optimize (smash/core/net.py:579-688, x.dot(weight) + bias) on the 5th-generation
// tensor cores, for domain-sized inputs (France: 906 044 cells x 1 554 neurons).
//
// Y[M][N] = act(X[M][K] . W[K][N] + b[N]) as a TF32 tcgen05 GEMM with float32 accumulation in tensor memory: operands staged in
// shared memory by TMA, `tcgen05.mma.kind::tf32` issued by one thread per CTA, accumulators read back with `tcgen05.ld` for the
// fused epilogue (bias + activation), results stored by TMA (SASS: UTMALDG / UTCHMMA / LDTM / UTMASTG).  The kernel is
// instantiated from the CUTLASS 4.x sm100 collective builders (header tree vendored under site-packages/flashinfer); the
// whole network runs on the device: only the descriptors go up and the predicted fields come back.
//
// Numerics: TF32 inputs (10-bit mantissa) against the reference's float64 NumPy -- 1e-3 of the output's scale (stated in
// tests/test_gpu_dense.py); the NumPy path stays the default for catchment-sized networks, where the golden values live.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/epilogue/fusion/operations.hpp"
#include "cutlass/epilogue/thread/activation.h"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/dispatch_policy.hpp"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/util/packed_stride.hpp"

namespace smash {

namespace {

using namespace cute;

// X row-major [M][K] (K contiguous), W as its transpose Wt[N][K] (K contiguous: "column-major B"), Y row-major [M][N]
template <template <class> class Act, class Scheduler = void>
struct DenseGemm {
    using Element = float;
    using LayoutA = cutlass::layout::RowMajor;
    using LayoutB = cutlass::layout::ColumnMajor;
    using LayoutC = cutlass::layout::RowMajor;
    static constexpr int Align = 4;                                       // 16 bytes: what TMA needs of every leading dimension
    using Arch = cutlass::arch::Sm100;
    using OpClass = cutlass::arch::OpClassTensorOp;
#ifndef DENSE_TILE_M
#define DENSE_TILE_M _256
#define DENSE_TILE_N _128
#define DENSE_TILE_K _32
#define DENSE_CLUSTER_M _2
#endif
    using TileShape = Shape<DENSE_TILE_M, DENSE_TILE_N, DENSE_TILE_K>;    // 256 x 128: one tcgen05.mma.cta_group::2 tile per CTA pair
    using ClusterShape = Shape<DENSE_CLUSTER_M, _1, _1>;
    using Fusion = cutlass::epilogue::fusion::LinCombPerColBiasEltAct<Act, Element, float, float>;
    using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
        Arch, OpClass, TileShape, ClusterShape, cutlass::epilogue::collective::EpilogueTileAuto, float, float, Element, LayoutC, Align,
        Element, LayoutC, Align, cutlass::epilogue::collective::EpilogueScheduleAuto, Fusion>::CollectiveOp;
    using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
        Arch, OpClass, Element, LayoutA, Align, Element, LayoutB, Align, float, TileShape, ClusterShape,
        cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
        cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;
    // Scheduler = cutlass::gemm::StreamKScheduler: the contraction dimension is split over the CTAs (grad_weight: a small result
    // contracted over all the rows)
    using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, Scheduler>;
    using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;

    static const char *run(int M, int N, int K, const float *X, const float *Wt, const float *bias, float *Y, cudaStream_t s) {
        using StrideA = typename Gemm::GemmKernel::StrideA;
        using StrideB = typename Gemm::GemmKernel::StrideB;
        using StrideC = typename Gemm::GemmKernel::StrideC;
        using StrideD = typename Gemm::GemmKernel::StrideD;
        const StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, cute::make_shape(M, K, 1));
        const StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, cute::make_shape(N, K, 1));
        const StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, cute::make_shape(M, N, 1));
        const StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, cute::make_shape(M, N, 1));
        typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, 1}, {X, sa, Wt, sb}, {{}, Y, sc, Y, sd}};
        args.epilogue.thread.alpha = 1.0f;
        args.epilogue.thread.beta = 0.0f;
        args.epilogue.thread.bias_ptr = bias;
        Gemm gemm;
        if (gemm.can_implement(args) != cutlass::Status::kSuccess) return "dense layer: shape not supported by the tensor-core kernel";
        const size_t ws = Gemm::get_workspace_size(args);
        void *wsp = nullptr;
        if (ws && cudaMalloc(&wsp, ws) != cudaSuccess) return "dense layer: no memory for the kernel's workspace";
        cutlass::Status st = gemm.initialize(args, wsp, s);
        if (st == cutlass::Status::kSuccess) st = gemm.run(s);
        if (wsp) { cudaStreamSynchronize(s); cudaFree(wsp); }
        return st == cutlass::Status::kSuccess ? nullptr : "dense layer: launch of the tensor-core kernel failed";
    }
};

// activations without a fused epilogue (net.py:778-828): 4 leaky_relu (0.2), 5 elu (0.1), 6 selu, 7 softplus
__global__ void activation_kernel(float *y, size_t n, int act) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = y[i];
    float r = x;
    if (act == 4) r = x >= 0.0f ? x : 0.2f * x;
    else if (act == 5) r = x >= 0.0f ? x : 0.1f * (expf(x) - 1.0f);
    else if (act == 6) r = 1.0507009873554805f * (x >= 0.0f ? x : 1.6732632423543772f * (expf(x) - 1.0f));
    else if (act == 7) r = log1pf(expf(x));
    y[i] = r;
}

// g *= act'(z) with the derivative written in terms of the layer's OUTPUT y = act(z) (Activation._backward_pass, net.py:484-486;
// the derivatives of net.py:778-828): relu 1[y > 0], sigmoid y (1 - y), tanh 1 - y^2, leaky_relu 1 / 0.2 by the sign of y,
// elu y + alpha below 0, selu y + scale alpha below 0, softplus 1 - exp(-y)
__global__ void activation_grad_kernel(float *g, const float *y, size_t n, int act) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = y[i];
    float d = 1.0f;
    if (act == 1) d = v > 0.0f ? 1.0f : 0.0f;
    else if (act == 2) d = v * (1.0f - v);
    else if (act == 3) d = 1.0f - v * v;
    else if (act == 4) d = v >= 0.0f ? 1.0f : 0.2f;
    else if (act == 5) d = v >= 0.0f ? 1.0f : v + 0.1f;
    else if (act == 6) d = v >= 0.0f ? 1.0507009873554805f : v + 1.0507009873554805f * 1.6732632423543772f;
    else if (act == 7) d = 1.0f - expf(-v);
    g[i] *= d;
}

__global__ void pad_rows_kernel(const float *src, int64_t rows, int cols, int pitch_src, float *dst, int pitch_dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * pitch_dst) return;
    const int64_t r = i / pitch_dst;
    const int c = (int)(i - r * pitch_dst);
    dst[i] = c < cols ? src[r * pitch_src + c] : 0.0f;
}

// dst[c][r] = src[r][c]: [rows][cols] -> [cols][pitch_dst] (pitch_dst >= rows; the columns beyond `rows` are zeroed by the caller)
__global__ void transpose_kernel(const float *src, int64_t rows, int cols, float *dst, int64_t pitch_dst) {
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = r0 + i;
        const int c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * cols + c] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const int64_t r = r0 + threadIdx.x;
        if (c < cols && r < pitch_dst) dst[(size_t)c * pitch_dst + r] = tile[threadIdx.x][i];
    }
}

// column sums of g [rows][cols] in two deterministic stages: partial[chunk][col] over 1024-row chunks, then float64 over the chunks
__global__ void colsum_partial_kernel(const float *g, int64_t rows, int cols, float *partial) {
    const int c = blockIdx.y * 32 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * 1024;
    __shared__ float acc[8][33];
    float s = 0.0f;
    if (c < cols)
        for (int64_t r = r0 + threadIdx.y; r < r0 + 1024 && r < rows; r += 8) s += g[r * cols + c];
    acc[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float t = 0.0f;
        for (int i = 0; i < 8; i++) t += acc[i][threadIdx.x];
        partial[(size_t)blockIdx.x * cols + c] = t;
    }
}
__global__ void colsum_final_kernel(const float *partial, int nchunk, int cols, float *out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double t = 0.0;
    for (int k = 0; k < nchunk; k++) t += (double)partial[(size_t)k * cols + c];
    out[c] = (float)t;
}

struct DevBuf {
    float *p = nullptr;
    size_t n = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    bool alloc(size_t want) {
        if (want <= n && p) return true;
        if (p) { cudaFree(p); p = nullptr; n = 0; }
        if (cudaMalloc(&p, (want ? want : 1) * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return false; }
        n = want;
        return true;
    }
};

inline int ceil4(int v) { return (v + 3) / 4 * 4; }

template <template <class> class Act>
const char *gemm(int M, int N, int K, const float *X, const float *Wt, const float *bias, float *Y, cudaStream_t s) {
    return DenseGemm<Act>::run(M, N, K, X, Wt, bias, Y, s);
}

}  // namespace

// A chain of Dense (+ activation) layers resident on the device: padded activations of the last forward pass, weights in both
// majors, scratch for the backward pass.  act[l]: 0 none, 1 relu, 2 sigmoid, 3 tanh, 4 leaky_relu, 5 elu, 6 selu, 7 softplus.
struct Mlp {
    int64_t M = 0;
    int nlayer = 0;
    std::vector<int> sizes, act;
    std::vector<DevBuf> a, wt, wp, bias;   // a[l]: [M][ceil4(sizes[l])]; wt[l]: W^T [Np][Kp]; wp[l]: W [Kp][Np]
    DevBuf g0, g1, at, gt, part, small;    // gradient ping-pong [M][Np], transposes [Kp][Mp] / [Np][Mp], column-sum partials
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool have_x = false, have_forward = false;
    ~Mlp() {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};

const char *mlp_create(int64_t M, int nlayer, const int32_t *sizes, const int32_t *act, Mlp **out) {
    if (M <= 0 || M > 0x7fffffff || nlayer < 1) return "mlp: bad shape";
    for (int l = 0; l <= nlayer; l++) if (sizes[l] < 1) return "mlp: bad layer size";
    for (int l = 0; l < nlayer; l++) if (act[l] < 0 || act[l] > 7) return "mlp: unknown activation code";
    Mlp *m = new Mlp();
    m->M = M; m->nlayer = nlayer;
    m->sizes.assign(sizes, sizes + nlayer + 1);
    m->act.assign(act, act + nlayer);
    m->a.resize(nlayer + 1); m->wt.resize(nlayer); m->wp.resize(nlayer); m->bias.resize(nlayer);
    if (cudaEventCreate(&m->e0) != cudaSuccess || cudaEventCreate(&m->e1) != cudaSuccess) { delete m; return "mlp: cudaEventCreate failed"; }
    for (int l = 0; l <= nlayer; l++)
        if (!m->a[l].alloc((size_t)M * ceil4(sizes[l]))) { delete m; return "mlp: out of device memory"; }
    *out = m;
    return nullptr;
}
void mlp_destroy(Mlp *m) { delete m; }

static const char *mlp_upload_weights(Mlp &m, const float *const *W, const float *const *b) {
    for (int l = 0; l < m.nlayer; l++) {
        const int K = m.sizes[l], N = m.sizes[l + 1], Kp = ceil4(K), Np = ceil4(N);
        std::vector<float> wt((size_t)Np * Kp, 0.0f), wp((size_t)Kp * Np, 0.0f), bb(Np, 0.0f);
        for (int k = 0; k < K; k++)
            for (int n = 0; n < N; n++) {
                const float v = W[l][(size_t)k * N + n];
                wt[(size_t)n * Kp + k] = v;
                wp[(size_t)k * Np + n] = v;
            }
        for (int n = 0; n < N; n++) bb[n] = b[l][n];
        if (!m.wt[l].alloc(wt.size()) || !m.wp[l].alloc(wp.size()) || !m.bias[l].alloc(Np)) return "mlp: out of device memory";
        if (cudaMemcpy(m.wt[l].p, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(m.wp[l].p, wp.data(), wp.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(m.bias[l].p, bb.data(), Np * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
            return "mlp: upload failed";
    }
    return nullptr;
}

// x: host [M][sizes[0]] or NULL (the rows uploaded by an earlier call); W[l]: host [sizes[l]][sizes[l+1]] row-major (net.py's
// weight); b[l]: host [sizes[l+1]]; y: host [M][sizes[nlayer]] or NULL.  ms: device time of the layers, flops: multiply-adds x 2.
const char *mlp_forward(Mlp &m, const float *x, const float *const *W, const float *const *b, float *y, float *ms, double *flops) {
    cudaStream_t s = nullptr;
    const int64_t M = m.M;
    if (x) {
        const int K = m.sizes[0], Kp = ceil4(K);
        if (!m.g0.alloc((size_t)M * K)) return "mlp: out of device memory";
        if (cudaMemcpy(m.g0.p, x, (size_t)M * K * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return "mlp: upload failed";
        pad_rows_kernel<<<(unsigned)(((size_t)M * Kp + 255) / 256), 256>>>(m.g0.p, M, K, K, m.a[0].p, Kp);
        if (cudaDeviceSynchronize() != cudaSuccess) return "mlp: padding kernel failed";
        m.have_x = true;
    }
    if (!m.have_x) return "mlp: no input rows on the device";
    const char *err = mlp_upload_weights(m, W, b);
    if (err) return err;
    double fl = 0.0;
    cudaEventRecord(m.e0, s);
    for (int l = 0; l < m.nlayer && !err; l++) {
        const int Kp = ceil4(m.sizes[l]), Np = ceil4(m.sizes[l + 1]);
        const int a = m.act[l];
        const float *X = m.a[l].p;
        float *Y = m.a[l + 1].p;
        if (a == 1) err = gemm<cutlass::epilogue::thread::ReLu>((int)M, Np, Kp, X, m.wt[l].p, m.bias[l].p, Y, s);
        else if (a == 2) err = gemm<cutlass::epilogue::thread::Sigmoid>((int)M, Np, Kp, X, m.wt[l].p, m.bias[l].p, Y, s);
        else if (a == 3) err = gemm<cutlass::epilogue::thread::Tanh>((int)M, Np, Kp, X, m.wt[l].p, m.bias[l].p, Y, s);
        else err = gemm<cutlass::epilogue::thread::Identity>((int)M, Np, Kp, X, m.wt[l].p, m.bias[l].p, Y, s);
        if (!err && a >= 4) activation_kernel<<<(unsigned)(((size_t)M * Np + 255) / 256), 256, 0, s>>>(Y, (size_t)M * Np, a);
        fl += 2.0 * (double)M * m.sizes[l] * m.sizes[l + 1];
    }
    if (err) return err;
    cudaEventRecord(m.e1, s);
    if (cudaEventSynchronize(m.e1) != cudaSuccess) { cudaGetLastError(); return "mlp: a layer kernel failed"; }
    if (ms) cudaEventElapsedTime(ms, m.e0, m.e1);
    if (flops) *flops = fl;
    m.have_forward = true;
    if (y) {
        const int N = m.sizes[m.nlayer], Np = ceil4(N);
        if (cudaMemcpy2D(y, (size_t)N * sizeof(float), m.a[m.nlayer].p, (size_t)Np * sizeof(float), (size_t)N * sizeof(float), (size_t)M,
                         cudaMemcpyDeviceToHost) != cudaSuccess)
            return "mlp: download failed";
    }
    return nullptr;
}

// Backward pass of the chain (Dense._backward_pass net.py:672-685, Activation._backward_pass :484-486) from gy = d loss / d y
// (host [M][sizes[nlayer]]), with the activations and weights of the last forward pass:
//   g <- g * act'(.);  grad_b = column sums of g;  grad_W = a^T g;  g <- g W^T
// Both contractions run on the same tensor-core kernel: g W^T reads W as stored ([K][N], N contiguous); a^T g is formed from
// the transposes a^T [K][M] and g^T [N][M] (M contiguous), a 1 554 x 777 result contracted over the 906 044 rows.
// gW[l]: host [sizes[l]][sizes[l+1]]; gb[l]: host [sizes[l+1]].
const char *mlp_backward(Mlp &m, const float *gy, float *const *gW, float *const *gb, float *ms) {
    if (!m.have_forward) return "mlp: backward pass without a forward pass";
    cudaStream_t s = nullptr;
    const int64_t M = m.M;
    const int64_t Mp = (M + 3) / 4 * 4;
    int maxp = 0;
    for (int l = 0; l <= m.nlayer; l++) maxp = std::max(maxp, ceil4(m.sizes[l]));
    if (!m.g0.alloc((size_t)M * maxp) || !m.g1.alloc((size_t)M * maxp) || !m.at.alloc((size_t)maxp * Mp) || !m.gt.alloc((size_t)maxp * Mp) ||
        !m.part.alloc((size_t)((M + 1023) / 1024) * maxp) || !m.small.alloc((size_t)maxp * maxp))
        return "mlp: out of device memory";
    {
        const int N = m.sizes[m.nlayer], Np = ceil4(N);
        if (!m.gt.alloc(std::max((size_t)maxp * Mp, (size_t)M * N))) return "mlp: out of device memory";
        if (cudaMemcpy(m.gt.p, gy, (size_t)M * N * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return "mlp: upload failed";
        pad_rows_kernel<<<(unsigned)(((size_t)M * Np + 255) / 256), 256>>>(m.gt.p, M, N, N, m.g0.p, Np);
    }
    float *g = m.g0.p, *gn = m.g1.p;
    const char *err = nullptr;
    cudaEventRecord(m.e0, s);
    for (int l = m.nlayer - 1; l >= 0 && !err; l--) {
        const int K = m.sizes[l], N = m.sizes[l + 1], Kp = ceil4(K), Np = ceil4(N);
        if (m.act[l] != 0)
            activation_grad_kernel<<<(unsigned)(((size_t)M * Np + 255) / 256), 256, 0, s>>>(g, m.a[l + 1].p, (size_t)M * Np, m.act[l]);
        // grad_b
        const int nchunk = (int)((M + 1023) / 1024);
        colsum_partial_kernel<<<dim3(nchunk, (Np + 31) / 32), dim3(32, 8), 0, s>>>(g, M, Np, m.part.p);
        colsum_final_kernel<<<(Np + 127) / 128, 128, 0, s>>>(m.part.p, nchunk, Np, m.small.p);
        if (cudaMemcpyAsync(gb[l], m.small.p, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost, s) != cudaSuccess) { err = "mlp: download failed"; break; }
        // grad_W = a^T g
        // (the transposes write every element of [Kp][Mp] / [Np][Mp], zeros beyond the M rows)
        transpose_kernel<<<dim3((unsigned)((M + 31) / 32), (Kp + 31) / 32), dim3(32, 8), 0, s>>>(m.a[l].p, M, Kp, m.at.p, Mp);
        transpose_kernel<<<dim3((unsigned)((M + 31) / 32), (Np + 31) / 32), dim3(32, 8), 0, s>>>(g, M, Np, m.gt.p, Mp);
        err = DenseGemm<cutlass::epilogue::thread::Identity, cutlass::gemm::StreamKScheduler>::run(Kp, Np, (int)Mp, m.at.p, m.gt.p, nullptr,
                                                                                                     m.small.p, s);
        if (err) break;
        if (cudaMemcpy2DAsync(gW[l], (size_t)N * sizeof(float), m.small.p, (size_t)Np * sizeof(float), (size_t)N * sizeof(float), (size_t)K,
                              cudaMemcpyDeviceToHost, s) != cudaSuccess) { err = "mlp: download failed"; break; }
        if (cudaStreamSynchronize(s) != cudaSuccess) { err = "mlp: a backward kernel failed"; break; }   // `small` is reused by the next layer
        // g <- g W^T  ([M][Np] . [Kp][Np]^T)
        if (l > 0) {
            err = gemm<cutlass::epilogue::thread::Identity>((int)M, Kp, Np, g, m.wp[l].p, nullptr, gn, s);
            std::swap(g, gn);
        }
    }
    if (err) { cudaGetLastError(); return err; }
    cudaEventRecord(m.e1, s);
    if (cudaEventSynchronize(m.e1) != cudaSuccess) { cudaGetLastError(); return "mlp: a backward kernel failed"; }
    if (ms) cudaEventElapsedTime(ms, m.e0, m.e1);
    return nullptr;
}

// one-shot forward pass (no context kept)
const char *mlp_forward_device(int64_t M, int nlayer, const int32_t *sizes, const float *x, const float *const *W, const float *const *b,
                               const int32_t *act, float *y, float *ms, double *flops) {
    Mlp *m = nullptr;
    const char *err = mlp_create(M, nlayer, sizes, act, &m);
    if (err) return err;
    err = mlp_forward(*m, x, W, b, y, ms, flops);
    mlp_destroy(m);
    return err;
}

}  // namespace smash
