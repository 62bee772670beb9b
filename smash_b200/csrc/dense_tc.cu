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
template <template <class> class Act>
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
    using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
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

__global__ void pad_rows_kernel(const float *src, int64_t rows, int cols, int pitch_src, float *dst, int pitch_dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * pitch_dst) return;
    const int64_t r = i / pitch_dst;
    const int c = (int)(i - r * pitch_dst);
    dst[i] = c < cols ? src[r * pitch_src + c] : 0.0f;
}

struct DevBuf {
    float *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    bool alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(float)) == cudaSuccess; }
};

inline int ceil4(int v) { return (v + 3) / 4 * 4; }

}  // namespace

// The network's Dense (+ activation) layers, chained on the device.  x: host [M][sizes[0]]; W[l]: host [sizes[l]][sizes[l+1]]
// row-major (net.py's weight); b[l]: host [sizes[l+1]]; act[l]: 0 none, 1 relu, 2 sigmoid, 3 tanh, 4 leaky_relu, 5 elu, 6 selu,
// 7 softplus; y: host [M][sizes[nlayer]].  ms: device time of the layers (CUDA events), flops: multiply-adds x 2.
// Returns nullptr or an error text.
const char *mlp_forward_device(int64_t M, int nlayer, const int32_t *sizes, const float *x, const float *const *W, const float *const *b,
                               const int32_t *act, float *y, float *ms, double *flops) {
    if (M <= 0 || M > 0x7fffffff || nlayer < 1) return "mlp: bad shape";
    for (int l = 0; l <= nlayer; l++) if (sizes[l] < 1) return "mlp: bad layer size";
    for (int l = 0; l < nlayer; l++) if (act[l] < 0 || act[l] > 7) return "mlp: unknown activation code";
    cudaStream_t s = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return "mlp: cudaEventCreate failed";
    const char *err = nullptr;
    std::vector<DevBuf> acts(nlayer + 1), wts(nlayer), bias(nlayer);
    // ---- inputs and weights up; every leading dimension padded to a multiple of 4 floats with zeros
    {
        const int K = sizes[0], Kp = ceil4(K);
        DevBuf raw;
        if (!raw.alloc((size_t)M * K) || !acts[0].alloc((size_t)M * Kp)) err = "mlp: out of device memory";
        if (!err && cudaMemcpy(raw.p, x, (size_t)M * K * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) err = "mlp: upload failed";
        if (!err) pad_rows_kernel<<<(unsigned)(((size_t)M * Kp + 255) / 256), 256>>>(raw.p, M, K, K, acts[0].p, Kp);
        if (!err && cudaDeviceSynchronize() != cudaSuccess) err = "mlp: padding kernel failed";
    }
    for (int l = 0; l < nlayer && !err; l++) {
        const int K = sizes[l], N = sizes[l + 1], Kp = ceil4(K), Np = ceil4(N);
        std::vector<float> wt((size_t)Np * Kp, 0.0f), bb(Np, 0.0f);
        for (int k = 0; k < K; k++)
            for (int n = 0; n < N; n++) wt[(size_t)n * Kp + k] = W[l][(size_t)k * N + n];
        for (int n = 0; n < N; n++) bb[n] = b[l][n];
        if (!wts[l].alloc(wt.size()) || !bias[l].alloc(Np) || !acts[l + 1].alloc((size_t)M * Np)) { err = "mlp: out of device memory"; break; }
        if (cudaMemcpy(wts[l].p, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(bias[l].p, bb.data(), Np * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
            err = "mlp: upload failed";
    }
    // ---- the layers
    double fl = 0.0;
    if (!err) cudaEventRecord(e0, s);
    for (int l = 0; l < nlayer && !err; l++) {
        const int Kp = ceil4(sizes[l]), Np = ceil4(sizes[l + 1]);
        const int a = act[l];
        const float *X = acts[l].p;
        float *Y = acts[l + 1].p;
        if (a == 1) err = DenseGemm<cutlass::epilogue::thread::ReLu>::run((int)M, Np, Kp, X, wts[l].p, bias[l].p, Y, s);
        else if (a == 2) err = DenseGemm<cutlass::epilogue::thread::Sigmoid>::run((int)M, Np, Kp, X, wts[l].p, bias[l].p, Y, s);
        else if (a == 3) err = DenseGemm<cutlass::epilogue::thread::Tanh>::run((int)M, Np, Kp, X, wts[l].p, bias[l].p, Y, s);
        else err = DenseGemm<cutlass::epilogue::thread::Identity>::run((int)M, Np, Kp, X, wts[l].p, bias[l].p, Y, s);
        if (!err && a >= 4) activation_kernel<<<(unsigned)(((size_t)M * Np + 255) / 256), 256, 0, s>>>(Y, (size_t)M * Np, a);
        fl += 2.0 * (double)M * sizes[l] * sizes[l + 1];
    }
    if (!err) {
        cudaEventRecord(e1, s);
        if (cudaEventSynchronize(e1) != cudaSuccess) err = "mlp: a layer kernel failed";
        else if (ms) cudaEventElapsedTime(ms, e0, e1);
    }
    if (flops) *flops = fl;
    // ---- the prediction down (the padded columns stay behind)
    if (!err) {
        const int N = sizes[nlayer], Np = ceil4(N);
        if (cudaMemcpy2D(y, (size_t)N * sizeof(float), acts[nlayer].p, (size_t)Np * sizeof(float), (size_t)N * sizeof(float), (size_t)M,
                         cudaMemcpyDeviceToHost) != cudaSuccess)
            err = "mlp: download failed";
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (err) cudaGetLastError();
    return err;
}

}  // namespace smash
