// api.cu -- C ABI of libsmash_b200.so (include/smash_b200.h): plan cache, host<->device plumbing and the
// O(nrow*ncol) host-side pieces of base_forward / BASE_FORWARD_B that are not worth a kernel
// ((de)normalisation, regularisation term and its adjoint, hyper-parameter mapping).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/smash_b200.h"
#include "kernels.cuh"
#include "route_graph.hpp"
#include "field_kernels.cuh"
#include "split_kernels.cuh"
#include "topology.hpp"

using namespace smash;

// ------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(e_ == cudaErrorMemoryAllocation ? SMASH_B200_ENOMEM : SMASH_B200_ECUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                      \
    } while (0)
#define TRY(call)                \
    do {                         \
        int rc_ = (call);        \
        if (rc_ != 0) return rc_; \
    } while (0)

static std::map<std::string, long long> &options() {
    static std::map<std::string, long long> o;
    return o;
}
static long long option(const char *name, long long dflt) {
    auto it = options().find(name);
    if (it != options().end()) return it->second;
    std::string env = std::string("SMASH_B200_") + name;
    for (auto &ch : env) ch = (char)toupper(ch);
    const char *v = getenv(env.c_str());
    return v ? atoll(v) : dflt;
}

// ------------------------------------------------------------------------------------------------
// device buffer helper
// ------------------------------------------------------------------------------------------------
template <typename T> struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    ~DBuf() { if (p) cudaFree(p); }
    int ensure(size_t count) {
        if (count <= n) return 0;
        if (p) { cudaFree(p); p = nullptr; n = 0; }
        if (count == 0) return 0;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) { cudaGetLastError(); return fail(SMASH_B200_ENOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e)); }
        n = count;
        return 0;
    }
    int upload(const std::vector<T> &v, cudaStream_t s) {
        TRY(ensure(v.size() ? v.size() : 1));
        if (v.size()) CU(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
        return 0;
    }
};

// state of the split engine (split_kernels.cuh): route graph on the device, row / tape buffers, tensor maps
struct SplitState {
    RouteGraph rg;
    int S = 0, W = 0, nwin = 0, Tp = 0;
    int64_t qpitch = 0;
    DBuf<int32_t> d_flwacc, d_up_begin, d_down, d_down_task, d_down_need, d_rlist, d_rindex, d_task_begin, d_task_cells, d_gfirst, d_gnext, d_cell_task;
    // dynamic scheduling of the ticketed chains (SplitArgs::dyn)
    DBuf<int32_t> d_queue, d_queue0, d_ndep, d_ndep0, d_cons, d_qid, d_qoff;
    DBuf<unsigned int> d_qctl, d_qctl0;
    int dyn_nq = 0;
    DBuf<RouteUp> d_up, d_tup;
    DBuf<TaskCell> d_tcell;
    DBuf<uint8_t> d_down_lag;
    DBuf<float> d_pk_prcp, d_pk_pet, d_rows, d_rows_hr, d_rows_w, d_tape_hp, d_tape_hft, d_hcar, d_gcar;
    DBuf<int> d_done, d_rdone;
    CUtensorMap tm_prcp, tm_pet, tm_hp, tm_hft;
    const float *tape_mapped = nullptr;
    size_t tape_rows = 0;
    SplitTopo topo{};
    DBuf<uint8_t> d_deep;          // graph of the deep cells only: SplitTopo::deep
    // checkpointed gradient runs (adjoint_checkpoint): the tape, hr and w rows hold ONE routing window; the states at every
    // window start are kept and the window is replayed with the tape on before its reverse sweep
    bool ckpt = false;
    DBuf<float> d_ckpt;            // [nwin][5][npad]: hp, hft, hlr of the source cells (fstates), hcar, q at the step before
    DBuf<float> d_rows_seg, d_wnext, d_qprev;
};

// tick pass (tick_kernels.cu): forward runs of large domains; the pit pairs are routed afterwards from rows
struct WindowState {
    bool on = false;
    TickTopoHost host;
    SplitState deep;               // route graph of the pit pairs and its device image (graph members only)
    DBuf<int32_t> d_meta, d_upoff, d_ups, d_reach, d_wunits, d_cons1, d_cons2, d_need;
    DBuf<uint8_t> d_rounds;
    DBuf<float4> d_cc;
    DBuf<float> d_X;
    DBuf<int> d_cnt, d_err;
    int nwarp = 0, maxu = 0, variant = 8;
    TkTopo topo{};
};

// subtree engine (sub_kernels.cu): one pass over the forcing with the engine's own cell order; pit pairs afterwards from rows
struct SubState {
    bool on = false, have_forcing = false;
    SubTopoHost host;
    SplitState pits;               // route graph of the pit pairs and its device image
    int npad2 = 0;
    DBuf<int32_t> d_cell, d_rec, d_xout, d_extoff, d_extlist, d_idx_raw, d_pit_j, d_pit_jp;
    int npit = 0;
    DBuf<uint32_t> d_child;
    DBuf<uint8_t> d_kmax, d_ext;
    DBuf<float> d_X, d_pk_prcp, d_pk_pet, d_qdom, d_netp;
    DBuf<int> d_err;
    CUtensorMap tm_prcp, tm_pet;
    std::vector<int32_t> idx_sparse, idx_dense;   // raw forcing column of every engine column (-1 on empty lanes)
    SbTopo topo{};
};

// Opt-in (option "pin_host" = 1): large caller-owned host arrays (forcing in, domain series out) are page-locked in place
// the first time they are seen (cudaHostRegister), so that the copies run at PCIe speed; the registration is remembered
// by address until smash_b200_clear_cache().  The caller promises not to free such an array before clearing the cache
// (a dangling registration makes later allocations fail), which is why it is off by default.
static std::map<uintptr_t, size_t> &pinned_ranges() {
    static std::map<uintptr_t, size_t> m;
    return m;
}
static void pin_host(const void *ptr, size_t bytes) {
    if (!ptr || bytes < ((size_t)option("pin_min_mb", 64) << 20) || !option("pin_host", 0)) return;
    const uintptr_t page = 4096, lo = reinterpret_cast<uintptr_t>(ptr) & ~(page - 1);
    const uintptr_t hi = (reinterpret_cast<uintptr_t>(ptr) + bytes + page - 1) & ~(page - 1);
    auto &m = pinned_ranges();
    auto it = m.find(lo);
    if (it != m.end()) {
        if (it->second >= hi - lo) return;
        cudaHostUnregister(reinterpret_cast<void *>(lo));
        m.erase(it);
    }
    if (cudaHostRegister(reinterpret_cast<void *>(lo), hi - lo, cudaHostRegisterDefault) == cudaSuccess) m[lo] = hi - lo;
    else cudaGetLastError();
}
static void unpin_all() {
    for (auto &kv : pinned_ranges()) cudaHostUnregister(reinterpret_cast<void *>(kv.first));
    cudaGetLastError();
    pinned_ranges().clear();
}

struct SmashPlan {
    Topology tp;
    DeviceTopology dtp{};
    int engine = 0;                 // 0: fused tick wavefront (kernels.cu), 1: split reservoirs / routing (split_kernels.cu)
    int sparse = 0;                 // setup.sparse_storage the plan was built for
    bool ensemble = false;          // plan made for compute_multiple_run / a multi-member plan: lane = member routing
    SplitState sp;
    WindowState win;
    SubState sub;
    bool sub_ran = false;           // the last forward sweep was the subtree engine: its error word is checked at the next synchronisation
    int structure = SMASH_STRUCTURE_GR_A;   // setup%structure; the others run forward only, on the row passes (struct_kernels.cu)
    bool plan_api = false;          // made by smash_b200_plan_create: results may stay in the engine's own column order
    std::vector<int32_t> col_cell;  // per device column (slot / cell j): flat rect index or -1 on padding
    int ncols = 0;                  // columns of the per-cell device arrays (fields, fstates, grad)
    int nmember = 0;
    float dt = 0, dx = 0;
    int ncell = 0;
    cudaStream_t stream = nullptr;
    // streamed forward runs (large host-resident forcing): 256-step windows, forcing in / domain series out on streams of
    // their own so that the two PCIe directions and the kernels overlap
    bool small_windows = false;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_cmp;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    cudaEvent_t evk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // per-kernel marks
    int kmark[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // 1: evk[i] was recorded in the last run
    int launches = 0;
    bool tick_ran = false;          // the last forward sweep was the tick pass: its error word is checked at the next synchronisation
    // topology on device
    DBuf<int32_t> d_cell, d_off, d_flwacc, d_up_begin, d_down_kind, d_down_lane, d_gfirst, d_gnext, d_hmax, d_sparse_k;
    DBuf<uint8_t> d_late, d_early, d_bflags;
    DBuf<UpEntry> d_up;
    DBuf<ExtRef> d_ext, d_rext;
    DBuf<int64_t> d_tick_base;
    // data
    DBuf<float> d_splanes, d_sfields, d_sfstates;   // structures other than gr-a: [24][ncell], [m][24][npad], [m][8][npad]
    DBuf<int32_t> d_sample_plane;
    DBuf<float> d_forcing, d_raw_prcp, d_raw_pet, d_planes, d_fields, d_fstates, d_qsim, d_qdom, d_netp, d_tape, d_qsim_b,
        d_wdom, d_grad, d_cost_jobs, d_qobs, d_area, d_wgauge, d_sample, d_out;
    DBuf<int32_t> d_gauge_flwacc, d_sample_field, d_mask_event;
    DBuf<float> d_mean_prcp, d_sig_scratch;
    DBuf<int> d_prog, d_rprog;
    DBuf<unsigned int> d_ticket;
    DBuf<double> d_sum;
    // rectangle-level work on the device (field_kernels.cu): descriptors, hyper-parameters, Jreg planes, gradient planes
    DBuf<float> d_desc, d_hyper, d_hyper_b, d_rect, d_rectb, d_jr_mat, d_jr_bgd, d_jr_b, d_jreg;
    DBuf<double> d_partial;
    DBuf<int32_t> d_active;
    const void *desc_ptr = nullptr;
    uint64_t desc_version = 0;
    bool have_active = false;
    bool need_qdom = false;
    bool have_forcing = false, have_qobs = false, have_tape = false;
    const void *forcing_ptr = nullptr;
    uint64_t forcing_version = 0;
    int forcing_sparse = -1;
    std::vector<int32_t> gauge_flwacc;

    ~SmashPlan() {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev2) cudaEventDestroy(ev2);
        for (auto &e : evk) if (e) cudaEventDestroy(e);
        for (auto &e : ev_in) if (e) cudaEventDestroy(e);
        for (auto &e : ev_cmp) if (e) cudaEventDestroy(e);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
        if (stream) cudaStreamDestroy(stream);
    }
};

// gr-b, gr-c, vic-a: their own reservoir pass (struct_kernels.cu), forward only.  gr-d runs on gr-a's kernels (forward, tape and
// reverse sweep) with the shares 1 / 0 and exc = 0 (SplitArgs::grd), which makes every statement gr_d_forward's / GR_D_FORWARD_B's.
static bool struct_pass(const SmashPlan &pl) { return pl.structure != SMASH_STRUCTURE_GR_A && pl.structure != SMASH_STRUCTURE_GR_D; }

static std::mutex g_mu;
static std::map<std::string, std::unique_ptr<SmashPlan>> g_plans;

static int check_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(SMASH_B200_ENODEV, "no CUDA device available (%s): libsmash_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    return 0;
}

static int pick_block(int nactive) {
    long long b = option("block", 0);
    if (b > 0) return (int)b;
    if (nactive <= 512) return std::max(32, ((nactive + 31) / 32) * 32);
    return 256;
}

static int math_mode();

// engine option: -1 (default) picks the split engine whenever the mesh supports it and the fast math mode is on
static int pick_engine(int want) {
    const long long o = option("engine", -1);
    if (o >= 0) return (int)o;
    if (want >= 0) return want;
    return math_mode() ? 1 : 0;
}

static int split_build(SmashPlan &pl, const SmashSetup *setup, const SmashMesh *mesh, bool *unsupported);

static int plan_build(SmashPlan &pl, const SmashSetup *setup, const SmashMesh *mesh, int nmember, int engine = -1) {
    if (!setup || !mesh) return fail(SMASH_B200_EINVAL, "setup / mesh is NULL");
    if (setup->structure < SMASH_STRUCTURE_GR_A || setup->structure > SMASH_STRUCTURE_VIC_A)
        return fail(SMASH_B200_EINVAL, "unknown structure %d", setup->structure);
    pl.structure = setup->structure;
    if (!mesh->flwdir || !mesh->flwacc || !mesh->active_cell || !mesh->path) return fail(SMASH_B200_EINVAL, "mesh arrays missing");
    if (mesh->ng > 0 && (!mesh->gauge_pos || !mesh->area)) return fail(SMASH_B200_EINVAL, "mesh.gauge_pos / area missing");
    // count computed cells to pick the block size
    int nact = 0;
    const int ncell = mesh->nrow * mesh->ncol;
    for (int c = 0; c < ncell; c++)
        if (mesh->active_cell[c] == 1 && (!mesh->local_active_cell || mesh->local_active_cell[c] == 1)) nact++;
    pl.sparse = setup->sparse_storage ? 1 : 0;
    pl.engine = pick_engine(engine);
    if (pl.structure != SMASH_STRUCTURE_GR_A) pl.engine = 1;   // their reservoir pass belongs to the row passes
    if (pl.engine == 1) {
        bool unsupported = false;
        const int rc = split_build(pl, setup, mesh, &unsupported);
        if (rc == 0) return 0;
        if (!unsupported) return rc;
        if (pl.structure != SMASH_STRUCTURE_GR_A)
            return fail(SMASH_B200_EUNSUPPORTED, "structure %d on a mesh the row passes do not handle: %s", pl.structure, g_err.c_str());
        pl.engine = 0;   // mesh shape the split engine does not handle: fused engine
        pl.sp = SplitState();
    }
    const int B = pick_block(nact);
    std::string err = build_topology(pl.tp, mesh->nrow, mesh->ncol, mesh->ng, setup->ntime_step, mesh->flwdir, mesh->flwacc,
                                     mesh->active_cell, mesh->local_active_cell, mesh->path, mesh->gauge_pos, B);
    if (!err.empty()) return fail(SMASH_B200_EINVAL, "%s", err.c_str());
    pl.dt = setup->dt; pl.dx = mesh->dx; pl.ncell = ncell; pl.nmember = 0;
    Topology &tp = pl.tp;
    CU(cudaStreamCreateWithFlags(&pl.stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&pl.ev0)); CU(cudaEventCreate(&pl.ev1)); CU(cudaEventCreate(&pl.ev2));
    for (auto &e : pl.evk) CU(cudaEventCreate(&e));
    cudaStream_t s = pl.stream;
    TRY(pl.d_cell.upload(tp.cell, s)); TRY(pl.d_off.upload(tp.off, s)); TRY(pl.d_flwacc.upload(tp.flwacc, s));
    TRY(pl.d_late.upload(tp.late, s)); TRY(pl.d_early.upload(tp.early, s)); TRY(pl.d_up_begin.upload(tp.up_begin, s));
    TRY(pl.d_up.upload(tp.up, s)); TRY(pl.d_ext.upload(tp.ext, s)); TRY(pl.d_rext.upload(tp.rext, s));
    TRY(pl.d_down_kind.upload(tp.down_kind, s)); TRY(pl.d_down_lane.upload(tp.down_lane, s));
    TRY(pl.d_gfirst.upload(tp.gauge_first, s)); TRY(pl.d_gnext.upload(tp.gauge_next, s));
    TRY(pl.d_hmax.upload(tp.hmax, s)); TRY(pl.d_tick_base.upload(tp.tick_base, s)); TRY(pl.d_bflags.upload(tp.flags, s));
    TRY(pl.d_sparse_k.upload(tp.sparse_k, s));
    TRY(pl.d_ticket.ensure(2)); TRY(pl.d_sum.ensure(1));
    DeviceTopology &d = pl.dtp;
    d.T = tp.T; d.B = tp.B; d.nblocks = tp.nblocks; d.nslots = tp.nslots; d.ng = tp.ng; d.total_ticks = tp.total_ticks;
    d.cell = pl.d_cell.p; d.off = pl.d_off.p; d.flwacc = pl.d_flwacc.p; d.late = pl.d_late.p; d.early = pl.d_early.p;
    d.up_begin = pl.d_up_begin.p; d.up = pl.d_up.p; d.ext = pl.d_ext.p; d.rext = pl.d_rext.p;
    d.down_kind = pl.d_down_kind.p; d.down_lane = pl.d_down_lane.p; d.gauge_first = pl.d_gfirst.p; d.gauge_next = pl.d_gnext.p;
    d.hmax = pl.d_hmax.p; d.tick_base = pl.d_tick_base.p; d.bflags = pl.d_bflags.p;
    pl.need_qdom = false;
    for (auto f : tp.flags) if (f & BLK_PUBLISH) pl.need_qdom = true;
    pl.gauge_flwacc.assign(std::max(1, mesh->ng), 1);
    if (mesh->ng > 0) {
        std::vector<float> area(mesh->area, mesh->area + mesh->ng);
        for (int g = 0; g < mesh->ng; g++)
            pl.gauge_flwacc[g] = mesh->flwacc[(mesh->gauge_pos[g] - 1) + (size_t)(mesh->gauge_pos[g + mesh->ng] - 1) * mesh->nrow];
        TRY(pl.d_area.upload(area, s));
        TRY(pl.d_gauge_flwacc.upload(pl.gauge_flwacc, s));
    }
    TRY(pl.d_planes.ensure((size_t)NFIELD * ncell));
    CU(cudaStreamSynchronize(s));
    (void)nmember;
    pl.col_cell = tp.cell;
    pl.ncols = tp.nslots;
    return 0;
}

// device image of a route graph (sp.rg -> sp.topo): the full graph of a mesh, or the deep-cell graph of the window pass
static int upload_route_graph(SplitState &sp, int ng, cudaStream_t s) {
    RouteGraph &rg = sp.rg;
    const int npad = rg.npad;
    std::vector<int32_t> flw(npad, 1), down(npad, -1), down_task(npad, -1), gfirst(npad, -1);
    std::vector<uint8_t> down_lag(npad, 0);
    std::vector<int32_t> down_need(npad, 0xffff), pos_in_task(rg.n, 0);
    for (int t = 0; t < rg.ntask; t++)
        for (int e = rg.task_begin[t]; e < rg.task_begin[t + 1]; e++) pos_in_task[rg.task_cells[e]] = e - rg.task_begin[t];
    for (int j = 0; j < rg.n; j++) {
        const int d = rg.down[j];
        if (d >= 0 && rg.down_task[j] >= 0 && rg.down_task[j] < rg.nchain) {
            const int len = rg.task_begin[rg.down_task[j] + 1] - rg.task_begin[rg.down_task[j]];
            if (len < 0xffff) down_need[j] = len - pos_in_task[d];   // cells done counted from the tail, d included
        }
    }
    for (int j = 0; j < rg.n; j++) {
        flw[j] = rg.flwacc[j]; down[j] = rg.down[j]; down_task[j] = rg.down_task[j]; gfirst[j] = rg.gauge_first[j];
        if (rg.down[j] >= 0 && j > rg.down[j]) down_lag[j] = 1;   // producer later in path: the reader sees its previous step
    }
    TRY(sp.d_flwacc.upload(flw, s)); TRY(sp.d_up_begin.upload(rg.up_begin, s)); TRY(sp.d_up.upload(rg.up, s));
    TRY(sp.d_down.upload(down, s)); TRY(sp.d_down_task.upload(down_task, s)); TRY(sp.d_down_need.upload(down_need, s)); TRY(sp.d_down_lag.upload(down_lag, s));
    TRY(sp.d_task_begin.upload(rg.task_begin, s)); TRY(sp.d_task_cells.upload(rg.task_cells, s));
    TRY(sp.d_gfirst.upload(gfirst, s)); TRY(sp.d_gnext.upload(rg.gauge_next, s));
    std::vector<int32_t> rlist, rindex(npad, -1);
    for (int j = 0; j < rg.n; j++) if (rg.flwacc[j] > 1) { rindex[j] = (int32_t)rlist.size(); rlist.push_back(j); }
    TRY(sp.d_rlist.upload(rlist, s)); TRY(sp.d_rindex.upload(rindex, s));
    TRY(sp.d_tcell.upload(rg.tcell, s)); TRY(sp.d_tup.upload(rg.tup, s));
    std::vector<int32_t> cell_task(npad, -1);
    for (int j = 0; j < rg.n; j++) cell_task[j] = rg.cell_task[j];
    TRY(sp.d_cell_task.upload(cell_task, s));
    {   // ---- ready queues of the dynamic scheduler: one per basin of a long river (longest first), one for the rest
        const int nticket = rg.nchain - rg.nded;
        std::vector<int32_t> cons(std::max(1, rg.nchain), -1), ndep0(std::max(1, nticket), 0), qid(std::max(1, nticket), 0);
        for (int t = 0; t < rg.nchain; t++) {
            const int tail = rg.task_cells[rg.task_begin[t + 1] - 1];
            const int ct = rg.down_task[tail];
            if (ct >= 0 && ct < nticket) cons[t] = ct;
        }
        for (int t = 0; t < nticket; t++)
            for (int i = rg.task_begin[t]; i < rg.task_begin[t + 1]; i++) {
                const TaskCell &tc = rg.tcell[i];
                for (int e = 0; e < (tc.meta >> 8); e++) if (rg.tup[tc.up_off + e].task >= 0) ndep0[t]++;
            }
        // basins (cells that drain to the same outlet) ranked by their longest chain
        std::vector<int32_t> root(rg.n), longest(rg.n, 0);
        for (int j = rg.n - 1; j >= 0; j--) root[j] = (rg.down[j] > j) ? root[rg.down[j]] : j;
        for (int t = 0; t < rg.nchain; t++) {
            const int r = root[rg.task_cells[rg.task_begin[t]]];
            longest[r] = std::max(longest[r], rg.task_begin[t + 1] - rg.task_begin[t]);
        }
        std::vector<int32_t> roots;
        for (int j = 0; j < rg.n; j++) if (root[j] == j && longest[j] >= (int)option("route_ded_min", 96)) roots.push_back(j);
        std::stable_sort(roots.begin(), roots.end(), [&](int x, int y) { return longest[x] > longest[y]; });
        const int nq = (int)std::min<size_t>(roots.size(), (size_t)std::min<long long>(15, std::max<long long>(0, option("route_queues", 12)))) + 1;
        std::vector<int32_t> qof_root(rg.n, nq - 1);
        for (int k = 0; k + 1 < nq; k++) qof_root[roots[k]] = k;
        std::vector<int32_t> qoff(nq + 1, 0), fill(nq, 0);
        for (int t = 0; t < nticket; t++) { qid[t] = qof_root[root[rg.task_cells[rg.task_begin[t]]]]; qoff[qid[t] + 1]++; }
        for (int q = 0; q < nq; q++) qoff[q + 1] += qoff[q];
        std::vector<int32_t> queue0(std::max(1, nticket), -1);
        std::vector<unsigned int> qctl0(32, 0u);
        for (int t = 0; t < nticket; t++)                        // chains without tributary chains are ready from the start, in task order
            if (ndep0[t] == 0) { queue0[qoff[qid[t]] + fill[qid[t]]++] = t; }
        for (int q = 0; q < nq; q++) qctl0[16 + q] = (unsigned int)fill[q];
        sp.dyn_nq = nq;
        TRY(sp.d_cons.upload(cons, s)); TRY(sp.d_ndep0.upload(ndep0, s)); TRY(sp.d_qid.upload(qid, s)); TRY(sp.d_qoff.upload(qoff, s));
        TRY(sp.d_queue0.upload(queue0, s)); TRY(sp.d_qctl0.upload(qctl0, s));
        TRY(sp.d_queue.ensure(queue0.size())); TRY(sp.d_ndep.ensure(ndep0.size())); TRY(sp.d_qctl.ensure(32));
    }
    SplitTopo &t = sp.topo;
    t.n = rg.n; t.npad = npad; t.ng = ng; t.ntask = rg.ntask; t.nchain = rg.nchain; t.nded = rg.nded;
    t.flwacc = sp.d_flwacc.p; t.up_begin = sp.d_up_begin.p; t.up = sp.d_up.p; t.down = sp.d_down.p; t.down_task = sp.d_down_task.p; t.down_need = sp.d_down_need.p;
    t.down_lag = sp.d_down_lag.p; t.task_begin = sp.d_task_begin.p; t.task_cells = sp.d_task_cells.p;
    t.gauge_first = sp.d_gfirst.p; t.gauge_next = sp.d_gnext.p; t.cell_task = sp.d_cell_task.p;
    t.nrouted = (int)rlist.size(); t.rlist = sp.d_rlist.p; t.rindex = sp.d_rindex.p;
    t.tcell = reinterpret_cast<const int4 *>(sp.d_tcell.p); t.tup = reinterpret_cast<const int2 *>(sp.d_tup.p);
    return 0;
}

// ---- tick pass: classes, reaches, stages, units per warp ------------------------------------------
// Forward runs of large domains (no tape, one member, math = 1): tick_kernels.cu computes reservoirs and routing of the whole
// domain in one kernel.  Not eligible (small domain, option off, unsupported mesh): win.on stays false and the run uses the
// row-based passes.
static int window_build(SmashPlan &pl, const SmashMesh *mesh) {
    WindowState &wn = pl.win;
    const RouteGraph &rg = pl.sp.rg;
    wn.on = false;
    if (!option("tick_pass", 0) || rg.n < option("tick_min_cells", 65536)) return 0;
    wn.variant = (int)option("tick_variant", 6);
    CU(tick_grid_warps(wn.variant, (int)option("tick_ctas_per_sm", 0), &wn.nwarp));
    // stages with at least a quarter of a unit per warp run two ticks after their producers (option tick_slack = 0: one tick)
    const int slack = option("tick_slack", 1) ? std::max(1, wn.nwarp / (int)std::max<long long>(1, option("tick_slack_div", 4))) : 0;
    if (!build_tick_topo(rg, (int)option("shallow_acc", 32), wn.host, slack).empty()) return 0;
    std::vector<int32_t> wunits;
    deal_tick_units(wn.host, wn.nwarp, wunits, wn.maxu);
    if (wn.maxu > tick_max_units()) return 0;                         // a mesh too large for one resident grid: row-based passes
    cudaStream_t s = pl.stream;
    const size_t npad = (size_t)rg.npad, ntile = npad / 32;
    if (wn.host.npair_cells > 0) {
        std::string err = build_route_graph(wn.deep.rg, mesh->nrow, mesh->ncol, mesh->ng, mesh->flwdir, mesh->flwacc, mesh->active_cell,
                                            mesh->local_active_cell, mesh->path, mesh->gauge_pos, (int)option("route_ded_min", 96),
                                            (int)option("route_ded_max", 64), 0, 0, wn.host.pair.data());
        if (!err.empty() || wn.deep.rg.n != rg.n || wn.deep.rg.nchain != 0) return 0;
        TRY(upload_route_graph(wn.deep, mesh->ng, s));
        std::vector<uint8_t> mask(npad, 0);
        std::copy(wn.host.pair.begin(), wn.host.pair.end(), mask.begin());
        TRY(wn.deep.d_deep.upload(mask, s));
        wn.deep.topo.deep = wn.deep.d_deep.p;
        TRY(wn.deep.d_done.ensure(std::max<size_t>(1, 2 * (size_t)wn.deep.rg.ntask)));
    }
    TRY(wn.d_meta.upload(wn.host.meta, s)); TRY(wn.d_upoff.upload(wn.host.upoff, s)); TRY(wn.d_ups.upload(wn.host.ups, s));
    TRY(wn.d_rounds.upload(wn.host.tile_rounds, s));
    if (wn.host.nreach > 0) TRY(wn.d_reach.upload(wn.host.reach_cells, s));
    TRY(wn.d_wunits.upload(wunits, s));
    TRY(wn.d_cons1.upload(wn.host.cons1, s)); TRY(wn.d_cons2.upload(wn.host.cons2, s)); TRY(wn.d_need.upload(wn.host.need, s));
    TRY(wn.d_cc.ensure(npad)); TRY(wn.d_err.ensure(1));
    TkTopo &t = wn.topo;
    t.n = rg.n; t.npad = (int)npad; t.ntile = (int)ntile; t.nreach = wn.host.nreach; t.ng = mesh->ng;
    t.meta = wn.d_meta.p; t.upoff = wn.d_upoff.p; t.ups = wn.d_ups.p; t.tile_rounds = wn.d_rounds.p; t.reach_cells = wn.d_reach.p;
    t.cons1 = wn.d_cons1.p; t.cons2 = wn.d_cons2.p; t.need = wn.d_need.p;
    t.gauge_first = pl.sp.d_gfirst.p; t.gauge_next = pl.sp.d_gnext.p;
    wn.on = true;
    return 0;
}

// ---- subtree engine: tiles, exchange slots, pit-pair graph --------------------------------------------
// Forward runs of large domains (no tape, one member, math = 1), option sub_engine.  Not eligible: sub.on stays false.
static int sub_build(SmashPlan &pl, const SmashMesh *mesh) {
    SubState &sb = pl.sub;
    const RouteGraph &rg = pl.sp.rg;
    sb.on = false;
    // option sub_engine: -1 (default) = plans of the plan API (results stay on the device, in engine order); 1 = every eligible
    // plan (the drop-in calls get their domain series back in cell order through a scatter pass); 0 = off
    const long long mode = option("sub_engine", -1);
    if (!(mode == 1 || (mode == -1 && pl.plan_api)) || rg.n < option("sub_min_cells", 65536)) return 0;
    if (!build_sub_topo(rg, 8, sb.host).empty()) return 0;
    cudaStream_t s = pl.stream;
    const size_t npad = (size_t)rg.npad;
    sb.npad2 = sb.host.ntile * 32;
    if (sb.host.pair_cells > 0) {
        std::string err = build_route_graph(sb.pits.rg, mesh->nrow, mesh->ncol, mesh->ng, mesh->flwdir, mesh->flwacc, mesh->active_cell,
                                            mesh->local_active_cell, mesh->path, mesh->gauge_pos, (int)option("route_ded_min", 96),
                                            (int)option("route_ded_max", 64), 0, 0, sb.host.pair.data());
        if (!err.empty() || sb.pits.rg.n != rg.n || sb.pits.rg.nchain != 0) return 0;
        TRY(upload_route_graph(sb.pits, mesh->ng, s));
        std::vector<uint8_t> mask(npad, 0);
        std::copy(sb.host.pair.begin(), sb.host.pair.end(), mask.begin());
        TRY(sb.pits.d_deep.upload(mask, s));
        sb.pits.topo.deep = sb.pits.d_deep.p;
        TRY(sb.pits.d_done.ensure(std::max<size_t>(1, 2 * (size_t)sb.pits.rg.ntask)));
    }
    sb.idx_sparse.assign(sb.npad2, -1); sb.idx_dense.assign(sb.npad2, -1);
    std::vector<int32_t> pit_j, pit_jp;                                   // the pit cells: cell order j and engine column j'
    for (int jp = 0; jp < sb.npad2; jp++) {
        const int j = sb.host.cell[jp];
        if (j < 0) continue;
        sb.idx_sparse[jp] = rg.sparse_k[j]; sb.idx_dense[jp] = rg.cell[j];
        if (sb.host.pair[j]) { pit_j.push_back(j); pit_jp.push_back(jp); }
    }
    sb.npit = (int)pit_j.size();
    if (pit_j.empty()) { pit_j.push_back(0); pit_jp.push_back(0); }
    TRY(sb.d_cell.upload(sb.host.cell, s)); TRY(sb.d_rec.upload(sb.host.rec, s)); TRY(sb.d_child.upload(sb.host.child, s));
    TRY(sb.d_xout.upload(sb.host.xout, s)); TRY(sb.d_extoff.upload(sb.host.extoff, s)); TRY(sb.d_extlist.upload(sb.host.extlist, s));
    TRY(sb.d_kmax.upload(sb.host.tile_kmax, s)); TRY(sb.d_ext.upload(sb.host.tile_ext, s));
    TRY(sb.d_pit_j.upload(pit_j, s)); TRY(sb.d_pit_jp.upload(pit_jp, s));
    TRY(sb.d_err.ensure(1));
    SbTopo &t = sb.topo;
    t.ntile = sb.host.ntile; t.ng = mesh->ng; t.nslot = sb.host.nslot; t.dmax = sb.host.dmax;
    t.cell = sb.d_cell.p; t.rec = sb.d_rec.p; t.child = reinterpret_cast<const uint2 *>(sb.d_child.p); t.xout = sb.d_xout.p;
    t.extoff = sb.d_extoff.p; t.extlist = sb.d_extlist.p; t.tile_kmax = sb.d_kmax.p; t.tile_ext = sb.d_ext.p;
    sb.on = true;
    return 0;
}

// ---- split engine: build -------------------------------------------------------------------------
// Gradient runs keep a tape of 24 bytes per cell-step (hp0, hft0, hr_imd, w, the q rows and their padding).  Beyond
// "tape_budget_mb" (default 16 GB) -- or always with option adjoint_checkpoint = 1 -- the reverse sweep runs window by window
// from checkpointed states instead (256-step windows), so that the tape holds one window.
static bool checkpoint_wanted(int ncells, int T) {
    const long long mode = option("adjoint_checkpoint", -1);
    if (mode == 0) return false;
    if (mode == 1) return T > 256;
    const double tape = 24.0 * (double)ncells * (double)T;
    return T > 256 && tape > (double)option("tape_budget_mb", 16384) * 1048576.0;
}

static int split_build(SmashPlan &pl, const SmashSetup *setup, const SmashMesh *mesh, bool *unsupported) {
    SplitState &sp = pl.sp;
    RouteGraph &rg = sp.rg;
    std::string err = build_route_graph(rg, mesh->nrow, mesh->ncol, mesh->ng, mesh->flwdir, mesh->flwacc, mesh->active_cell,
                                        mesh->local_active_cell, mesh->path, mesh->gauge_pos, (int)option("route_ded_min", 96),
                                        (int)option("route_ded_max", 64), 0, 0);
    if (!err.empty()) {
        *unsupported = err.rfind("unsupported", 0) == 0;
        return fail(SMASH_B200_EINVAL, "%s", err.c_str());
    }
    const int ncell = mesh->nrow * mesh->ncol;
    pl.dt = setup->dt; pl.dx = mesh->dx; pl.ncell = ncell; pl.nmember = 0;
    Topology &tp = pl.tp;
    tp = Topology();
    tp.nrow = mesh->nrow; tp.ncol = mesh->ncol; tp.ng = mesh->ng; tp.T = setup->ntime_step; tp.B = 256;
    tp.nactive = rg.n; tp.nslots = rg.npad; tp.nblocks = (rg.n + 255) / 256; tp.n_pairs = rg.npair;
    tp.sparse_k.assign(rg.npad, -1);
    for (int j = 0; j < rg.n; j++) tp.sparse_k[j] = rg.sparse_k[j];
    tp.gauge_slot = rg.gauge_cell;
    sp.ckpt = checkpoint_wanted(rg.n, tp.T);
    sp.W = split_pick_window(tp.T, &sp.S, &sp.nwin, pl.small_windows || sp.ckpt);
    sp.Tp = sp.W * sp.nwin;
    CU(cudaStreamCreateWithFlags(&pl.stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&pl.ev0)); CU(cudaEventCreate(&pl.ev1)); CU(cudaEventCreate(&pl.ev2));
    for (auto &e : pl.evk) CU(cudaEventCreate(&e));
    cudaStream_t s = pl.stream;
    const int npad = rg.npad;
    std::vector<int32_t> cell(npad, -1);
    for (int j = 0; j < rg.n; j++) cell[j] = rg.cell[j];
    TRY(pl.d_cell.upload(cell, s)); TRY(pl.d_sparse_k.upload(tp.sparse_k, s));
    TRY(upload_route_graph(sp, mesh->ng, s));
    TRY(pl.d_ticket.ensure(2)); TRY(pl.d_sum.ensure(1));
    // the field gather kernel only needs the column -> cell map
    pl.dtp = DeviceTopology{};
    pl.dtp.T = tp.T; pl.dtp.B = 256; pl.dtp.nblocks = tp.nblocks; pl.dtp.nslots = npad; pl.dtp.ng = mesh->ng; pl.dtp.cell = pl.d_cell.p;
    pl.gauge_flwacc.assign(std::max(1, mesh->ng), 1);
    if (mesh->ng > 0) {
        std::vector<float> area(mesh->area, mesh->area + mesh->ng);
        for (int g = 0; g < mesh->ng; g++)
            pl.gauge_flwacc[g] = mesh->flwacc[(mesh->gauge_pos[g] - 1) + (size_t)(mesh->gauge_pos[g + mesh->ng] - 1) * mesh->nrow];
        TRY(pl.d_area.upload(area, s));
        TRY(pl.d_gauge_flwacc.upload(pl.gauge_flwacc, s));
    }
    TRY(pl.d_planes.ensure((size_t)NFIELD * ncell));
    if (pl.structure == SMASH_STRUCTURE_GR_A) {
        TRY(window_build(pl, mesh));
        TRY(sub_build(pl, mesh));
    }
    CU(cudaStreamSynchronize(s));
    pl.col_cell = cell;
    pl.ncols = npad;
    sp.qpitch = (pl.sparse && rg.direct) ? rg.n : npad;
    return 0;
}

static int split_members(SmashPlan &pl, int nmember, bool save_q, bool save_netp, bool gradient) {
    SplitState &sp = pl.sp;
    const Topology &tp = pl.tp;
    const size_t nm = (size_t)nmember, npad = (size_t)sp.rg.npad;
    TRY(pl.d_fields.ensure(nm * NFIELD * npad));
    TRY(pl.d_fstates.ensure(nm * 3 * npad));
    if (pl.structure != SMASH_STRUCTURE_GR_A) {
        if (gradient && struct_pass(pl))
            return fail(SMASH_B200_EUNSUPPORTED, "structure %d: the adjoint is implemented for gr-a and gr-d only", pl.structure);
        TRY(pl.d_sfields.ensure(nm * (SMASH_B200_GNP + SMASH_B200_GNS) * npad));
        TRY(pl.d_sfstates.ensure(nm * SMASH_B200_GNS * npad));
    }
    TRY(pl.d_qsim.ensure(std::max<size_t>(1, nm * tp.T * tp.ng)));
    TRY(pl.d_cost_jobs.ensure(nm));
    const size_t nrows = nm * npad * sp.Tp;
    if (sp.d_rows.n < nrows) { TRY(sp.d_rows.ensure(nrows)); CU(cudaMemsetAsync(sp.d_rows.p, 0, nrows * sizeof(float), pl.stream)); }
    TRY(sp.d_hcar.ensure(nm * npad));
    TRY(sp.d_done.ensure(std::max<size_t>(1, 2 * nm * sp.rg.ntask)));   // done flags + block counters of the river reaches
    if (save_q) TRY(pl.d_qdom.ensure(nm * tp.T * sp.qpitch));
    if (save_netp) TRY(pl.d_netp.ensure(nm * tp.T * sp.qpitch));
    if (gradient) {
        const size_t trows = sp.ckpt ? (size_t)sp.W : (size_t)tp.T;     // tape rows per member
        const size_t nseg = sp.ckpt ? nm * npad * sp.W : nrows;
        if (sp.ckpt && nm != 1) return fail(SMASH_B200_EUNSUPPORTED, "checkpointed gradient with more than one member");
        TRY(sp.d_tape_hp.ensure(nm * trows * npad)); TRY(sp.d_tape_hft.ensure(nm * trows * npad));
        if (sp.d_rows_hr.n < nseg) { TRY(sp.d_rows_hr.ensure(nseg)); CU(cudaMemsetAsync(sp.d_rows_hr.p, 0, nseg * sizeof(float), pl.stream)); }
        if (sp.d_rows_w.n < nseg) { TRY(sp.d_rows_w.ensure(nseg)); CU(cudaMemsetAsync(sp.d_rows_w.p, 0, nseg * sizeof(float), pl.stream)); }
        if (sp.ckpt) {
            if (sp.d_rows_seg.n < nseg) { TRY(sp.d_rows_seg.ensure(nseg)); CU(cudaMemsetAsync(sp.d_rows_seg.p, 0, nseg * sizeof(float), pl.stream)); }
            TRY(sp.d_ckpt.ensure((size_t)sp.nwin * 5 * npad)); TRY(sp.d_wnext.ensure(npad)); TRY(sp.d_qprev.ensure(npad));
        }
        TRY(sp.d_gcar.ensure(nm * npad));
        TRY(sp.d_rdone.ensure(std::max<size_t>(1, nm * sp.rg.ntask)));
        TRY(pl.d_qsim_b.ensure(std::max<size_t>(1, nm * tp.T * tp.ng)));
        TRY(pl.d_grad.ensure(nm * NFIELD * npad));
        if (sp.tape_mapped != sp.d_tape_hp.p || sp.tape_rows != nm * trows) {
            const char *err = nullptr;
            if (make_tensor_map_2d(&sp.tm_hp, sp.d_tape_hp.p, npad, nm * trows, npad, &err) ||
                make_tensor_map_2d(&sp.tm_hft, sp.d_tape_hft.p, npad, nm * trows, npad, &err))
                return fail(SMASH_B200_ECUDA, "%s", err);
            sp.tape_mapped = sp.d_tape_hp.p; sp.tape_rows = nm * trows;
        }
    }
    pl.nmember = nmember;
    return 0;
}

// graph: the route graph the routing kernels walk (default: the full graph of the mesh; the window pass hands the chain
// scans the graph of the deep cells)
static SplitArgs split_args(SmashPlan &pl, bool save_q, bool save_netp, SplitState *graph = nullptr) {
    SplitState &sp = pl.sp;
    SplitState &gr = graph ? *graph : pl.sp;
    SplitArgs a{};
    a.tp = gr.topo; a.T = pl.tp.T; a.Tp = sp.Tp; a.W = sp.W; a.nwin = sp.nwin; a.nmember = pl.nmember; a.dt = pl.dt; a.dx = pl.dx;
    a.first_routed = gr.rg.first_routed;
    a.t_begin = 0; a.t_end = pl.tp.T;
    a.dbg_prof = nullptr;
    if (option("dbg_prof", 0) && sp.rg.nded > 0 && sp.rg.nded <= 256) {
        // per dedicated chain / reach: [0] cells [1] cycles [2] cycles blocked on whole-series tributaries (sum over threads)
        // [3] end time ns [4] cycles blocked on streamed blocks (sum over threads) [5] ticks
        static unsigned long long *dp = nullptr;
        const size_t nb = 8 * (size_t)sp.rg.nded * sizeof(unsigned long long);
        if (!dp) { cudaMalloc(&dp, 8 * 256 * sizeof(unsigned long long)); cudaMemset(dp, 0, 8 * 256 * sizeof(unsigned long long)); }
        std::vector<unsigned long long> hv(8 * (size_t)sp.rg.nded);
        cudaMemcpy(hv.data(), dp, nb, cudaMemcpyDeviceToHost);
        unsigned long long tmax = 0, tmin = ~0ull;
        for (int i = 0; i < sp.rg.nded; i++) { tmax = std::max(tmax, hv[8 * i + 3]); if (hv[8 * i + 3]) tmin = std::min(tmin, hv[8 * i + 3]); }
        if (tmax) {
            std::vector<int> idx(sp.rg.nded);
            for (int i = 0; i < sp.rg.nded; i++) idx[i] = i;
            std::sort(idx.begin(), idx.end(), [&](int x, int y) { return hv[8 * x + 3] > hv[8 * y + 3]; });
            fprintf(stderr, "[dbg_prof] %d dedicated tasks, first ended %.3f ms before the last\n", sp.rg.nded, (tmax - tmin) * 1e-6);
            for (int kk = 0; kk < 16 && kk < sp.rg.nded; kk++) {
                const int k = kk < 8 ? kk : sp.rg.nded - 1 - (kk - 8);       // the eight last and the eight first to end
                const unsigned long long *h = &hv[8 * (size_t)idx[k]];
                fprintf(stderr, "[dbg_prof] task %d: %llu cells, %llu cycles, %llu ticks, blocked on tributaries %llu (thread-cycles), on streamed "
                        "blocks %llu, ended %.3f ms before the last\n", idx[k], h[0], h[1], h[5], h[2], h[4], (tmax - h[3]) * 1e-6);
            }
        }
        a.dbg_prof = dp;
    }
    a.save_q = save_q ? 1 : 0; a.save_netp = save_netp ? 1 : 0;
    a.fuse_export = 0;
    a.dyn = (option("route_dynamic", 1) != 0 && pl.nmember == 1 && gr.rg.nchain > gr.rg.nded) ? 1 : 0;
    a.dyn_nq = gr.dyn_nq; a.qctl = gr.d_qctl.p; a.queue = gr.d_queue.p; a.ndep = gr.d_ndep.p; a.cons = gr.d_cons.p; a.qid = gr.d_qid.p;
    a.qoff = gr.d_qoff.p; a.qctl0 = gr.d_qctl0.p; a.queue0 = gr.d_queue0.p; a.ndep0 = gr.d_ndep0.p;
    a.sfields = pl.d_sfields.p; a.sfstates = pl.d_sfstates.p;
    a.grd = pl.structure == SMASH_STRUCTURE_GR_D ? 1 : 0;
    a.fields = pl.d_fields.p; a.fstates = pl.d_fstates.p; a.rows = sp.d_rows.p; a.qdom = pl.d_qdom.p; a.netp = pl.d_netp.p;
    a.qpitch = sp.qpitch; a.qsim = pl.d_qsim.p; a.tape_hp = sp.d_tape_hp.p; a.tape_hft = sp.d_tape_hft.p; a.rows_hr = sp.d_rows_hr.p;
    a.hcar = sp.d_hcar.p; a.done = graph ? gr.d_done.p : sp.d_done.p; a.ticket = pl.d_ticket.p; a.qsim_b = pl.d_qsim_b.p; a.rows_w = sp.d_rows_w.p;
    a.gcar = sp.d_gcar.p; a.grad = pl.d_grad.p; a.rdone = sp.d_rdone.p;
    return a;
}

// forward sweep of the split engine: reservoirs, routing, optional [t][cell] export of the routed cells
static int split_forward(SmashPlan &pl, bool save_q, bool save_netp, bool tape) {
    SplitState &sp = pl.sp;
    const size_t npad = (size_t)sp.rg.npad;
    SplitArgs a = split_args(pl, save_q, save_netp);
    for (int i = 0; i < 8; i++) pl.kmark[i] = 0;
    auto mark = [&](int i) { pl.kmark[i] = cudaEventRecord(pl.evk[i], pl.stream) == cudaSuccess; };
    mark(0);
    // routing state of window 0 = the hlr field
    CU(cudaMemcpy2DAsync(sp.d_hcar.p, npad * sizeof(float), pl.d_fields.p + (size_t)F_HLR * npad, (size_t)NFIELD * npad * sizeof(float),
                         npad * sizeof(float), (size_t)pl.nmember, cudaMemcpyDeviceToDevice, pl.stream));
    if (tape && sp.ckpt) {
        // checkpointed gradient run, pass 1: the plain forward sweep window by window, keeping the states every window starts
        // from (reservoirs, routing stores, the discharge of the step before).  The tape is written when the window is replayed
        // right before its reverse sweep (split_reverse).
        const size_t pb = npad * sizeof(float);
        a.Tp = sp.W; a.qprev = sp.d_qprev.p;
        CU(cudaMemsetAsync(sp.d_qprev.p, 0, pb, pl.stream));
        for (int w = 0; w < sp.nwin; w++) {
            float *ck = sp.d_ckpt.p + (size_t)w * 5 * npad;
            if (w == 0) {
                CU(cudaMemcpyAsync(ck, pl.d_fields.p + (size_t)F_HP * npad, 3 * pb, cudaMemcpyDeviceToDevice, pl.stream));   // hp, hft, hlr
            } else {
                CU(cudaMemcpyAsync(ck, pl.d_fstates.p, 3 * pb, cudaMemcpyDeviceToDevice, pl.stream));
            }
            CU(cudaMemcpyAsync(ck + 3 * npad, sp.d_hcar.p, pb, cudaMemcpyDeviceToDevice, pl.stream));
            CU(cudaMemcpyAsync(ck + 4 * npad, sp.d_qprev.p, pb, cudaMemcpyDeviceToDevice, pl.stream));
            a.t_begin = w * sp.W; a.t_end = std::min(pl.tp.T, (w + 1) * sp.W);
            a.rows = sp.d_rows_seg.p - (ptrdiff_t)w * sp.W;              // the window buffer holds steps [w W, (w + 1) W)
            CU(launch_vertical_forward(a, sp.tm_prcp, sp.tm_pet, math_mode(), false, pl.stream));
            CU(launch_route_forward_window(a, w, false, pl.stream));
            CU(launch_first_step(sp.d_rows_seg.p + (sp.W - 1), sp.W, (int)npad, sp.d_qprev.p, pl.stream));
            pl.launches += 3 + (sp.rg.npair > 0 ? 1 : 0);
        }
        mark(1); mark(2); mark(3);
        return 0;
    }
    if (pl.sub.on && pl.sub.have_forcing && !tape && pl.nmember == 1 && !pl.ensemble && math_mode() == 1) {
        // subtree engine: one pass over the forcing; the results are in engine order (column j' = tile * 32 + lane)
        SubState &sb = pl.sub;
        SbArgs sa{};
        sa.tp = sb.topo; sa.T = pl.tp.T; sa.Tp = sp.Tp; sa.nwin = (pl.tp.T + SB_W - 1) / SB_W; sa.npad = (int)npad;
        sa.dt = pl.dt; sa.dx = pl.dx; sa.save_q = save_q ? 1 : 0; sa.save_netp = save_netp ? 1 : 0;
        sa.fields = pl.d_fields.p; sa.flwacc = sp.d_flwacc.p; sa.gauge_first = sp.d_gfirst.p; sa.gauge_next = sp.d_gnext.p;
        sa.fstates = pl.d_fstates.p; sa.rows = sp.d_rows.p; sa.qpitch = sb.npad2; sa.qsim = pl.d_qsim.p; sa.err = sb.d_err.p;
        sa.nowait = (int)option("sub_nowait", 0);
        const size_t nx = (size_t)std::max(1, sb.host.nslot) * sa.nwin * SB_W, nq = (size_t)pl.tp.T * sb.npad2;
        if (sb.d_X.n < nx) TRY(sb.d_X.ensure(nx));
        if (save_q && sb.d_qdom.n < nq) { TRY(sb.d_qdom.ensure(nq)); CU(cudaMemsetAsync(sb.d_qdom.p, 0, nq * sizeof(float), pl.stream)); }
        if (save_netp && sb.d_netp.n < nq) { TRY(sb.d_netp.ensure(nq)); CU(cudaMemsetAsync(sb.d_netp.p, 0, nq * sizeof(float), pl.stream)); }
        sa.X = sb.d_X.p; sa.qdom = sb.d_qdom.p; sa.netp = sb.d_netp.p;
        CU(launch_sub_forward(sa, sb.tm_prcp, sb.tm_pet, pl.stream));
        mark(1);
        pl.launches += 1;
        const bool scatter = !pl.plan_api || option("sub_scatter", 0) != 0;   // results also in cell order j (what the exports read)
        if (scatter) {
            if (save_q) CU(launch_scatter_columns(sb.d_qdom.p, sb.npad2, sb.d_cell.p, sb.npad2, pl.tp.T, sp.qpitch, pl.d_qdom.p, pl.stream));
            if (save_netp) CU(launch_scatter_columns(sb.d_netp.p, sb.npad2, sb.d_cell.p, sb.npad2, pl.tp.T, sp.qpitch, pl.d_netp.p, pl.stream));
            pl.launches += (save_q ? 1 : 0) + (save_netp ? 1 : 0);
        }
        if (sb.host.pair_cells > 0) {
            SplitArgs b = split_args(pl, save_q, save_netp, &sb.pits);
            b.fuse_export = save_q ? 1 : 0;                                 // route_pair writes the pit cells' series to qdom (cell order) itself
            CU(launch_route_forward(b, false, pl.stream));
            pl.launches += sp.nwin * 2;
            // the pit cells' columns into the engine-order arrays too
            if (save_q) CU(launch_copy_columns(pl.d_qdom.p, sp.qpitch, sb.d_pit_j.p, sb.d_pit_jp.p, sb.npit, pl.tp.T, sb.npad2, sb.d_qdom.p, pl.stream));
        }
        mark(2); mark(3);
        pl.sub_ran = true;
        return 0;
    }
    if (pl.win.on && !tape && pl.nmember == 1 && !pl.ensemble && math_mode() == 1) {
        // tick pass: reservoirs + routing of every cell but the pit pairs; those are routed from rows afterwards
        WindowState &wn = pl.win;
        TkArgs wa{};
        wa.tp = wn.topo; wa.T = pl.tp.T; wa.Tp = sp.Tp; wa.nwin = (pl.tp.T + TK_W - 1) / TK_W;
        wa.nwarp = wn.nwarp; wa.maxu = wn.maxu; wa.wunits = wn.d_wunits.p;
        wa.dt = pl.dt; wa.dx = pl.dx; wa.save_q = save_q ? 1 : 0; wa.save_netp = save_netp ? 1 : 0;
        const size_t nx = (size_t)wa.nwin * npad * TK_W;
        if (wn.d_X.n < nx) TRY(wn.d_X.ensure(nx));
        wa.nb = (int)std::max<long long>(1, option("tick_nb", 2));
        const size_t ncnt = (size_t)((wa.nwin + wa.nb - 1) / wa.nb) * (wn.topo.ntile + wn.topo.nreach);
        if (wn.d_cnt.n < ncnt) TRY(wn.d_cnt.ensure(ncnt));
        wa.cc = wn.d_cc.p; wa.fstates = pl.d_fstates.p; wa.X = wn.d_X.p; wa.rows = sp.d_rows.p; wa.qdom = pl.d_qdom.p;
        wa.netp = pl.d_netp.p; wa.qpitch = sp.qpitch; wa.qsim = pl.d_qsim.p; wa.cnt = wn.d_cnt.p; wa.err = wn.d_err.p;
        if (option("tick_dbg", 0)) {
            static unsigned long long *dp = nullptr;
            if (!dp) CU(cudaMalloc(&dp, 1025 * sizeof(unsigned long long)));
            std::vector<unsigned long long> hv(1025);
            CU(cudaMemcpy(hv.data(), dp, 1025 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            if (hv[1024] != 0 && hv[1024] != ~0ull) {
                fprintf(stderr, "[tick_dbg] previous run: tick -> ms since start:");
                for (int kk = 0; kk < 1024; kk++) if (hv[kk]) fprintf(stderr, " %d:%.3f", kk, (hv[kk] - hv[1024]) * 1e-6);
                fprintf(stderr, "\n");
            }
            CU(cudaMemset(dp, 0, 1024 * sizeof(unsigned long long)));
            CU(cudaMemset(dp + 1024, 0xff, sizeof(unsigned long long)));
            wa.dbg = dp;
        }
        CU(launch_tick_forward(wa, pl.d_fields.p, sp.tm_prcp, sp.tm_pet, pl.stream, wn.variant));
        mark(1);
        pl.launches += 2;
        if (wn.host.npair_cells > 0) {
            SplitArgs b = split_args(pl, save_q, save_netp, &wn.deep);
            b.fuse_export = save_q ? 1 : 0;                                 // route_pair writes the pit cells' series to qdom itself
            CU(launch_route_forward(b, false, pl.stream));
            pl.launches += sp.nwin * 2;
        }
        mark(2); mark(3);
        pl.tick_ran = true;
        return 0;
    }
    if (struct_pass(pl)) {
        if (tape) return fail(SMASH_B200_EUNSUPPORTED, "structure %d: the adjoint is implemented for gr-a and gr-d only", pl.structure);
        CU(launch_vertical_struct(a, sp.tm_prcp, sp.tm_pet, pl.structure, math_mode(), pl.stream));
    } else {
        CU(launch_vertical_forward(a, sp.tm_prcp, sp.tm_pet, math_mode(), tape, pl.stream));
    }
    mark(1);
    // ensembles on a small mesh: lane = member, exact sequential routing; otherwise the per-chain scan
    const bool by_member = pl.ensemble && sp.rg.npair == 0 && a.tp.nrouted <= 12000;   // whatever the size of this launch
    // the routing warps export the routed cells' series themselves, in the shadow of the river walks (option fuse_export)
    a.fuse_export = (save_q && !by_member) ? (int)option("fuse_export", 4) : 0;
    if (by_member) CU(launch_route_members(a, tape, pl.stream));
    else CU(launch_route_forward(a, tape, pl.stream));
    mark(2);
    pl.launches += 1 + sp.nwin * (1 + (sp.rg.npair > 0 ? 1 : 0));   // reservoir pass + per window: chains (+ pit pairs)
    if (save_q && !a.fuse_export) { CU(launch_rows_to_domain(a, pl.stream)); pl.launches++; }
    mark(3);
    return 0;
}

static int split_reverse(SmashPlan &pl) {
    SplitState &sp = pl.sp;
    SplitArgs a = split_args(pl, false, false);
    CU(cudaMemsetAsync(pl.d_grad.p, 0, sizeof(float) * (size_t)pl.nmember * NFIELD * sp.rg.npad, pl.stream));
    auto mark = [&](int i) { pl.kmark[i] = cudaEventRecord(pl.evk[i], pl.stream) == cudaSuccess; };
    mark(4);
    if (sp.ckpt) {
        // windows in reverse order: restore the window's start states, replay it with the tape on (one extra forward sweep in
        // total: recompute factor 2 on the forward side), then its reverse routing sweep and its reverse reservoir sweep.  The
        // adjoint states travel from window to window in gcar (hlr_b) and the gradient planes (hp_b, hft_b, running sums).
        const size_t npad = (size_t)sp.rg.npad, pb = npad * sizeof(float);
        a.Tp = sp.W; a.wnext = sp.d_wnext.p; a.qprev = sp.d_qprev.p;
        for (int w = sp.nwin - 1; w >= 0; w--) {
            const float *ck = sp.d_ckpt.p + (size_t)w * 5 * npad;
            CU(cudaMemcpyAsync(pl.d_fstates.p, ck, 3 * pb, cudaMemcpyDeviceToDevice, pl.stream));
            CU(cudaMemcpyAsync(sp.d_hcar.p, ck + 3 * npad, pb, cudaMemcpyDeviceToDevice, pl.stream));
            CU(cudaMemcpyAsync(sp.d_qprev.p, ck + 4 * npad, pb, cudaMemcpyDeviceToDevice, pl.stream));
            a.t_begin = w * sp.W; a.t_end = std::min(pl.tp.T, (w + 1) * sp.W); a.tape_t0 = w * sp.W;
            const ptrdiff_t off = (ptrdiff_t)w * sp.W;
            a.rows = sp.d_rows_seg.p - off; a.rows_hr = sp.d_rows_hr.p - off; a.rows_w = sp.d_rows_w.p - off;
            CU(launch_vertical_forward(a, sp.tm_prcp, sp.tm_pet, math_mode(), true, pl.stream));
            CU(launch_route_forward_window(a, w, true, pl.stream, true));
            CU(launch_route_adjoint_window(a, w, pl.stream));
            CU(launch_vertical_adjoint(a, sp.tm_prcp, sp.tm_pet, sp.tm_hp, sp.tm_hft, math_mode(), pl.stream));
            CU(launch_first_step(sp.d_rows_w.p, sp.W, (int)npad, sp.d_wnext.p, pl.stream));
            pl.launches += 5 + (sp.rg.npair > 0 ? 1 : 0);
        }
        mark(5); mark(6);
        return 0;
    }
    CU(launch_route_adjoint(a, pl.stream));
    mark(5);
    CU(launch_vertical_adjoint(a, sp.tm_prcp, sp.tm_pet, sp.tm_hp, sp.tm_hft, math_mode(), pl.stream));
    mark(6);
    pl.launches += 1 + sp.nwin;
    return 0;
}

static int plan_members(SmashPlan &pl, int nmember, bool save_q, bool save_netp, bool gradient) {
    if (pl.engine == 1) return split_members(pl, nmember, save_q, save_netp, gradient);
    const Topology &tp = pl.tp;
    const size_t nm = (size_t)nmember;
    TRY(pl.d_fields.ensure(nm * NFIELD * tp.nslots));
    TRY(pl.d_fstates.ensure(nm * 3 * tp.nslots));
    TRY(pl.d_qsim.ensure(std::max<size_t>(1, nm * tp.T * tp.ng)));
    TRY(pl.d_cost_jobs.ensure(nm));
    TRY(pl.d_prog.ensure(nm * tp.nblocks));
    if (pl.need_qdom || save_q || gradient) TRY(pl.d_qdom.ensure(nm * tp.total_ticks * tp.B));
    if (save_netp) TRY(pl.d_netp.ensure(nm * tp.total_ticks * tp.B));
    if (gradient) {
        TRY(pl.d_tape.ensure(nm * tp.total_ticks * 4 * tp.B));
        TRY(pl.d_qsim_b.ensure(std::max<size_t>(1, nm * tp.T * tp.ng)));
        TRY(pl.d_wdom.ensure(nm * tp.total_ticks * tp.B));
        TRY(pl.d_grad.ensure(nm * NFIELD * tp.nslots));
        TRY(pl.d_rprog.ensure(nm * tp.nblocks));
    }
    pl.nmember = nmember;
    return 0;
}

static SolverArgs solver_args(SmashPlan &pl, bool save_q, bool save_netp, bool tape) {
    SolverArgs a{};
    a.tp = pl.dtp; a.nmember = pl.nmember; a.dt = pl.dt; a.dx = pl.dx;
    a.save_q = save_q ? 1 : 0; a.save_netp = save_netp ? 1 : 0; a.tape_on = tape ? 1 : 0;
    a.debug_nowait = (int)option("debug_nowait", 0);
    a.forcing = pl.d_forcing.p; a.fields = pl.d_fields.p; a.fstates = pl.d_fstates.p; a.qsim = pl.d_qsim.p;
    a.qdom = pl.d_qdom.p; a.netp = pl.d_netp.p; a.tape = pl.d_tape.p; a.prog = pl.d_prog.p; a.ticket = pl.d_ticket.p;
    a.qsim_b = pl.d_qsim_b.p; a.wdom = pl.d_wdom.p; a.grad = pl.d_grad.p; a.rprog = pl.d_rprog.p;
    return a;
}

// ---- forcing ------------------------------------------------------------------------------------
static int plan_set_forcing(SmashPlan &pl, const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *in, bool upload = true,
                            bool *fresh = nullptr) {
    if (!in) return fail(SMASH_B200_EINVAL, "input_data is NULL");
    const Topology &tp = pl.tp;
    const bool sparse = setup->sparse_storage != 0;
    const float *prcp = sparse ? in->sparse_prcp : in->prcp;
    const float *pet = sparse ? in->sparse_pet : in->pet;
    if (!prcp || !pet) return fail(SMASH_B200_EINVAL, "input_data.%sprcp / pet is NULL", sparse ? "sparse_" : "");
    if (pl.have_forcing && in->forcing_version != 0 && pl.forcing_version == in->forcing_version && pl.forcing_ptr == prcp &&
        pl.forcing_sparse == (int)sparse) {
        if (fresh) *fresh = false;
        return 0;
    }
    if (fresh) *fresh = true;
    const int64_t stride = sparse ? mesh->nac : (int64_t)mesh->nrow * mesh->ncol;
    const size_t nraw = (size_t)stride * tp.T;
    TRY(pl.d_raw_prcp.ensure(nraw)); TRY(pl.d_raw_pet.ensure(nraw));
    if (pl.engine == 0) TRY(pl.d_forcing.ensure((size_t)tp.total_ticks * 2 * tp.B));
    pin_host(prcp, nraw * sizeof(float)); pin_host(pet, nraw * sizeof(float));
    if (upload) {
        CU(cudaMemcpyAsync(pl.d_raw_prcp.p, prcp, nraw * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
        CU(cudaMemcpyAsync(pl.d_raw_pet.p, pet, nraw * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
    }
    if (pl.engine == 1) {
        // the reservoir pass reads [t][cell j] tiles by TMA: sparse arrays whose order is already j are used in place
        SplitState &sp = pl.sp;
        const float *fp = pl.d_raw_prcp.p, *fe = pl.d_raw_pet.p;
        uint64_t cols = (uint64_t)sp.rg.n, pitch = (uint64_t)stride;
        if (!(sparse && sp.rg.direct && stride % 4 == 0)) {
            const size_t npk = (size_t)sp.rg.npad * tp.T;
            TRY(sp.d_pk_prcp.ensure(npk)); TRY(sp.d_pk_pet.ensure(npk));
            const int32_t *idx = sparse ? pl.d_sparse_k.p : pl.d_cell.p;
            CU(launch_pack_columns(pl.d_raw_prcp.p, stride, idx, sp.rg.n, sp.rg.npad, tp.T, sp.d_pk_prcp.p, pl.stream));
            CU(launch_pack_columns(pl.d_raw_pet.p, stride, idx, sp.rg.n, sp.rg.npad, tp.T, sp.d_pk_pet.p, pl.stream));
            pl.launches += 2;
            fp = sp.d_pk_prcp.p; fe = sp.d_pk_pet.p; cols = pitch = (uint64_t)sp.rg.npad;
        }
        const char *err = nullptr;
        if (make_tensor_map_2d(&sp.tm_prcp, fp, cols, (uint64_t)tp.T, pitch, &err) ||
            make_tensor_map_2d(&sp.tm_pet, fe, cols, (uint64_t)tp.T, pitch, &err))
            return fail(SMASH_B200_ECUDA, "%s", err);
    } else {
        CU(launch_relayout_forcing(pl.dtp, sparse ? pl.d_sparse_k.p : pl.d_cell.p, pl.d_raw_prcp.p, pl.d_raw_pet.p, stride,
                                   pl.d_forcing.p, pl.stream));
        pl.launches++;
    }
    if (pl.engine == 1 && pl.sub.on) {
        // the subtree engine reads the forcing in its own column order: packed once per forcing
        SubState &sb = pl.sub;
        const size_t npk = (size_t)sb.npad2 * tp.T;
        TRY(sb.d_pk_prcp.ensure(npk)); TRY(sb.d_pk_pet.ensure(npk));
        TRY(sb.d_idx_raw.upload(sparse ? sb.idx_sparse : sb.idx_dense, pl.stream));
        CU(launch_pack_columns(pl.d_raw_prcp.p, stride, sb.d_idx_raw.p, sb.npad2, sb.npad2, tp.T, sb.d_pk_prcp.p, pl.stream));
        CU(launch_pack_columns(pl.d_raw_pet.p, stride, sb.d_idx_raw.p, sb.npad2, sb.npad2, tp.T, sb.d_pk_pet.p, pl.stream));
        const char *err = nullptr;
        if (make_tensor_map_2d(&sb.tm_prcp, sb.d_pk_prcp.p, (uint64_t)sb.npad2, (uint64_t)tp.T, (uint64_t)sb.npad2, &err) ||
            make_tensor_map_2d(&sb.tm_pet, sb.d_pk_pet.p, (uint64_t)sb.npad2, (uint64_t)tp.T, (uint64_t)sb.npad2, &err))
            return fail(SMASH_B200_ECUDA, "%s", err);
        sb.have_forcing = true;
    }
    pl.have_forcing = true; pl.forcing_ptr = prcp; pl.forcing_version = in->forcing_version; pl.forcing_sparse = (int)sparse;
    if (mesh->ng > 0 && in->qobs) {
        TRY(pl.d_qobs.ensure((size_t)mesh->ng * tp.T));
        CU(cudaMemcpyAsync(pl.d_qobs.p, in->qobs, (size_t)mesh->ng * tp.T * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
        pl.have_qobs = true;
    }
    return 0;
}

// ---- fields -------------------------------------------------------------------------------------
static const int FIELD_PARAM[4] = {SMASH_P_CP, SMASH_P_CFT, SMASH_P_EXC, SMASH_P_LR};
static const int FIELD_STATE[3] = {SMASH_S_HP, SMASH_S_HFT, SMASH_S_HLR};

// stacked index (1-based, 16 parameters then 8 states) -> device field or -1
static int stacked_to_field(int ind1) {
    const int k = ind1 - 1;
    if (k < SMASH_B200_GNP) { for (int f = 0; f < 4; f++) if (FIELD_PARAM[f] == k) return f; return -1; }
    for (int f = 0; f < 3; f++) if (FIELD_STATE[f] == k - SMASH_B200_GNP) return 4 + f;
    return -1;
}

// planes a structure reads and updates (the reference's STRUCTURE_PARAMETERS / STRUCTURE_STATES, smash/core/_constant.py:20-33)
struct StructurePlanes { bool par[SMASH_B200_GNP]; bool st[SMASH_B200_GNS]; };
static const StructurePlanes &structure_planes(int structure) {
    static StructurePlanes tab[6];
    static bool init = false;
    if (!init) {
        auto set = [&](int s, std::initializer_list<int> p, std::initializer_list<int> h) {
            for (int k : p) tab[s].par[k] = true;
            for (int k : h) tab[s].st[k] = true;
        };
        set(SMASH_STRUCTURE_GR_A, {SMASH_P_CP, SMASH_P_CFT, SMASH_P_EXC, SMASH_P_LR}, {SMASH_S_HP, SMASH_S_HFT, SMASH_S_HLR});
        set(SMASH_STRUCTURE_GR_B, {SMASH_P_CI, SMASH_P_CP, SMASH_P_CFT, SMASH_P_EXC, SMASH_P_LR}, {SMASH_S_HI, SMASH_S_HP, SMASH_S_HFT, SMASH_S_HLR});
        set(SMASH_STRUCTURE_GR_C, {SMASH_P_CI, SMASH_P_CP, SMASH_P_CFT, SMASH_P_CST, SMASH_P_EXC, SMASH_P_LR},
            {SMASH_S_HI, SMASH_S_HP, SMASH_S_HFT, SMASH_S_HST, SMASH_S_HLR});
        set(SMASH_STRUCTURE_GR_D, {SMASH_P_CP, SMASH_P_CFT, SMASH_P_LR}, {SMASH_S_HP, SMASH_S_HFT, SMASH_S_HLR});
        set(SMASH_STRUCTURE_VIC_A, {SMASH_P_B, SMASH_P_CUSL1, SMASH_P_CUSL2, SMASH_P_CLSL, SMASH_P_KS, SMASH_P_DS, SMASH_P_DSM, SMASH_P_WS, SMASH_P_LR},
            {SMASH_S_HUSL1, SMASH_S_HUSL2, SMASH_S_HLSL, SMASH_S_HLR});
        init = true;
    }
    return tab[structure];
}

static int plan_set_fields(SmashPlan &pl, const SmashParameters *par, const SmashStates *st, const float *sample,
                           const int32_t *ind, int nvar, int nmember) {
    const size_t nc = (size_t)pl.ncell;
    const bool other = pl.structure != SMASH_STRUCTURE_GR_A;
    if (other) {
        // every plane the structure reads (STRUCTURE_PARAMETERS / STRUCTURE_STATES of the reference's _constant.py:20-33)
        TRY(pl.d_splanes.ensure((size_t)(SMASH_B200_GNP + SMASH_B200_GNS) * nc));
        const StructurePlanes &need = structure_planes(pl.structure);
        for (int k = 0; k < SMASH_B200_GNP + SMASH_B200_GNS; k++) {
            const float *src = k < SMASH_B200_GNP ? par->v[k] : st->v[k - SMASH_B200_GNP];
            float *dst = pl.d_splanes.p + (size_t)k * nc;
            const bool used = k < SMASH_B200_GNP ? need.par[k] : need.st[k - SMASH_B200_GNP];
            if (src) CU(cudaMemcpyAsync(dst, src, nc * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
            else if (used) return fail(SMASH_B200_EINVAL, "%s plane %d is NULL", k < SMASH_B200_GNP ? "parameters" : "states",
                                       k < SMASH_B200_GNP ? k : k - SMASH_B200_GNP);
            else CU(cudaMemsetAsync(dst, 0, nc * sizeof(float), pl.stream));
        }
    }
    for (int f = 0; f < 4; f++) {
        // the routing passes read lr / hlr from here; the gr-a reservoir planes are placeholders for the other structures
        if (other) { CU(cudaMemcpyAsync(pl.d_planes.p + f * nc, pl.d_splanes.p + (size_t)FIELD_PARAM[f] * nc, nc * sizeof(float), cudaMemcpyDeviceToDevice, pl.stream)); continue; }
        if (!par->v[FIELD_PARAM[f]]) return fail(SMASH_B200_EINVAL, "parameters plane %d is NULL", FIELD_PARAM[f]);
        CU(cudaMemcpyAsync(pl.d_planes.p + f * nc, par->v[FIELD_PARAM[f]], nc * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
    }
    for (int f = 0; f < 3; f++) {
        if (other) { CU(cudaMemcpyAsync(pl.d_planes.p + (4 + f) * nc, pl.d_splanes.p + (size_t)(SMASH_B200_GNP + FIELD_STATE[f]) * nc, nc * sizeof(float), cudaMemcpyDeviceToDevice, pl.stream)); continue; }
        if (!st->v[FIELD_STATE[f]]) return fail(SMASH_B200_EINVAL, "states plane %d is NULL", FIELD_STATE[f]);
        CU(cudaMemcpyAsync(pl.d_planes.p + (4 + f) * nc, st->v[FIELD_STATE[f]], nc * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
    }
    int nv = 0;
    if (sample && nvar > 0) {
        std::vector<int32_t> sf(nvar);
        for (int j = 0; j < nvar; j++) sf[j] = stacked_to_field(ind[j]);  // -1: field unused by gr-a, no effect on the run
        TRY(pl.d_sample_field.upload(sf, pl.stream));
        TRY(pl.d_sample.ensure((size_t)nvar * nmember));
        CU(cudaMemcpyAsync(pl.d_sample.p, sample, (size_t)nvar * nmember * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
        nv = nvar;
    }
    CU(launch_gather_fields(pl.dtp, nmember, pl.d_planes.p, (int64_t)nc, nv ? pl.d_sample.p : nullptr, pl.d_sample_field.p, nv,
                            pl.d_fields.p, pl.stream));
    pl.launches++;
    if (other) {
        if (nv) {
            std::vector<int32_t> spl(nvar);
            for (int j = 0; j < nvar; j++) spl[j] = ind[j] - 1;             // stacked index: 16 parameters then 8 states
            TRY(pl.d_sample_plane.upload(spl, pl.stream));
        }
        CU(launch_gather_struct_fields(pl.d_cell.p, pl.ncols, nmember, pl.d_splanes.p, (int64_t)nc, nv ? pl.d_sample.p : nullptr,
                                       pl.d_sample_plane.p, nv, pl.d_sfields.p, pl.stream));
        pl.launches++;
    }
    return 0;
}

// ---- cost ---------------------------------------------------------------------------------------
static int make_cost_args(SmashPlan &pl, const SmashSetup *setup, const SmashMesh *mesh, float jobs_b, bool adjoint, CostArgs &c,
                          const SmashInputData *in = nullptr) {
    c = CostArgs{};
    c.T = pl.tp.T; c.ng = mesh->ng; c.nmember = pl.nmember; c.start = setup->optimize_start_step - 1;
    if (c.start < 0 || c.start >= c.T) return fail(SMASH_B200_EINVAL, "optimize_start_step %d out of range", setup->optimize_start_step);
    c.dt = setup->dt; c.dx = mesh->dx;
    c.qsim = pl.d_qsim.p; c.qobs = pl.d_qobs.p; c.area = pl.d_area.p; c.wgauge = pl.d_wgauge.p; c.gauge_flwacc = pl.d_gauge_flwacc.p;
    c.njf = setup->njf;
    if (c.njf > 32) return fail(SMASH_B200_EUNSUPPORTED, "more than 32 objective functions");
    for (int j = 0; j < c.njf; j++) {
        c.jobs_fun[j] = setup->jobs_fun[j]; c.wjobs_fun[j] = setup->wjobs_fun[j];
        if (c.jobs_fun[j] < SMASH_JOBS_NSE || c.jobs_fun[j] > SMASH_JOBS_EPF)
            return fail(SMASH_B200_EUNSUPPORTED, "jobs_fun code %d is not implemented", c.jobs_fun[j]);
        if (c.jobs_fun[j] >= SMASH_JOBS_CRC) {
            // signature objectives (mwd_cost.f90:770-970): catchment-mean precipitation and the event mask on the device
            if (adjoint) return fail(SMASH_B200_EUNSUPPORTED, "forward_b with a signature-based objective (jobs_fun code %d)", c.jobs_fun[j]);
            if (!in || !in->mean_prcp || !setup->mask_event)
                return fail(SMASH_B200_EINVAL, "signature-based objective: input_data.mean_prcp / setup.optimize.mask_event missing");
            if (!c.mean_prcp) {
                const size_t nq = (size_t)mesh->ng * c.T;
                TRY(pl.d_mean_prcp.ensure(nq)); TRY(pl.d_mask_event.ensure(nq)); TRY(pl.d_sig_scratch.ensure((size_t)pl.nmember * 2 * nq));
                CU(cudaMemcpyAsync(pl.d_mean_prcp.p, in->mean_prcp, nq * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
                CU(cudaMemcpyAsync(pl.d_mask_event.p, setup->mask_event, nq * sizeof(int32_t), cudaMemcpyHostToDevice, pl.stream));
                c.mean_prcp = pl.d_mean_prcp.p; c.mask_event = pl.d_mask_event.p; c.scratch = pl.d_sig_scratch.p;
            }
        }
    }
    c.jobs_b = jobs_b; c.cost_jobs = pl.d_cost_jobs.p; c.qsim_b = adjoint ? pl.d_qsim_b.p : nullptr;
    return 0;
}

static int run_cost(SmashPlan &pl, const SmashSetup *setup, const SmashMesh *mesh, float jobs_b, bool adjoint,
                    const SmashInputData *in = nullptr) {
    if (mesh->ng <= 0) {
        CU(cudaMemsetAsync(pl.d_cost_jobs.p, 0, sizeof(float) * pl.nmember, pl.stream));
        return 0;
    }
    if (!pl.have_qobs) return fail(SMASH_B200_EINVAL, "input_data.qobs is NULL");
    if (!setup->wgauge) return fail(SMASH_B200_EINVAL, "setup.optimize.wgauge is NULL");
    std::vector<float> wg(setup->wgauge, setup->wgauge + mesh->ng);
    TRY(pl.d_wgauge.upload(wg, pl.stream));
    CostArgs c;
    TRY(make_cost_args(pl, setup, mesh, jobs_b, adjoint, c, in));
    CU(launch_cost(c, pl.stream));
    pl.launches++;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// the O(nrow * ncol) pieces of base_forward / BASE_FORWARD_B: (de)normalisation of the caller's planes, Jreg and its
// adjoint, hyper mapping and its adjoint.  The arithmetic runs on the device (field_kernels.cu); what stays on the host is
// the in-place affine map of the caller's own arrays, which the boundary demands (forward.f90:33-38: parameters and states
// are left denormalised in the caller's memory), spread over a few threads.
// ------------------------------------------------------------------------------------------------
static int download(SmashPlan &pl, void *dst, const void *src, size_t bytes);

static void normalize_planes(float *const *v, int nplanes, size_t nc, const float *lb, const float *ub, bool inverse) {
    auto one = [&](int i) {
        if (!v[i]) return;
        const float l = lb[i], u = ub[i];
        if (!inverse) for (size_t c = 0; c < nc; c++) v[i][c] = (v[i][c] - l) / (u - l);   // mwd_parameters_manipulation.f90:154-179
        else for (size_t c = 0; c < nc; c++) v[i][c] = v[i][c] * (u - l) + l;               // :181-206
    };
    if (nc * (size_t)nplanes < (1u << 20)) { for (int i = 0; i < nplanes; i++) one(i); return; }
    std::vector<std::thread> th;
    const int nth = std::min(nplanes, 8);
    for (int k = 0; k < nth; k++) th.emplace_back([&, k]() { for (int i = k; i < nplanes; i += nth) one(i); });
    for (auto &t : th) t.join();
}

// stacked plane index (16 parameters, then 8 states) of the device field f
static int live_plane(int f) { return f < 4 ? FIELD_PARAM[f] : SMASH_B200_GNP + FIELD_STATE[f - 4]; }

static int ensure_active(SmashPlan &pl, const SmashMesh *mesh) {
    if (pl.have_active) return 0;
    const size_t nc = (size_t)pl.ncell;
    TRY(pl.d_active.ensure(nc));
    CU(cudaMemcpyAsync(pl.d_active.p, mesh->active_cell, nc * sizeof(int32_t), cudaMemcpyHostToDevice, pl.stream));
    pl.have_active = true;
    return 0;
}

// compute_jreg mwd_cost.f90:159-245 on the device; with jreg_b != 0 also COMPUTE_JREG_B (forward_db.f90:2927-3092): the
// adjoint planes (normalised space) of the optimised fields stay in pl.d_jr_b, `planes` lists their stacked indices.
// The caller's planes are denormalised when setup->denormalize_forward is set; the kernels normalise them on the fly
// (compute_cost mwd_cost.f90:284-298).
static int jreg_device(SmashPlan &pl, const SmashSetup *s, const SmashMesh *mesh, const SmashParameters *par, const SmashParameters *par_bgd,
                       const SmashStates *st, const SmashStates *st_bgd, float jreg_b, float *jreg, std::vector<int> &planes) {
    planes.clear();
    for (int i = 0; i < SMASH_B200_GNP; i++) if (s->optim_parameters[i] > 0 && par->v[i] && par_bgd->v[i]) planes.push_back(i);
    for (int i = 0; i < SMASH_B200_GNS; i++) if (s->optim_states[i] > 0 && st->v[i] && st_bgd->v[i]) planes.push_back(SMASH_B200_GNP + i);
    for (int i = 0; i < s->njr; i++)
        if (s->jreg_fun[i] != SMASH_JREG_PRIOR && s->jreg_fun[i] != SMASH_JREG_SMOOTHING && s->jreg_fun[i] != SMASH_JREG_HARD_SMOOTHING)
            return fail(SMASH_B200_EUNSUPPORTED, "jreg_fun code %d is not implemented", s->jreg_fun[i]);
    *jreg = 0.0f;
    if (planes.empty()) return 0;
    const size_t nc = (size_t)pl.ncell, np = planes.size();
    TRY(ensure_active(pl, mesh));
    TRY(pl.d_jr_mat.ensure(np * nc)); TRY(pl.d_jr_bgd.ensure(np * nc)); TRY(pl.d_jr_b.ensure(np * nc)); TRY(pl.d_jreg.ensure(1));
    TRY(pl.d_partial.ensure(std::max<size_t>((size_t)jreg_blocks((int)nc) * np, (size_t)hyper_reduce_blocks(pl.ncols) * NFIELD * (1 + 2 * HYPER_MAX_ND))));
    JregArgs a{};
    a.nrow = mesh->nrow; a.ncol = mesh->ncol; a.ncell = (int)nc; a.nplanes = (int)np; a.normalize = s->denormalize_forward ? 1 : 0;
    a.active = pl.d_active.p; a.mat = pl.d_jr_mat.p; a.bgd = pl.d_jr_bgd.p; a.mat_b = pl.d_jr_b.p;
    for (size_t k = 0; k < np; k++) {
        const int i = planes[k];
        const float *m = i < SMASH_B200_GNP ? par->v[i] : st->v[i - SMASH_B200_GNP];
        const float *b = i < SMASH_B200_GNP ? par_bgd->v[i] : st_bgd->v[i - SMASH_B200_GNP];
        a.lb[k] = i < SMASH_B200_GNP ? s->lb_parameters[i] : s->lb_states[i - SMASH_B200_GNP];
        a.ub[k] = i < SMASH_B200_GNP ? s->ub_parameters[i] : s->ub_states[i - SMASH_B200_GNP];
        CU(cudaMemcpyAsync(pl.d_jr_mat.p + k * nc, m, nc * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
        CU(cudaMemcpyAsync(pl.d_jr_bgd.p + k * nc, b, nc * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
    }
    CU(cudaMemsetAsync(pl.d_jr_b.p, 0, np * nc * sizeof(float), pl.stream));
    CU(cudaMemsetAsync(pl.d_jreg.p, 0, sizeof(float), pl.stream));
    for (int i = 0; i < s->njr; i++) {
        const float w = s->wjreg_fun[i];
        const bool prior = s->jreg_fun[i] == SMASH_JREG_PRIOR;
        const float wt = prior ? w : powf(w, 2.0f);                      // mwd_cost.f90:206-226
        CU(launch_jreg_term(a, s->jreg_fun[i], wt, wt * jreg_b, pl.d_partial.p, pl.d_jreg.p, pl.stream));
        pl.launches += 2;
    }
    TRY(download(pl, jreg, pl.d_jreg.p, sizeof(float)));
    return 0;
}

// descriptors + hyper-parameters on the device; fills the argument block of the hyper kernels
static int hyper_args(SmashPlan &pl, const SmashSetup *s, const SmashMesh *mesh, const SmashInputData *in, const SmashParameters *hp,
                      const SmashStates *hs, HyperArgs &a) {
    const size_t nc = (size_t)pl.ncell;
    if (s->nd > HYPER_MAX_ND) return fail(SMASH_B200_EUNSUPPORTED, "more than %d descriptors", HYPER_MAX_ND);
    (void)mesh;
    if (s->nd > 0 && !(pl.desc_ptr == in->descriptor && in->forcing_version != 0 && pl.desc_version == in->forcing_version)) {
        TRY(pl.d_desc.ensure((size_t)s->nd * nc));
        CU(cudaMemcpyAsync(pl.d_desc.p, in->descriptor, (size_t)s->nd * nc * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
        pl.desc_ptr = in->descriptor; pl.desc_version = in->forcing_version;
    }
    const int nh = s->nhyper;
    std::vector<float> h((size_t)HYPER_NPLANE * nh, 0.0f);
    for (int i = 0; i < SMASH_B200_GNP; i++) if (hp->v[i]) std::copy(hp->v[i], hp->v[i] + nh, h.begin() + (size_t)i * nh);
    for (int i = 0; i < SMASH_B200_GNS; i++) if (hs->v[i]) std::copy(hs->v[i], hs->v[i] + nh, h.begin() + (size_t)(SMASH_B200_GNP + i) * nh);
    TRY(pl.d_hyper.upload(h, pl.stream));
    CU(cudaStreamSynchronize(pl.stream));                                // h is a local
    a = HyperArgs{};
    a.n = pl.ncols; a.npad = pl.ncols; a.ncell = (int)nc; a.nd = s->nd; a.nh = nh; a.poly = s->mapping == SMASH_MAPPING_HYPER_POLYNOMIAL;
    a.cell = pl.d_cell.p; a.desc = pl.d_desc.p; a.hyper = pl.d_hyper.p; a.fields = pl.d_fields.p;
    for (int f = 0; f < NFIELD; f++) a.live[f] = live_plane(f);
    for (int i = 0; i < SMASH_B200_GNP; i++) { a.lb[i] = s->lb_parameters[i]; a.ub[i] = s->ub_parameters[i]; }
    for (int i = 0; i < SMASH_B200_GNS; i++) { a.lb[SMASH_B200_GNP + i] = s->lb_states[i]; a.ub[SMASH_B200_GNP + i] = s->ub_states[i]; }
    return 0;
}

// hyper_parameters_to_parameters / hyper_states_to_states (mwd_parameters_manipulation.f90:304-362,
// mwd_states_manipulation.f90:271-329): the plan's field planes are written directly; the caller's 16 + 8 rectangles,
// which the reference rewrites as a side effect, are filled from one device pass over the rectangle.
static int hyper_apply(SmashPlan &pl, const HyperArgs &a, SmashParameters *par, SmashStates *st) {
    const size_t nc = (size_t)pl.ncell;
    CU(launch_hyper_fields(a, pl.stream));
    TRY(pl.d_rect.ensure((size_t)HYPER_NPLANE * nc));
    CU(launch_hyper_rect(a, pl.d_rect.p, pl.stream));
    pl.launches += 2;
    for (int i = 0; i < SMASH_B200_GNP; i++) if (par->v[i]) TRY(download(pl, par->v[i], pl.d_rect.p + (size_t)i * nc, nc * sizeof(float)));
    for (int i = 0; i < SMASH_B200_GNS; i++) if (st->v[i]) TRY(download(pl, st->v[i], pl.d_rect.p + (size_t)(SMASH_B200_GNP + i) * nc, nc * sizeof(float)));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// plan cache
// ------------------------------------------------------------------------------------------------
static int get_plan(const SmashSetup *setup, const SmashMesh *mesh, SmashPlan **out, int engine = -1) {
    if (!setup || !mesh) return fail(SMASH_B200_EINVAL, "setup / mesh is NULL");
    TRY(check_device());
    const uint64_t h = hash_mesh(mesh->nrow, mesh->ncol, mesh->ng, mesh->flwdir, mesh->flwacc, mesh->active_cell,
                                 mesh->local_active_cell, mesh->path, mesh->gauge_pos);
    int dev = 0;
    cudaGetDevice(&dev);
    // large sparse forcing that lives on the host: plans of the ABI entry points use 256-step windows so that a forward run
    // can stream (forward_streamed); option "stream" 0 turns it off, "stream_min_mb" is the forcing size it starts at
    const bool small = option("stream", 1) != 0 && setup->sparse_storage &&
                       (size_t)mesh->nac * setup->ntime_step * 8 >= ((size_t)option("stream_min_mb", 256) << 20);
    char key[192];
    snprintf(key, sizeof key, "%016llx:%d:%d:%d:%g:%g:%lld:%d:%d:%d:%lld", (unsigned long long)h, dev, setup->structure, setup->ntime_step,
             (double)setup->dt, (double)mesh->dx, option("block", 0), pick_engine(engine), setup->sparse_storage ? 1 : 0, small ? 1 : 0,
             option("route_ded_min", 96) + 1024 * option("route_ded_max", 64) + (option("route_queues", 12) << 20) +
                 ((option("adjoint_checkpoint", -1) + 1) << 26) + (option("tape_budget_mb", 16384) << 28) +
                 (option("tick_pass", 0) << 48) + (option("tick_slack", 1) << 49) + (option("shallow_acc", 32) << 50) + ((option("sub_engine", -1) + 1) << 58));
    auto it = g_plans.find(key);
    if (it == g_plans.end()) {
        std::unique_ptr<SmashPlan> pl(new SmashPlan());
        pl->small_windows = small;
        TRY(plan_build(*pl, setup, mesh, 1, engine));
        it = g_plans.emplace(key, std::move(pl)).first;
    }
    *out = it->second.get();
    return 0;
}

static int download(SmashPlan &pl, void *dst, const void *src, size_t bytes) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, pl.stream));
    return 0;
}

// qsim_domain / net_prcp_domain in the reference's layout
static int export_domain(SmashPlan &pl, const SmashSetup *setup, const SmashMesh *mesh, const float *skewed, float *dense, float *sparse) {
    const Topology &tp = pl.tp;
    if (pl.engine == 1) {   // [t][qpitch] in cell order j
        SplitState &sp = pl.sp;
        if (setup->sparse_storage) {
            if (!sparse) return 0;
            const size_t n = (size_t)mesh->nac * tp.T;
            pin_host(sparse, n * sizeof(float));
            if (sp.rg.direct) { CU(cudaMemcpyAsync(sparse, skewed, n * sizeof(float), cudaMemcpyDeviceToHost, pl.stream)); return 0; }
            TRY(pl.d_out.ensure(n));
            CU(cudaMemsetAsync(pl.d_out.p, 0, n * sizeof(float), pl.stream));
            CU(launch_scatter_columns(skewed, sp.qpitch, pl.d_sparse_k.p, sp.rg.n, tp.T, mesh->nac, pl.d_out.p, pl.stream));
            pl.launches++;
            CU(cudaMemcpyAsync(sparse, pl.d_out.p, n * sizeof(float), cudaMemcpyDeviceToHost, pl.stream));
        } else {
            if (!dense) return 0;
            const size_t n = (size_t)pl.ncell * tp.T;
            TRY(pl.d_out.ensure(n));
            CU(cudaMemcpyAsync(pl.d_out.p, dense, n * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
            CU(launch_scatter_columns(skewed, sp.qpitch, pl.d_cell.p, sp.rg.n, tp.T, pl.ncell, pl.d_out.p, pl.stream));
            pl.launches++;
            CU(cudaMemcpyAsync(dense, pl.d_out.p, n * sizeof(float), cudaMemcpyDeviceToHost, pl.stream));
        }
        return 0;
    }
    if (setup->sparse_storage) {
        if (!sparse) return 0;
        const size_t n = (size_t)mesh->nac * tp.T;
        TRY(pl.d_out.ensure(n));
        CU(cudaMemsetAsync(pl.d_out.p, 0, n * sizeof(float), pl.stream));
        CU(launch_unskew(pl.dtp, pl.d_sparse_k.p, skewed, mesh->nac, 0.0f, pl.d_out.p, pl.stream));
        pl.launches++;
        CU(cudaMemcpyAsync(sparse, pl.d_out.p, n * sizeof(float), cudaMemcpyDeviceToHost, pl.stream));
    } else {
        if (!dense) return 0;
        // only active cells are written by the reference (md_forward_structure.f90:164-194): read-modify-write
        const size_t n = (size_t)pl.ncell * tp.T;
        TRY(pl.d_out.ensure(n));
        CU(cudaMemcpyAsync(pl.d_out.p, dense, n * sizeof(float), cudaMemcpyHostToDevice, pl.stream));
        CU(launch_unskew(pl.dtp, pl.d_cell.p, skewed, pl.ncell, 0.0f, pl.d_out.p, pl.stream));
        pl.launches++;
        CU(cudaMemcpyAsync(dense, pl.d_out.p, n * sizeof(float), cudaMemcpyDeviceToHost, pl.stream));
    }
    return 0;
}

static void scatter_sorted(const SmashPlan &pl, const float *sorted, float *plane) {
    for (int s = 0; s < pl.ncols; s++) if (pl.col_cell[s] >= 0) plane[pl.col_cell[s]] = sorted[s];
}

// Streamed forward run of the split engine.  The sparse forcing arrays are used in place on the device ([t][k], k = j), so
// a time window of them is one contiguous piece of the caller's arrays, and so is a window of sparse_qsim_domain.  Window w
// is uploaded on s_in while window w - 1 is computed on the plan's stream and window w - 2 travels back on s_out: both PCIe
// directions and the kernels overlap, and the run costs about as long as the larger of the two transfers.
static bool can_stream(const SmashPlan &pl, const SmashSetup *setup) {
    return pl.engine == 1 && pl.structure == SMASH_STRUCTURE_GR_A && pl.small_windows && pl.sp.nwin > 1 && setup->sparse_storage && pl.sp.rg.direct && pl.tp.nactive % 4 == 0 &&
           option("stream", 1) != 0;
}
static int forward_streamed(SmashPlan &pl, const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *in, SmashOutput *out,
                            bool save_q, bool save_netp) {
    SplitState &sp = pl.sp;
    const Topology &tp = pl.tp;
    const size_t nac = (size_t)mesh->nac, npad = (size_t)sp.rg.npad;
    bool fresh = true;
    TRY(plan_set_forcing(pl, setup, mesh, in, false, &fresh));           // buffers, tensor maps, qobs; no forcing copy
    if (!pl.s_in) { CU(cudaStreamCreateWithFlags(&pl.s_in, cudaStreamNonBlocking)); CU(cudaStreamCreateWithFlags(&pl.s_out, cudaStreamNonBlocking)); }
    while ((int)pl.ev_in.size() < sp.nwin) {
        cudaEvent_t e0, e1;
        CU(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming)); pl.ev_in.push_back(e0);
        CU(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming)); pl.ev_cmp.push_back(e1);
    }
    float *hq = save_q ? out->sparse_qsim_domain : nullptr, *hn = save_netp ? out->sparse_net_prcp_domain : nullptr;
    if (hq) pin_host(hq, nac * tp.T * sizeof(float));
    if (hn) pin_host(hn, nac * tp.T * sizeof(float));
    SplitArgs a = split_args(pl, save_q, save_netp);
    for (int i = 0; i < 8; i++) pl.kmark[i] = 0;
    // the fields are uploaded on the plan's stream: nothing of this run may start before them
    CU(cudaMemcpy2DAsync(sp.d_hcar.p, npad * sizeof(float), pl.d_fields.p + (size_t)F_HLR * npad, (size_t)NFIELD * npad * sizeof(float),
                         npad * sizeof(float), 1, cudaMemcpyDeviceToDevice, pl.stream));
    for (int w = 0; w < sp.nwin; w++) {
        const int t0 = w * sp.W, t1 = std::min(tp.T, (w + 1) * sp.W);
        const size_t off = (size_t)t0 * nac, cnt = (size_t)(t1 - t0) * nac;
        if (fresh) {
            CU(cudaMemcpyAsync(pl.d_raw_prcp.p + off, in->sparse_prcp + off, cnt * sizeof(float), cudaMemcpyHostToDevice, pl.s_in));
            CU(cudaMemcpyAsync(pl.d_raw_pet.p + off, in->sparse_pet + off, cnt * sizeof(float), cudaMemcpyHostToDevice, pl.s_in));
            CU(cudaEventRecord(pl.ev_in[w], pl.s_in));
            CU(cudaStreamWaitEvent(pl.stream, pl.ev_in[w], 0));
        }
        a.t_begin = t0; a.t_end = t1;
        CU(launch_vertical_forward(a, sp.tm_prcp, sp.tm_pet, math_mode(), false, pl.stream));
        a.fuse_export = save_q ? (int)option("fuse_export", 4) : 0;
        CU(launch_route_forward_window(a, w, false, pl.stream));
        pl.launches += 2 + (sp.rg.npair > 0 ? 1 : 0);
        if (save_q && !a.fuse_export) { CU(launch_rows_to_domain(a, pl.stream)); pl.launches++; }
        CU(cudaEventRecord(pl.ev_cmp[w], pl.stream));
        if (hq || hn) CU(cudaStreamWaitEvent(pl.s_out, pl.ev_cmp[w], 0));
        if (hq) CU(cudaMemcpyAsync(hq + off, pl.d_qdom.p + (size_t)t0 * sp.qpitch, cnt * sizeof(float), cudaMemcpyDeviceToHost, pl.s_out));
        if (hn) CU(cudaMemcpyAsync(hn + off, pl.d_netp.p + (size_t)t0 * sp.qpitch, cnt * sizeof(float), cudaMemcpyDeviceToHost, pl.s_out));
    }
    return 0;
}

// after a synchronisation of the plan's stream: a wait of the tick pass that did not end leaves 1 + unit in its error word
static int tick_check(SmashPlan &pl) {
    if (pl.sub_ran) {
        pl.sub_ran = false;
        int h = 0;
        CU(cudaMemcpyAsync(&h, pl.sub.d_err.p, sizeof(int), cudaMemcpyDeviceToHost, pl.stream));
        CU(cudaStreamSynchronize(pl.stream));
        if (h != 0) return fail(SMASH_B200_ECUDA, "subtree engine: tile %d waited for an inflow block that never came (results invalid)", h - 1);
    }
    if (!pl.tick_ran) return 0;
    pl.tick_ran = false;
    int h = 0;
    CU(cudaMemcpyAsync(&h, pl.win.d_err.p, sizeof(int), cudaMemcpyDeviceToHost, pl.stream));
    CU(cudaStreamSynchronize(pl.stream));
    if (h != 0) return fail(SMASH_B200_ECUDA, "tick pass: a wait for unit %d did not end (results invalid)", h - 1);
    return 0;
}

static int run_forward_engine(SmashPlan &pl, bool save_q, bool save_netp, bool tape) {
    pl.tick_ran = false; pl.sub_ran = false;
    if (pl.engine == 1) return split_forward(pl, save_q, save_netp, tape);
    SolverArgs a = solver_args(pl, save_q, save_netp, tape);
    CU(launch_forward(a, math_mode(), pl.stream));
    pl.launches++;
    return 0;
}
static int run_reverse_engine(SmashPlan &pl) {
    if (pl.engine == 1) return split_reverse(pl);
    SolverArgs a = solver_args(pl, false, false, true);
    CU(launch_reverse(a, math_mode(), pl.stream));
    pl.launches++;
    return 0;
}

static int math_mode() { return (int)option("math", 1); }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// fields_ready: the plan's field planes were written on the device already (hyper mapping), nothing is taken from par / st
static int forward_common(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *in, SmashParameters *par,
                          SmashStates *st, SmashOutput *out, SmashPlan **plan_out, float *jobs_out, bool restore_states,
                          bool fields_ready = false) {
    SmashPlan *pl;
    TRY(get_plan(setup, mesh, &pl));
    *plan_out = pl;
    pl->launches = 0;
    pl->ensemble = false;
    const Topology &tp = pl->tp;
    const bool save_q = setup->save_qsim_domain && out && (setup->sparse_storage ? out->sparse_qsim_domain : out->qsim_domain);
    const bool save_n = setup->save_net_prcp_domain && out && (setup->sparse_storage ? out->sparse_net_prcp_domain : out->net_prcp_domain);
    TRY(plan_members(*pl, 1, save_q, save_n, false));
    const bool streamed = can_stream(*pl, setup) && in && in->sparse_prcp && in->sparse_pet;
    if (streamed) {
        if (!fields_ready) TRY(plan_set_fields(*pl, par, st, nullptr, nullptr, 0, 1));
        TRY(forward_streamed(*pl, setup, mesh, in, out, save_q, save_n));
    } else {
        TRY(plan_set_forcing(*pl, setup, mesh, in));
        if (!fields_ready) TRY(plan_set_fields(*pl, par, st, nullptr, nullptr, 0, 1));
        TRY(run_forward_engine(*pl, save_q, save_n, false));
    }
    TRY(run_cost(*pl, setup, mesh, 0.0f, false, in));
    std::vector<float> fs((size_t)3 * pl->ncols), sfs;
    float jobs = 0.0f;
    TRY(download(*pl, fs.data(), pl->d_fstates.p, fs.size() * sizeof(float)));
    const bool other = struct_pass(*pl);
    if (other) {
        sfs.resize((size_t)SMASH_B200_GNS * pl->ncols);
        TRY(download(*pl, sfs.data(), pl->d_sfstates.p, sfs.size() * sizeof(float)));
    }
    TRY(download(*pl, &jobs, pl->d_cost_jobs.p, sizeof(float)));
    if (out && out->qsim && mesh->ng > 0) TRY(download(*pl, out->qsim, pl->d_qsim.p, (size_t)mesh->ng * tp.T * sizeof(float)));
    if (save_q && !streamed) TRY(export_domain(*pl, setup, mesh, pl->d_qdom.p, out->qsim_domain, out->sparse_qsim_domain));
    if (save_n && !streamed) TRY(export_domain(*pl, setup, mesh, pl->d_netp.p, out->net_prcp_domain, out->sparse_net_prcp_domain));
    const size_t nc = (size_t)pl->ncell;
    // output%fstates = states (forward.f90:71): every plane is copied, the three prognostic ones hold final values.  The
    // host copies run while the device works.
    if (out)
        for (int i = 0; i < SMASH_B200_GNS; i++)
            if (out->fstates.v[i] && st->v[i]) memcpy(out->fstates.v[i], st->v[i], nc * sizeof(float));
    CU(cudaStreamSynchronize(pl->stream));
    if (streamed) CU(cudaStreamSynchronize(pl->s_out));
    TRY(tick_check(*pl));
    const StructurePlanes &used = structure_planes(pl->structure);
    for (int k = 0; k < SMASH_B200_GNS; k++) {
        if (!used.st[k]) continue;
        // gr-a: hp, hft, hlr from the engine's three planes; the other structures: hlr from there, the rest from their own block
        const float *src = nullptr;
        if (other && k != SMASH_S_HLR) src = sfs.data() + (size_t)k * pl->ncols;
        else for (int f = 0; f < 3; f++) if (FIELD_STATE[f] == k) src = fs.data() + (size_t)f * pl->ncols;
        float *dst = restore_states ? (out ? out->fstates.v[k] : nullptr) : st->v[k];
        if (dst) scatter_sorted(*pl, src, dst);
        if (!restore_states && out && out->fstates.v[k]) memcpy(out->fstates.v[k], st->v[k], nc * sizeof(float));
    }
    *jobs_out = jobs;
    return 0;
}

extern "C" int smash_b200_forward(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *in, SmashParameters *par,
                                  const SmashParameters *par_bgd, SmashStates *st, const SmashStates *st_bgd, SmashOutput *out,
                                  float *cost) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!setup || !mesh || !par || !st) return fail(SMASH_B200_EINVAL, "NULL argument");
    const size_t nc = (size_t)mesh->nrow * mesh->ncol;
    if (setup->denormalize_forward) {                                   // forward.f90:33-38
        normalize_planes(par->v, SMASH_B200_GNP, nc, setup->lb_parameters, setup->ub_parameters, true);
        normalize_planes(st->v, SMASH_B200_GNS, nc, setup->lb_states, setup->ub_states, true);
    }
    SmashPlan *pl;
    float jobs = 0.0f;
    TRY(forward_common(setup, mesh, in, par, st, out, &pl, &jobs, true));
    // compute_cost mwd_cost.f90:247-306: Jreg on normalised values (the kernels normalise the planes on the fly)
    float jreg = 0.0f;
    if (setup->njr > 0) {
        if (!par_bgd || !st_bgd) return fail(SMASH_B200_EINVAL, "parameters_bgd / states_bgd required by the regularisation term");
        std::vector<int> planes;
        TRY(jreg_device(*pl, setup, mesh, par, par_bgd, st, st_bgd, 0.0f, &jreg, planes));
        CU(cudaStreamSynchronize(pl->stream));
    }
    if (setup->denormalize_forward) {
        // the reference normalises the caller's planes around Jreg and denormalises them again (mwd_cost.f90:284-303): the
        // float32 round trip is a visible side effect on the caller's arrays, so it is reproduced
        normalize_planes(par->v, SMASH_B200_GNP, nc, setup->lb_parameters, setup->ub_parameters, false);
        normalize_planes(st->v, SMASH_B200_GNS, nc, setup->lb_states, setup->ub_states, false);
        normalize_planes(par->v, SMASH_B200_GNP, nc, setup->lb_parameters, setup->ub_parameters, true);
        normalize_planes(st->v, SMASH_B200_GNS, nc, setup->lb_states, setup->ub_states, true);
    }
    const float c = jobs + setup->wjreg * jreg;
    if (out) { out->cost = c; out->cost_jobs = jobs; out->cost_jreg = jreg; }
    if (cost) *cost = c;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// forward_b
// ------------------------------------------------------------------------------------------------
static int gradient_common(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *in, SmashParameters *par,
                           SmashStates *st, SmashOutput *out, float cost_b, SmashPlan **plan_out, float *jobs_out,
                           bool fields_ready = false) {
    SmashPlan *pl;
    TRY(get_plan(setup, mesh, &pl));
    *plan_out = pl;
    pl->launches = 0;
    pl->ensemble = false;
    const Topology &tp = pl->tp;
    TRY(plan_members(*pl, 1, false, false, true));
    TRY(plan_set_forcing(*pl, setup, mesh, in));
    if (!fields_ready) TRY(plan_set_fields(*pl, par, st, nullptr, nullptr, 0, 1));
    TRY(run_forward_engine(*pl, false, false, true));
    if (mesh->ng > 0) TRY(run_cost(*pl, setup, mesh, cost_b, true));
    else CU(cudaMemsetAsync(pl->d_cost_jobs.p, 0, sizeof(float), pl->stream));
    TRY(run_reverse_engine(*pl));                                       // the gradient planes stay on the device (pl->d_grad)
    float jobs = 0.0f;
    TRY(download(*pl, &jobs, pl->d_cost_jobs.p, sizeof(float)));
    if (out && out->qsim && mesh->ng > 0) TRY(download(*pl, out->qsim, pl->d_qsim.p, (size_t)mesh->ng * tp.T * sizeof(float)));
    CU(cudaStreamSynchronize(pl->stream));
    *jobs_out = jobs;
    return 0;
}

extern "C" int smash_b200_forward_b(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *in, SmashParameters *par,
                                    SmashParameters *par_b, const SmashParameters *par_bgd, SmashStates *st, SmashStates *st_b,
                                    const SmashStates *st_bgd, SmashOutput *out, float *cost, float *cost_b) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!setup || !mesh || !par || !st || !par_b || !st_b) return fail(SMASH_B200_EINVAL, "NULL argument");
    const size_t nc = (size_t)mesh->nrow * mesh->ncol;
    const float seed = cost_b ? *cost_b : 1.0f;
    if (setup->denormalize_forward) {                                   // forward_db.f90:10697-10703
        normalize_planes(par->v, SMASH_B200_GNP, nc, setup->lb_parameters, setup->ub_parameters, true);
        normalize_planes(st->v, SMASH_B200_GNS, nc, setup->lb_states, setup->ub_states, true);
    }
    SmashPlan *pl;
    float jobs = 0.0f;
    TRY(gradient_common(setup, mesh, in, par, st, out, seed, &pl, &jobs));
    for (int i = 0; i < SMASH_B200_GNP; i++) if (par_b->v[i]) memset(par_b->v[i], 0, nc * sizeof(float));   // :10869
    for (int i = 0; i < SMASH_B200_GNS; i++) if (st_b->v[i]) memset(st_b->v[i], 0, nc * sizeof(float));     // :10870
    // the seven live gradient planes are assembled on the device in rectangle layout: Jreg adjoint (normalised space,
    // COMPUTE_COST_B :3252-3353) -> NORMALIZE_*_B -> + GR_A_FORWARD_B (:10885) -> DENORMALIZE_*_B (:10931-10935)
    TRY(pl->d_rectb.ensure((size_t)NFIELD * nc));
    CU(cudaMemsetAsync(pl->d_rectb.p, 0, (size_t)NFIELD * nc * sizeof(float), pl->stream));
    float jreg = 0.0f;
    if (setup->njr > 0) {
        if (!par_bgd || !st_bgd) return fail(SMASH_B200_EINVAL, "parameters_bgd / states_bgd required by the regularisation term");
        std::vector<int> planes;
        TRY(jreg_device(*pl, setup, mesh, par, par_bgd, st, st_bgd, setup->wjreg * seed, &jreg, planes));
        for (size_t k = 0; k < planes.size(); k++) {
            int f = -1;
            for (int q = 0; q < NFIELD; q++) if (live_plane(q) == planes[k]) f = q;
            if (f >= 0) {
                CU(cudaMemcpyAsync(pl->d_rectb.p + (size_t)f * nc, pl->d_jr_b.p + k * nc, nc * sizeof(float), cudaMemcpyDeviceToDevice, pl->stream));
            } else {
                // an optimised plane the structure does not use: its gradient is the Jreg adjoint alone
                const int i = planes[k];
                float *dst = i < SMASH_B200_GNP ? par_b->v[i] : st_b->v[i - SMASH_B200_GNP];
                if (!dst) continue;
                TRY(download(*pl, dst, pl->d_jr_b.p + k * nc, nc * sizeof(float)));
                CU(cudaStreamSynchronize(pl->stream));
                if (setup->denormalize_forward) {
                    const float span = i < SMASH_B200_GNP ? setup->ub_parameters[i] - setup->lb_parameters[i]
                                                          : setup->ub_states[i - SMASH_B200_GNP] - setup->lb_states[i - SMASH_B200_GNP];
                    for (size_t c = 0; c < nc; c++) dst[c] = span * (dst[c] / span);
                }
            }
        }
    }
    GradScale sc{};
    sc.on = setup->denormalize_forward ? 1 : 0;
    for (int f = 0; f < NFIELD; f++)
        sc.span[f] = f < 4 ? setup->ub_parameters[FIELD_PARAM[f]] - setup->lb_parameters[FIELD_PARAM[f]]
                           : setup->ub_states[FIELD_STATE[f - 4]] - setup->lb_states[FIELD_STATE[f - 4]];
    CU(launch_scatter_grad(pl->d_grad.p, pl->d_cell.p, pl->ncols, pl->ncols, (int)nc, sc, pl->d_rectb.p, pl->stream));
    pl->launches++;
    for (int f = 0; f < NFIELD; f++) {
        float *dst = f < 4 ? par_b->v[FIELD_PARAM[f]] : st_b->v[FIELD_STATE[f - 4]];
        if (dst) TRY(download(*pl, dst, pl->d_rectb.p + (size_t)f * nc, nc * sizeof(float)));
    }
    CU(cudaStreamSynchronize(pl->stream));
    const float c = jobs + setup->wjreg * jreg;
    if (out) { out->cost = c; out->cost_jobs = jobs; out->cost_jreg = jreg; }
    if (cost) *cost = c;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// hyper_forward / hyper_forward_b
// ------------------------------------------------------------------------------------------------
static int hyper_check(const SmashSetup *setup, const SmashInputData *in) {
    if (setup->structure != SMASH_STRUCTURE_GR_A)
        return fail(SMASH_B200_EUNSUPPORTED, "structure %d: the descriptor mappings are implemented for gr-a only", setup->structure);
    if (setup->mapping != SMASH_MAPPING_HYPER_LINEAR && setup->mapping != SMASH_MAPPING_HYPER_POLYNOMIAL)
        return fail(SMASH_B200_EINVAL, "hyper_forward needs mapping hyper-linear or hyper-polynomial");
    const int want = setup->mapping == SMASH_MAPPING_HYPER_LINEAR ? 1 + setup->nd : 1 + 2 * setup->nd;
    if (setup->nhyper != want) return fail(SMASH_B200_EINVAL, "nhyper = %d, expected %d", setup->nhyper, want);
    if (setup->nd > 0 && (!in || !in->descriptor)) return fail(SMASH_B200_EINVAL, "input_data.descriptor is NULL");
    return 0;
}

extern "C" int smash_b200_hyper_forward(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *in,
                                        SmashParameters *par, const SmashParameters *hyper_par, const SmashParameters *hyper_par_bgd,
                                        SmashStates *st, const SmashStates *hyper_st, const SmashStates *hyper_st_bgd,
                                        SmashOutput *out, float *cost) {
    (void)hyper_par_bgd; (void)hyper_st_bgd;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!setup || !mesh || !par || !st || !hyper_par || !hyper_st) return fail(SMASH_B200_EINVAL, "NULL argument");
    TRY(hyper_check(setup, in));
    SmashPlan *pl;
    TRY(get_plan(setup, mesh, &pl));
    const bool save_q = setup->save_qsim_domain && out && (setup->sparse_storage ? out->sparse_qsim_domain : out->qsim_domain);
    const bool save_n = setup->save_net_prcp_domain && out && (setup->sparse_storage ? out->sparse_net_prcp_domain : out->net_prcp_domain);
    TRY(plan_members(*pl, 1, save_q, save_n, false));
    HyperArgs ha;
    TRY(hyper_args(*pl, setup, mesh, in, hyper_par, hyper_st, ha));
    TRY(hyper_apply(*pl, ha, par, st));
    float jobs = 0.0f;
    TRY(forward_common(setup, mesh, in, par, st, out, &pl, &jobs, false, true));   // states keep their final values (forward.f90:145)
    const float c = jobs + setup->wjreg * 0.0f;                          // hyper_compute_cost mwd_cost.f90:309-348
    if (out) { out->cost = c; out->cost_jobs = jobs; }
    if (cost) *cost = c;
    return 0;
}

extern "C" int smash_b200_hyper_forward_b(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *in,
                                          SmashParameters *par, const SmashParameters *hyper_par, SmashParameters *hyper_par_b,
                                          SmashStates *st, const SmashStates *hyper_st, SmashStates *hyper_st_b, SmashOutput *out,
                                          float *cost, float *cost_b) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!setup || !mesh || !par || !st || !hyper_par || !hyper_st || !hyper_par_b || !hyper_st_b)
        return fail(SMASH_B200_EINVAL, "NULL argument");
    TRY(hyper_check(setup, in));
    const float seed = cost_b ? *cost_b : 1.0f;
    SmashPlan *pl;
    TRY(get_plan(setup, mesh, &pl));
    TRY(plan_members(*pl, 1, false, false, true));
    HyperArgs ha;
    TRY(hyper_args(*pl, setup, mesh, in, hyper_par, hyper_st, ha));
    TRY(hyper_apply(*pl, ha, par, st));
    float jobs = 0.0f;
    TRY(gradient_common(setup, mesh, in, par, st, out, seed, &pl, &jobs, true));
    // HYPER_*_B forward_db.f90:1434-1537, 2272-2369: reductions of the seven live gradient planes; every other plane has a
    // zero adjoint (beta / alpha are not even read, :1489-1490)
    const int nh = setup->nhyper;
    TRY(pl->d_hyper_b.ensure((size_t)NFIELD * nh));
    TRY(pl->d_partial.ensure((size_t)hyper_reduce_blocks(pl->ncols) * NFIELD * (1 + 2 * HYPER_MAX_ND)));
    CU(launch_hyper_reduce(ha, pl->d_grad.p, pl->d_partial.p, pl->d_hyper_b.p, pl->stream));
    pl->launches += 2;
    std::vector<float> hb((size_t)NFIELD * nh);
    TRY(download(*pl, hb.data(), pl->d_hyper_b.p, hb.size() * sizeof(float)));
    CU(cudaStreamSynchronize(pl->stream));
    for (int i = 0; i < SMASH_B200_GNP; i++) if (hyper_par_b->v[i]) for (int k = 0; k < nh; k++) hyper_par_b->v[i][k] = 0.0f;
    for (int i = 0; i < SMASH_B200_GNS; i++) if (hyper_st_b->v[i]) for (int k = 0; k < nh; k++) hyper_st_b->v[i][k] = 0.0f;
    for (int f = 0; f < NFIELD; f++) {
        float *dst = f < 4 ? hyper_par_b->v[FIELD_PARAM[f]] : hyper_st_b->v[FIELD_STATE[f - 4]];
        const float *src_h = f < 4 ? hyper_par->v[FIELD_PARAM[f]] : hyper_st->v[FIELD_STATE[f - 4]];
        if (dst && src_h) std::copy(hb.begin() + (size_t)f * nh, hb.begin() + (size_t)(f + 1) * nh, dst);
    }
    if (out) { out->cost = jobs; out->cost_jobs = jobs; }
    if (cost) *cost = jobs;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// compute_multiple_run
// ------------------------------------------------------------------------------------------------
extern "C" int smash_b200_compute_multiple_run(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *in,
                                               const SmashParameters *par, const SmashStates *st, SmashOutput *out,
                                               const float *sample, const int32_t *ind, int32_t nvar, int32_t ns,
                                               float *res_cost, float *res_qsim) {
    (void)out;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!setup || !mesh || !par || !st || !sample || !ind || !res_cost) return fail(SMASH_B200_EINVAL, "NULL argument");
    if (ns <= 0) return 0;
    if (setup->denormalize_forward) return fail(SMASH_B200_EUNSUPPORTED, "compute_multiple_run with denormalize_forward");
    for (int j = 0; j < nvar; j++)
        if (ind[j] < 1 || ind[j] > SMASH_B200_GNP + SMASH_B200_GNS) return fail(SMASH_B200_EINVAL, "ind_parameters_states[%d] = %d out of range", j, ind[j]);
    SmashPlan *pl;
    TRY(get_plan(setup, mesh, &pl, (int)option("ensemble_engine", -1)));
    pl->ensemble = true;
    pl->launches = 0;
    const Topology &tp = pl->tp;
    // members per launch bounded by a memory budget
    const size_t per_member = ((size_t)NFIELD + 3) * pl->ncols * 4 + (size_t)tp.T * tp.ng * 4 + 64 +
                              (pl->engine == 1 ? (size_t)pl->ncols * (pl->sp.Tp + 1) * 4 + (size_t)pl->sp.rg.ntask * 4
                                               : (pl->need_qdom ? (size_t)tp.total_ticks * tp.B * 4 : 0));
    const size_t budget = (size_t)option("member_budget_mb", 16384) << 20;
    int chunk = (int)std::min<size_t>((size_t)ns, std::max<size_t>(1, budget / per_member));
    TRY(plan_set_forcing(*pl, setup, mesh, in));
    const size_t nq = (size_t)mesh->ng * tp.T;
    std::vector<float> jobs(chunk);
    for (int m0 = 0; m0 < ns; m0 += chunk) {
        const int nm = std::min(chunk, ns - m0);
        TRY(plan_members(*pl, nm, false, false, false));
        TRY(plan_set_fields(*pl, par, st, sample + (size_t)m0 * nvar, ind, nvar, nm));
        TRY(run_forward_engine(*pl, false, false, false));
        TRY(run_cost(*pl, setup, mesh, 0.0f, false, in));
        TRY(download(*pl, jobs.data(), pl->d_cost_jobs.p, (size_t)nm * sizeof(float)));
        if (res_qsim && nq) TRY(download(*pl, res_qsim + (size_t)m0 * nq, pl->d_qsim.p, (size_t)nm * nq * sizeof(float)));
        CU(cudaStreamSynchronize(pl->stream));
        for (int m = 0; m < nm; m++) res_cost[m0 + m] = jobs[m];
    }
    // regularisation term of each member (uniform planes against the caller's background), mwd_cost.f90:280-303
    if (setup->njr > 0 && setup->wjreg != 0.0f) {
        const size_t nc = (size_t)mesh->nrow * mesh->ncol;
        std::vector<std::vector<float>> pc(SMASH_B200_GNP), sc(SMASH_B200_GNS);
        SmashParameters pm{};
        SmashStates sm{};
        float **pv = pm.v, **sv = sm.v;
        std::vector<int> planes;
        for (int m = 0; m < ns; m++) {
            for (int i = 0; i < SMASH_B200_GNP; i++) { pv[i] = nullptr; if (par->v[i]) { pc[i].assign(par->v[i], par->v[i] + nc); pv[i] = pc[i].data(); } }
            for (int i = 0; i < SMASH_B200_GNS; i++) { sv[i] = nullptr; if (st->v[i]) { sc[i].assign(st->v[i], st->v[i] + nc); sv[i] = sc[i].data(); } }
            for (int j = 0; j < nvar; j++) {
                const int k = ind[j] - 1;
                float *plane = k < SMASH_B200_GNP ? pv[k] : sv[k - SMASH_B200_GNP];
                if (plane) std::fill(plane, plane + nc, sample[(size_t)m * nvar + j]);
            }
            float jreg = 0.0f;
            TRY(jreg_device(*pl, setup, mesh, &pm, par, &sm, st, 0.0f, &jreg, planes));
            CU(cudaStreamSynchronize(pl->stream));
            res_cost[m] = res_cost[m] + setup->wjreg * jreg;
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// services
// ------------------------------------------------------------------------------------------------
extern "C" const char *smash_b200_last_error(void) { return g_err.c_str(); }
extern "C" const char *smash_b200_version(void) { return "smash_b200 0.1.0 (sm_100a; reference smash v0.5.0 solver ABI)"; }
extern "C" int smash_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int smash_b200_set_device(int device) {
    TRY(check_device());
    CU(cudaSetDevice(device));
    return 0;
}
extern "C" void smash_b200_clear_cache(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_plans.clear();
    unpin_all();
}
extern "C" int smash_b200_set_option(const char *name, long long value) {
    if (!name) return fail(SMASH_B200_EINVAL, "option name is NULL");
    std::lock_guard<std::mutex> lk(g_mu);
    options()[name] = value;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// device-resident plan API
// ------------------------------------------------------------------------------------------------
extern "C" int smash_b200_plan_create(const SmashSetup *setup, const SmashMesh *mesh, int32_t nmember, SmashPlan **plan) {
    if (!plan) return fail(SMASH_B200_EINVAL, "plan is NULL");
    TRY(check_device());
    std::unique_ptr<SmashPlan> pl(new SmashPlan());
    pl->small_windows = option("plan_small_windows", 0) != 0;          // diagnostics: 256-step routing windows in a resident plan
    pl->plan_api = true;
    TRY(plan_build(*pl, setup, mesh, nmember, nmember > 1 ? (int)option("ensemble_engine", -1) : -1));
    pl->nmember = nmember > 0 ? nmember : 1;
    pl->ensemble = nmember > 1;
    *plan = pl.release();
    return 0;
}
extern "C" void smash_b200_plan_destroy(SmashPlan *plan) { delete plan; }

struct PlanSetupCopy { SmashSetup setup; SmashMesh mesh; };

extern "C" int smash_b200_plan_set_forcing(SmashPlan *plan, const SmashSetup *setup, const SmashInputData *in) {
    if (!plan || !setup) return fail(SMASH_B200_EINVAL, "NULL argument");
    SmashMesh m{};
    m.nrow = plan->tp.nrow; m.ncol = plan->tp.ncol; m.ng = plan->tp.ng;
    int nac = 0;
    for (int s = 0; s < plan->tp.nslots; s++) nac = std::max(nac, plan->tp.sparse_k[s] + 1);
    m.nac = nac;
    SmashInputData i2 = *in;
    i2.forcing_version = 0;
    TRY(plan_set_forcing(*plan, setup, &m, &i2));
    CU(cudaStreamSynchronize(plan->stream));
    return 0;
}

extern "C" int smash_b200_plan_set_fields(SmashPlan *plan, const SmashParameters *par, const SmashStates *st, const float *sample,
                                          const int32_t *ind, int32_t nvar) {
    if (!plan || !par || !st) return fail(SMASH_B200_EINVAL, "NULL argument");
    const int nm = plan->nmember > 0 ? plan->nmember : 1;
    TRY(plan_members(*plan, nm, true, false, false));
    TRY(plan_set_fields(*plan, par, st, sample, ind, nvar, nm));
    CU(cudaStreamSynchronize(plan->stream));
    return 0;
}

static int plan_cost_setup(SmashPlan *plan) {
    // plans created through the plan API evaluate NSE with unit weight on every gauge that has observations
    (void)plan;
    return 0;
}

extern "C" int smash_b200_plan_run_forward(SmashPlan *plan, float *elapsed_ms) {
    if (!plan) return fail(SMASH_B200_EINVAL, "plan is NULL");
    if (!plan->have_forcing) return fail(SMASH_B200_EINVAL, "plan has no forcing");
    TRY(plan_cost_setup(plan));
    plan->launches = 0;
    CU(cudaEventRecord(plan->ev0, plan->stream));
    TRY(run_forward_engine(*plan, plan->d_qdom.p != nullptr && !option("debug_nosave", 0), false, false));
    CU(cudaEventRecord(plan->ev1, plan->stream));
    CU(cudaEventSynchronize(plan->ev1));
    if (elapsed_ms) CU(cudaEventElapsedTime(elapsed_ms, plan->ev0, plan->ev1));
    TRY(tick_check(*plan));
    return 0;
}

extern "C" int smash_b200_plan_run_gradient(SmashPlan *plan, float *ms_fwd, float *ms_rev) {
    if (!plan) return fail(SMASH_B200_EINVAL, "plan is NULL");
    if (!plan->have_forcing) return fail(SMASH_B200_EINVAL, "plan has no forcing");
    const int nm = plan->nmember > 0 ? plan->nmember : 1;
    TRY(plan_members(*plan, nm, true, false, true));
    plan->launches = 0;
    const Topology &tp = plan->tp;
    CU(cudaEventRecord(plan->ev0, plan->stream));
    TRY(run_forward_engine(*plan, plan->engine == 0, false, true));
    CU(cudaEventRecord(plan->ev1, plan->stream));
    if (tp.ng > 0 && plan->have_qobs) {
        // NSE on every gauge, weight 1/ng (Optimize_SetupDT default wgauge, mwd_setup.f90:172)
        std::vector<float> wg(tp.ng, 1.0f / tp.ng);
        TRY(plan->d_wgauge.upload(wg, plan->stream));
        CostArgs c{};
        c.T = tp.T; c.ng = tp.ng; c.nmember = nm; c.start = 0; c.dt = plan->dt; c.dx = plan->dx;
        c.qsim = plan->d_qsim.p; c.qobs = plan->d_qobs.p; c.area = plan->d_area.p; c.wgauge = plan->d_wgauge.p;
        c.gauge_flwacc = plan->d_gauge_flwacc.p; c.njf = 1; c.jobs_fun[0] = SMASH_JOBS_NSE; c.wjobs_fun[0] = 1.0f;
        c.jobs_b = 1.0f; c.cost_jobs = plan->d_cost_jobs.p; c.qsim_b = plan->d_qsim_b.p;
        CU(launch_cost(c, plan->stream));
        plan->launches++;
    } else {
        CU(cudaMemsetAsync(plan->d_qsim_b.p, 0, sizeof(float) * std::max<size_t>(1, (size_t)nm * tp.T * tp.ng), plan->stream));
    }
    TRY(run_reverse_engine(*plan));
    CU(cudaEventRecord(plan->ev2, plan->stream));
    CU(cudaEventSynchronize(plan->ev2));
    if (ms_fwd) CU(cudaEventElapsedTime(ms_fwd, plan->ev0, plan->ev1));
    if (ms_rev) CU(cudaEventElapsedTime(ms_rev, plan->ev1, plan->ev2));
    return 0;
}

// Device-resident regionalisation step (bench / calibration loops): hyper-parameters -> field planes (hyper mapping kernel),
// forward + reverse sweeps, hyper adjoint reductions.  hyper_b: [7][nhyper] gradient of the live fields cp, cft, exc, lr,
// hp, hft, hlr.  ms[0..3]: mapping, forward sweep, reverse sweep, reductions (CUDA events on the plan's stream).
extern "C" int smash_b200_plan_run_hyper_gradient(SmashPlan *plan, const SmashSetup *setup, const SmashInputData *in,
                                                  const SmashParameters *hyper_par, const SmashStates *hyper_st, float *hyper_b,
                                                  float ms[4]) {
    if (!plan || !setup || !in || !hyper_par || !hyper_st) return fail(SMASH_B200_EINVAL, "NULL argument");
    if (!plan->have_forcing) return fail(SMASH_B200_EINVAL, "plan has no forcing");
    TRY(hyper_check(setup, in));
    TRY(plan_members(*plan, 1, true, false, true));
    HyperArgs ha;
    TRY(hyper_args(*plan, setup, nullptr, in, hyper_par, hyper_st, ha));
    const int nh = setup->nhyper;
    TRY(plan->d_hyper_b.ensure((size_t)NFIELD * nh));
    TRY(plan->d_partial.ensure((size_t)hyper_reduce_blocks(plan->ncols) * NFIELD * (1 + 2 * HYPER_MAX_ND)));
    cudaEvent_t e[5];
    for (auto &x : e) CU(cudaEventCreate(&x));
    CU(cudaEventRecord(e[0], plan->stream));
    CU(launch_hyper_fields(ha, plan->stream));
    CU(cudaEventRecord(e[1], plan->stream));
    float f = 0, r = 0;
    TRY(smash_b200_plan_run_gradient(plan, &f, &r));
    CU(cudaEventRecord(e[3], plan->stream));
    CU(launch_hyper_reduce(ha, plan->d_grad.p, plan->d_partial.p, plan->d_hyper_b.p, plan->stream));
    CU(cudaEventRecord(e[4], plan->stream));
    if (hyper_b) TRY(download(*plan, hyper_b, plan->d_hyper_b.p, (size_t)NFIELD * nh * sizeof(float)));
    CU(cudaStreamSynchronize(plan->stream));
    if (ms) {
        CU(cudaEventElapsedTime(&ms[0], e[0], e[1]));
        ms[1] = f; ms[2] = r;
        CU(cudaEventElapsedTime(&ms[3], e[3], e[4]));
    }
    for (auto &x : e) cudaEventDestroy(x);
    return 0;
}

extern "C" int smash_b200_plan_get_qsim(SmashPlan *plan, float *qsim, float *cost) {
    if (!plan) return fail(SMASH_B200_EINVAL, "plan is NULL");
    const int nm = plan->nmember > 0 ? plan->nmember : 1;
    if (qsim && plan->tp.ng > 0) TRY(download(*plan, qsim, plan->d_qsim.p, (size_t)nm * plan->tp.T * plan->tp.ng * sizeof(float)));
    if (cost) TRY(download(*plan, cost, plan->d_cost_jobs.p, (size_t)nm * sizeof(float)));
    CU(cudaStreamSynchronize(plan->stream));
    return 0;
}

extern "C" int smash_b200_plan_get_gradient(SmashPlan *plan, SmashParameters *par_b, SmashStates *st_b) {
    if (!plan || !par_b || !st_b) return fail(SMASH_B200_EINVAL, "NULL argument");
    const size_t nc = (size_t)plan->ncols;
    std::vector<float> grad((size_t)NFIELD * nc);
    TRY(download(*plan, grad.data(), plan->d_grad.p, grad.size() * sizeof(float)));
    CU(cudaStreamSynchronize(plan->stream));
    for (int f = 0; f < 4; f++) if (par_b->v[FIELD_PARAM[f]]) scatter_sorted(*plan, grad.data() + (size_t)f * nc, par_b->v[FIELD_PARAM[f]]);
    for (int f = 0; f < 3; f++) if (st_b->v[FIELD_STATE[f]]) scatter_sorted(*plan, grad.data() + (size_t)(4 + f) * nc, st_b->v[FIELD_STATE[f]]);
    return 0;
}

extern "C" int smash_b200_plan_checksum(SmashPlan *plan, double *sum_q) {
    if (!plan || !sum_q) return fail(SMASH_B200_EINVAL, "NULL argument");
    if (!plan->d_qdom.p) return fail(SMASH_B200_EINVAL, "plan keeps no domain discharge");
    if (plan->engine == 1 && plan->sub.on && plan->sub.d_qdom.p && plan->plan_api && !option("sub_scatter", 0))
        CU(launch_sum_domain(plan->sub.d_qdom.p, plan->sub.npad2, plan->sub.npad2, plan->tp.T, plan->d_sum.p, plan->stream));
    else if (plan->engine == 1) CU(launch_sum_domain(plan->d_qdom.p, plan->sp.qpitch, plan->sp.rg.n, plan->tp.T, plan->d_sum.p, plan->stream));
    else CU(launch_checksum(plan->dtp, plan->d_qdom.p, plan->d_sum.p, plan->stream));
    CU(cudaMemcpyAsync(sum_q, plan->d_sum.p, sizeof(double), cudaMemcpyDeviceToHost, plan->stream));
    CU(cudaStreamSynchronize(plan->stream));
    return 0;
}

extern "C" int smash_b200_plan_info(const SmashPlan *plan, int64_t info[12]) {
    if (!plan || !info) return fail(SMASH_B200_EINVAL, "NULL argument");
    const Topology &tp = plan->tp;
    info[0] = tp.nactive; info[1] = tp.nblocks; info[2] = tp.B; info[3] = tp.max_skew; info[4] = tp.total_ticks;
    info[5] = tp.n_cross_edges; info[6] = tp.n_pairs; info[7] = plan->launches; info[8] = tp.critical_ticks; info[9] = tp.max_chain_blocks; info[10] = tp.n_clusters; info[11] = tp.cluster_levels;
    if (plan->engine == 1) {   // split engine: chains instead of blocks
        const RouteGraph &rg = plan->sp.rg;
        info[3] = rg.max_height; info[4] = plan->sp.Tp; info[5] = rg.nchain; info[8] = rg.critical_cells; info[9] = rg.max_chain;
        info[10] = rg.ntask; info[11] = -1;
    }
    return 0;
}

extern "C" int smash_b200_plan_order(const SmashPlan *plan, int32_t *order, int32_t *block_of, int32_t *offset_of) {
    if (!plan) return fail(SMASH_B200_EINVAL, "plan is NULL");
    const Topology &tp = plan->tp;
    int k = 0;
    for (int s = 0; s < plan->ncols; s++) {
        if (plan->col_cell[s] < 0) continue;
        if (order) order[k] = plan->col_cell[s];
        if (block_of) block_of[k] = s / tp.B;
        if (offset_of) offset_of[k] = plan->engine == 1 ? 0 : tp.off[s];
        k++;
    }
    return 0;
}

// host-only: the device ordering of a mesh without touching CUDA (used by the CPU test-suite)
extern "C" int smash_b200_mesh_order(const SmashSetup *setup, const SmashMesh *mesh, int32_t block, int64_t info[12],
                                     int32_t *order, int32_t *block_of, int32_t *offset_of) {
    if (!setup || !mesh || !info) return fail(SMASH_B200_EINVAL, "NULL argument");
    int nact = 0;
    const int ncell = mesh->nrow * mesh->ncol;
    for (int c = 0; c < ncell; c++)
        if (mesh->active_cell[c] == 1 && (!mesh->local_active_cell || mesh->local_active_cell[c] == 1)) nact++;
    Topology tp;
    std::string err = build_topology(tp, mesh->nrow, mesh->ncol, mesh->ng, setup->ntime_step, mesh->flwdir, mesh->flwacc,
                                     mesh->active_cell, mesh->local_active_cell, mesh->path, mesh->gauge_pos,
                                     block > 0 ? block : pick_block(nact));
    if (!err.empty()) return fail(SMASH_B200_EINVAL, "%s", err.c_str());
    info[0] = tp.nactive; info[1] = tp.nblocks; info[2] = tp.B; info[3] = tp.max_skew; info[4] = tp.total_ticks;
    info[5] = tp.n_cross_edges; info[6] = tp.n_pairs; info[7] = 0; info[8] = tp.critical_ticks; info[9] = tp.max_chain_blocks; info[10] = tp.n_clusters; info[11] = tp.cluster_levels;
    int k = 0;
    for (int s = 0; s < tp.nslots; s++) {
        if (tp.cell[s] < 0) continue;
        if (order) order[k] = tp.cell[s];
        if (block_of) block_of[k] = s / tp.B;
        if (offset_of) offset_of[k] = tp.off[s];
        k++;
    }
    return 0;
}

// ---- preprocessing on the device (pre_kernels.cu) ---------------------------------------------------
namespace smash {
cudaError_t pre_flow_accumulation(int nrow, int ncol, const int32_t *flwdir, const int32_t *mask, int32_t *flwacc);
cudaError_t pre_gauge_masks(int nrow, int ncol, int ng, const int32_t *flwdir, const int32_t *gauge_pos, uint8_t *mask, uint8_t *d_mask_out);
cudaError_t pre_mean_forcing(int nrow, int ncol, int ng, int T, const int32_t *flwdir, const int32_t *gauge_pos, int n, const int32_t *cell_of,
                             const float *prcp, const float *pet, float *mean_prcp, float *mean_pet);
cudaError_t pre_interception(int n, int T, const int64_t *src_index, int64_t slab, const float *prcp, const float *pet, const int32_t *day_index,
                             float *ci_out, float *ms);
}

// replaces mw_meshing::flow_accumulation (smash/mesh/mw_meshing.f90:204-233)
extern "C" int smash_b200_flow_accumulation(int32_t nrow, int32_t ncol, const int32_t *flwdir, const int32_t *mask, int32_t *flwacc) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!flwdir || !flwacc || nrow <= 0 || ncol <= 0) return fail(SMASH_B200_EINVAL, "bad argument");
    TRY(check_device());
    CU(pre_flow_accumulation(nrow, ncol, flwdir, mask, flwacc));
    return 0;
}

// replaces mw_mask::mask_upstream_cells (solver/routine/mw_mask.f90:11-55) called once per gauge
extern "C" int smash_b200_gauge_masks(const SmashMesh *mesh, uint8_t *mask) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!mesh || !mask || !mesh->flwdir || (mesh->ng > 0 && !mesh->gauge_pos)) return fail(SMASH_B200_EINVAL, "NULL argument");
    TRY(check_device());
    CU(pre_gauge_masks(mesh->nrow, mesh->ncol, mesh->ng, mesh->flwdir, mesh->gauge_pos, mask, nullptr));
    return 0;
}

// replaces mw_forcing_statistic::compute_mean_forcing (solver/routine/mw_forcing_statistic.f90:18-75)
extern "C" int smash_b200_compute_mean_forcing(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *in, float *mean_prcp,
                                               float *mean_pet) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!setup || !mesh || !in) return fail(SMASH_B200_EINVAL, "NULL argument");
    TRY(check_device());
    const int ncell = mesh->nrow * mesh->ncol;
    if (setup->sparse_storage) {
        if (!in->sparse_prcp || !in->sparse_pet) return fail(SMASH_B200_EINVAL, "input_data.sparse_prcp / sparse_pet is NULL");
        std::vector<int32_t> cell_of;                                   // mw_sparse_storage.f90:28-45: path order over the active cells
        for (int i = 0; i < ncell; i++) {
            const int row = mesh->path[2 * i], col = mesh->path[2 * i + 1];
            if (!(row > 0 && col > 0)) continue;
            if (row > mesh->nrow || col > mesh->ncol) return fail(SMASH_B200_EINVAL, "mesh.path holds an index outside the grid");
            const int c = (row - 1) + (col - 1) * mesh->nrow;
            if (mesh->active_cell[c] == 1) cell_of.push_back(c);
        }
        if ((int)cell_of.size() != mesh->nac) return fail(SMASH_B200_EINVAL, "mesh.nac does not match the active cells on mesh.path");
        CU(pre_mean_forcing(mesh->nrow, mesh->ncol, mesh->ng, setup->ntime_step, mesh->flwdir, mesh->gauge_pos, mesh->nac, cell_of.data(),
                            in->sparse_prcp, in->sparse_pet, mean_prcp, mean_pet));
    } else {
        if (!in->prcp || !in->pet) return fail(SMASH_B200_EINVAL, "input_data.prcp / pet is NULL");
        CU(pre_mean_forcing(mesh->nrow, mesh->ncol, mesh->ng, setup->ntime_step, mesh->flwdir, mesh->gauge_pos, ncell, nullptr, in->prcp,
                            in->pet, mean_prcp, mean_pet));
    }
    return 0;
}

// replaces mw_interception_store::adjust_interception_store (solver/routine/mw_interception_store.f90:19-160)
extern "C" int smash_b200_adjust_interception_store(const SmashSetup *setup, const SmashMesh *mesh, const SmashInputData *in, int32_t nday,
                                                    const int32_t *day_index, float *ci, float *kernel_ms) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!setup || !mesh || !in || !day_index || !ci) return fail(SMASH_B200_EINVAL, "NULL argument");
    if (!mesh->active_cell) return fail(SMASH_B200_EINVAL, "mesh.active_cell is NULL");
    const int T = setup->ntime_step;
    int days = 1;
    for (int t = 1; t < T; t++) days += day_index[t] != day_index[t - 1];
    if (nday < days) return fail(SMASH_B200_EINVAL, "nday = %d but day_index holds %d days", nday, days);
    TRY(check_device());
    const int ncell = mesh->nrow * mesh->ncol;
    const float *prcp = setup->sparse_storage ? in->sparse_prcp : in->prcp, *pet = setup->sparse_storage ? in->sparse_pet : in->pet;
    if (!prcp || !pet) return fail(SMASH_B200_EINVAL, "input_data forcing arrays are NULL");
    if (setup->sparse_storage && !mesh->rowcol_to_ind_sparse) return fail(SMASH_B200_EINVAL, "mesh.rowcol_to_ind_sparse is NULL");
    std::vector<int64_t> src;
    std::vector<int32_t> cell;
    for (int c = 0; c < ncell; c++) {                                    // :100, :138 the computed cells
        if (!(mesh->active_cell[c] == 1 && (!mesh->local_active_cell || mesh->local_active_cell[c] == 1))) continue;
        int64_t k = c;
        if (setup->sparse_storage) {
            k = (int64_t)mesh->rowcol_to_ind_sparse[c] - 1;
            if (k < 0 || k >= mesh->nac) return fail(SMASH_B200_EINVAL, "mesh.rowcol_to_ind_sparse holds an index outside 1..nac");
        }
        src.push_back(k);
        cell.push_back(c);
    }
    std::vector<float> out(src.size());
    float ms = 0.0f;
    CU(pre_interception((int)src.size(), T, src.data(), setup->sparse_storage ? (int64_t)mesh->nac : (int64_t)ncell, prcp, pet, day_index,
                        out.data(), &ms));
    for (size_t i = 0; i < cell.size(); i++) ci[cell[i]] = out[i];
    if (kernel_ms) *kernel_ms = ms;
    return 0;
}

// ---- the network's Dense layers on the tensor cores (dense_tc.cu) -----------------------------------
namespace smash {
struct Mlp;
const char *mlp_create(int64_t M, int nlayer, const int32_t *sizes, const int32_t *act, Mlp **out);
void mlp_destroy(Mlp *m);
const char *mlp_forward(Mlp &m, const float *x, const float *const *W, const float *const *b, float *y, float *ms, double *flops);
const char *mlp_backward(Mlp &m, const float *gy, float *const *gW, float *const *gb, float *ms);
const char *mlp_forward_device(int64_t M, int nlayer, const int32_t *sizes, const float *x, const float *const *W, const float *const *b,
                               const int32_t *act, float *y, float *ms, double *flops);
}

// replaces the forward pass of Net (smash/core/net.py:281-299, Dense :664-670, Activation :478-482) for Dense (+ activation) chains
extern "C" int smash_b200_mlp_forward(int64_t nrows, int32_t nlayer, const int32_t *sizes, const float *x, const float *const *weight,
                                      const float *const *bias, const int32_t *activation, float *y, float *ms, double *flops) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!sizes || !x || !weight || !bias || !activation || !y) return fail(SMASH_B200_EINVAL, "NULL argument");
    TRY(check_device());
    const char *err = mlp_forward_device(nrows, nlayer, sizes, x, weight, bias, activation, y, ms, flops);
    if (err) return fail(SMASH_B200_ECUDA, "%s", err);
    return 0;
}

// the same chain kept on the device between the forward and the backward pass of a training epoch (Net._fit_d2p net.py:353-415)
extern "C" int smash_b200_mlp_create(int64_t nrows, int32_t nlayer, const int32_t *sizes, const int32_t *activation, SmashMlp **mlp) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!sizes || !activation || !mlp) return fail(SMASH_B200_EINVAL, "NULL argument");
    TRY(check_device());
    Mlp *m = nullptr;
    const char *err = mlp_create(nrows, nlayer, sizes, activation, &m);
    if (err) return fail(SMASH_B200_ECUDA, "%s", err);
    *mlp = reinterpret_cast<SmashMlp *>(m);
    return 0;
}
extern "C" void smash_b200_mlp_destroy(SmashMlp *mlp) { mlp_destroy(reinterpret_cast<Mlp *>(mlp)); }
extern "C" int smash_b200_mlp_run_forward(SmashMlp *mlp, const float *x, const float *const *weight, const float *const *bias, float *y,
                                          float *ms, double *flops) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!mlp || !weight || !bias) return fail(SMASH_B200_EINVAL, "NULL argument");
    const char *err = mlp_forward(*reinterpret_cast<Mlp *>(mlp), x, weight, bias, y, ms, flops);
    if (err) return fail(SMASH_B200_ECUDA, "%s", err);
    return 0;
}
extern "C" int smash_b200_mlp_run_backward(SmashMlp *mlp, const float *grad_y, float *const *grad_weight, float *const *grad_bias, float *ms) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!mlp || !grad_y || !grad_weight || !grad_bias) return fail(SMASH_B200_EINVAL, "NULL argument");
    const char *err = mlp_backward(*reinterpret_cast<Mlp *>(mlp), grad_y, grad_weight, grad_bias, ms);
    if (err) return fail(SMASH_B200_ECUDA, "%s", err);
    return 0;
}

// host-only: the heavy-path decomposition of the split engine (used by the CPU test-suite)
extern "C" int smash_b200_mesh_chains(const SmashMesh *mesh, int64_t info[8], int32_t *cell, int32_t *task_of, int32_t *pos_of,
                                      int32_t *down_of) {
    if (!mesh || !info) return fail(SMASH_B200_EINVAL, "NULL argument");
    RouteGraph rg;
    std::string err = build_route_graph(rg, mesh->nrow, mesh->ncol, mesh->ng, mesh->flwdir, mesh->flwacc, mesh->active_cell,
                                        mesh->local_active_cell, mesh->path, mesh->gauge_pos);
    if (!err.empty()) return fail(err.rfind("unsupported", 0) == 0 ? SMASH_B200_EUNSUPPORTED : SMASH_B200_EINVAL, "%s", err.c_str());
    info[0] = rg.n; info[1] = rg.nchain; info[2] = rg.npair; info[3] = rg.max_height; info[4] = rg.max_chain;
    info[5] = rg.critical_cells; info[6] = rg.nsrc; info[7] = (rg.direct ? 1 : 0) | ((int64_t)rg.nded << 1);
    for (int j = 0; j < rg.n; j++) {
        if (cell) cell[j] = rg.cell[j];
        if (task_of) task_of[j] = rg.cell_task[j];
        if (down_of) down_of[j] = rg.down[j];
        if (pos_of) pos_of[j] = -1;
    }
    if (pos_of)
        for (int t = 0; t < rg.ntask; t++)
            for (int e = rg.task_begin[t]; e < rg.task_begin[t + 1]; e++) pos_of[rg.task_cells[e]] = e - rg.task_begin[t];
    return 0;
}

// host-only: the schedule of the tick pass (used by the CPU test-suite).  Builds classes, reaches, stages and the deal of the
// units to `nwarp` warps, then replays the tickets the way the kernel's warps would -- every warp in its own key order,
// a ticket runs only when everything it reads has been published -- and reports whether the replay completes.
// info: [0] cells [1] tiles [2] reaches [3] largest stage [4] shallow routed cells [5] deep cells [6] pit cells [7] units per
// warp [8] longest deep chain [9] tickets replayed [10] 1 = every dependency points to a smaller key and the replay completed
// unit_of / sigma_of (per cell, or NULL): the unit that publishes the cell's discharge and its stage
extern "C" int smash_b200_mesh_tick_schedule(const SmashMesh *mesh, int32_t shallow_acc, int32_t nwarp, int32_t nwin, int64_t info[12],
                                             int32_t *unit_of, int32_t *sigma_of) {
    if (!mesh || !info || nwarp < 1 || nwin < 1) return fail(SMASH_B200_EINVAL, "bad argument");
    RouteGraph rg;
    std::string err = build_route_graph(rg, mesh->nrow, mesh->ncol, mesh->ng, mesh->flwdir, mesh->flwacc, mesh->active_cell,
                                        mesh->local_active_cell, mesh->path, mesh->gauge_pos);
    if (!err.empty()) return fail(err.rfind("unsupported", 0) == 0 ? SMASH_B200_EUNSUPPORTED : SMASH_B200_EINVAL, "%s", err.c_str());
    TickTopoHost tk;
    err = build_tick_topo(rg, shallow_acc, tk, option("tick_slack", 1) ? std::max(1, nwarp / 4) : 0);   // as window_build does
    if (!err.empty()) return fail(err.rfind("unsupported", 0) == 0 ? SMASH_B200_EUNSUPPORTED : SMASH_B200_EINVAL, "%s", err.c_str());
    std::vector<int32_t> pub;
    long long done = 0;
    int maxu = 0;
    const bool ok = replay_tick_schedule(rg, tk, nwarp, nwin, pub, done, maxu);
    const int ntile = tk.ntile;
    const long long total = (long long)(ntile + tk.nreach) * nwin;
    info[0] = rg.n; info[1] = ntile; info[2] = tk.nreach; info[3] = tk.max_sigma; info[4] = tk.nshallow; info[5] = tk.ndeep;
    info[6] = tk.npair_cells; info[7] = maxu; info[8] = tk.max_chain; info[9] = done; info[10] = (ok && done == total) ? 1 : 0;
    info[11] = tk.nx_cells;
    for (int j = 0; j < rg.n; j++) {
        if (unit_of) unit_of[j] = pub[j];
        if (sigma_of) sigma_of[j] = pub[j] >= 0 ? tk.sigma[pub[j]] : -1;
    }
    return 0;
}

// device time of the kernels of the last plan run (split engine), milliseconds, -1 where a kernel did not run:
// [0] vertical_forward, [1] route_forward, [2] rows_to_domain, [3] route_adjoint, [4] vertical_adjoint
extern "C" int smash_b200_plan_kernel_times(SmashPlan *plan, float ms[5]) {
    if (!plan || !ms) return fail(SMASH_B200_EINVAL, "NULL argument");
    for (int i = 0; i < 5; i++) ms[i] = -1.0f;
    if (plan->engine != 1) return 0;
    CU(cudaStreamSynchronize(plan->stream));
    const int from[5] = {0, 1, 2, 4, 5}, to[5] = {1, 2, 3, 5, 6};
    for (int i = 0; i < 5; i++)
        if (plan->kmark[from[i]] && plan->kmark[to[i]]) CU(cudaEventElapsedTime(&ms[i], plan->evk[from[i]], plan->evk[to[i]]));
    return 0;
}

// named integer facts about a plan (bench / diagnostics): "routed_cells", "inflow_edges", "source_cells", "engine"
extern "C" double smash_b200_plan_stat(const SmashPlan *plan, const char *name) {
    if (!plan || !name) return -1.0;
    const std::string n(name);
    if (n == "engine") return plan->engine;
    if (plan->engine == 1) {
        const RouteGraph &rg = plan->sp.rg;
        if (n == "routed_cells") return (double)(rg.n - rg.nsrc);
        if (n == "source_cells") return (double)rg.nsrc;
        if (n == "inflow_edges") return (double)rg.up.size();
        const SplitState &sp = plan->sp;
        if (n == "checkpoint") return sp.ckpt ? 1.0 : 0.0;
        if (n == "route_window") return (double)sp.W;
        if (n == "route_windows") return (double)sp.nwin;
        // bytes a gradient run keeps between its sweeps: tape (hp0, hft0), hr and w rows, q rows, checkpoints
        if (n == "tape_bytes")
            return 4.0 * ((double)sp.d_tape_hp.n + sp.d_tape_hft.n + sp.d_rows_hr.n + sp.d_rows_w.n + (sp.ckpt ? (double)sp.d_rows_seg.n : (double)sp.d_rows.n) +
                          sp.d_ckpt.n);
        if (n == "sub_engine") return plan->sub.on ? 1.0 : 0.0;
        if (n == "sub_tiles") return plan->sub.on ? (double)plan->sub.host.ntile : -1.0;
        if (n == "sub_slots") return plan->sub.on ? (double)plan->sub.host.nslot : -1.0;
        if (n == "sub_levels") return plan->sub.on ? (double)plan->sub.host.nlevel : -1.0;
        if (n == "tick_pass") return plan->win.on ? 1.0 : 0.0;
        if (n == "tick_stages") return plan->win.on ? (double)plan->win.host.max_sigma : -1.0;
        if (n == "tick_reaches") return plan->win.on ? (double)plan->win.host.nreach : -1.0;
        if (n == "tick_units_per_warp") return plan->win.on ? (double)plan->win.maxu : -1.0;
        if (n == "deep_cells") return plan->win.on ? (double)plan->win.host.ndeep : -1.0;
        if (n == "shallow_cells") return plan->win.on ? (double)plan->win.host.nshallow : -1.0;
    }
    return -1.0;
}
