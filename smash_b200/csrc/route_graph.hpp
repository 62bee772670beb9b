// route_graph.hpp -- host-side preprocessing for the "split" engine.
//
// The split engine separates the two halves of gr_a_forward (forward/md_forward_structure.f90:30-214):
//   * the per-cell reservoirs (interception, production, exchange, transfer: :106-144) have no
//     inter-cell dependency and run as one thread per cell over the whole time loop;
//   * the D8 routing (upstream_discharge + linear_routing, operator/md_routing_operator.f90:17-79)
//     is a linear recurrence in time per cell, fed by the upstream cells' discharge series.  It is
//     evaluated cell by cell on whole time series ("row" = one cell, all time steps), one warp per
//     chain of the heavy-path decomposition of the drainage forest, the time axis spread over the
//     32 lanes (scan).  A chain keeps its running discharge series in registers from cell to cell;
//     only tributaries ("laterals") come from memory, behind one done-flag per chain.
//
// Cells are numbered j = 0..n-1 in the reference's `path` order restricted to computed cells
// (active_cell == 1 and local_active_cell == 1), which is also the order of sparse storage
// (routine/mw_sparse_storage.f90:28-45) when every active cell is computed.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace smash {

enum : int32_t { UP_NOWAIT = -1, UP_HEAVY = -2, UP_PARTNER = -3 };

struct RouteUp {     // one inflow of a cell, in the reference's summation order i = 1..8
    int32_t src;     // producer cell j
    int32_t task;    // >= 0: wait for this task's done flag; UP_NOWAIT: source cell (final after the reservoir pass);
                     // UP_HEAVY: previous cell of the same chain (series held in registers); UP_PARTNER: pit partner
};

struct TaskCell {    // one cell of a task, stored contiguously in execution order
    int32_t j;       // cell
    int32_t meta;    // bit 0: routed (flwacc > 1), bit 1: a gauge sits on the cell, bits 8..: number of inflow entries
    int32_t up_off;  // first inflow entry of the cell in RouteGraph::tup
    int32_t pad_;
};

struct RouteGraph {
    int nrow = 0, ncol = 0, ng = 0;
    int n = 0;                          // computed cells
    int nsrc = 0;                       // cells with flwacc == 1 come first in path order?  (count only)
    int npad = 0;                       // n rounded up to a multiple of 32
    bool direct = false;                // j == sparse index for every cell and n == nac (sparse arrays usable as they are)
    std::vector<int32_t> cell;          // flat rect index row + col*nrow
    std::vector<int32_t> sparse_k;      // sparse storage index
    std::vector<int32_t> flwacc;
    std::vector<int32_t> up_begin;      // n + 1
    std::vector<RouteUp> up;
    std::vector<int32_t> down;          // consumer cell (its up list holds j) or -1
    std::vector<int32_t> down_task;     // task of the consumer cell (reverse sweep waits for it) or -1
    std::vector<int32_t> cell_task;     // task that routes cell j, -1 for source cells outside every task
    // tasks: chains first (sorted by dependency height, longest first inside a height), then pit pairs
    int ntask = 0, nchain = 0, npair = 0;
    int nded = 0;                       // the last nded chain tasks are the longest chains (dedicated warps in the forward pass)
    std::vector<int32_t> task_begin;    // ntask + 1
    std::vector<int32_t> task_cells;    // cells of each task, upstream -> downstream (pairs: early cell, late cell)
    std::vector<TaskCell> tcell;        // parallel to task_cells
    std::vector<RouteUp> tup;           // inflow entries of the task cells, task order (reference summation order per cell)
    int first_routed = 0;               // smallest j with flwacc > 1 (cells are sorted by flwacc, so sources come first)
    int max_height = 0, max_chain = 0;
    int64_t critical_cells = 0;         // longest dependency chain in cells
    // gauges (md_forward_structure.f90:206-210)
    std::vector<int32_t> gauge_first;   // per cell: first gauge on it or -1
    std::vector<int32_t> gauge_next;    // per gauge
    std::vector<int32_t> gauge_cell;    // per gauge: cell j or -1
    std::vector<int32_t> j_of_cell;     // rect index -> j or -1
};

// Returns "" on success; "unsupported: ..." when the mesh needs the fused engine (lagged inflows outside pit pairs).
std::string build_route_graph(RouteGraph &g, int nrow, int ncol, int ng, const int32_t *flwdir, const int32_t *flwacc,
                              const int32_t *active_cell, const int32_t *local_active_cell, const int32_t *path,
                              const int32_t *gauge_pos, int ded_min = 96, int ded_max = 64, int reach = 0, int order = 0,
                              const uint8_t *deep_mask = nullptr);
// deep_mask (per computed cell j, or nullptr): graph of the chain scans that follow the window pass (window_kernels.cu).
// Only cells with deep_mask[j] != 0 gather and are routed; every other cell is final (a "source") when the scans start.
// order: ticket order of the chains -- 0 by the topological level of the chain's last cell, 1 by its distance to the outlet
// (farthest first)
// ded_min / ded_max: chains of at least ded_min cells, the longest first, get CTAs of their own in the forward routing
// pass, at most ded_max CTAs.  reach > 0: those chains are cut into reaches of at most `reach` cells, one CTA per reach
// (tick wavefront, the reaches of a river pipelined); reach = 0: one CTA per whole chain (window scan per cell).

// host side of the window pass topology (WfTopo, split_kernels.cuh)
struct WindowTopoHost {
    std::vector<int32_t> meta, upoff, ups;
    std::vector<uint8_t> tile_rounds, deep;
    int nshallow = 0, ndeep = 0, nrow = 0, nx_cells = 0, max_round = 0;
};
// Classes of the window pass from the full route graph; shallow_acc = largest flwacc routed inside the window pass.
// Returns "" or "unsupported: ...".
std::string build_window_topo(const RouteGraph &g, int shallow_acc, WindowTopoHost &out);

}  // namespace smash
