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

// host side of the tick pass topology (TkTopo, split_kernels.cuh; kernels in tick_kernels.cu)
struct TickTopoHost {
    int ntile = 0, nreach = 0;
    std::vector<int32_t> meta;          // [npad] bits 0-1 class (0 S source, 1 R shallow routed, 2 P pit pair, 3 D deep), bit 2: a
                                        // shallow cell or a reach reads the cell's exchange block, bit 3: a pit cell reads its row,
                                        // bit 4 gauge, bits 5-7 round inside the tile (R), bits 8-11 inflow entries in `ups`,
                                        // bit 12 (D): the previous lane of the reach holds the heavy inflow, bits 13-31 flwacc - 1
    std::vector<int32_t> upoff;         // [npad] first inflow entry (pairs of int32 in ups)
    std::vector<int32_t> ups;           // (producer cell, unit that publishes it) in the reference's summation order
    std::vector<uint8_t> tile_rounds;   // [ntile] dependency rounds inside the tile (0: no R cell)
    std::vector<uint8_t> pair;          // [n] 1 = cell of a pit pair (routed afterwards by route_pairs_kernel from rows)
    std::vector<int32_t> reach_cells;   // [nreach][32] cells of a reach, upstream -> downstream, -1 = unused lane
    std::vector<int32_t> cons1;         // [npad] unit that reads the block the cell's TILE ticket writes (S / R: the consumer's unit when
                                        // it is another unit; D: the cell's own reach, which reads the runoff), or -1
    std::vector<int32_t> cons2;         // [npad] D cells: unit that reads the block the cell's REACH ticket writes, or -1
    std::vector<int32_t> need;          // [ntile + nreach] blocks of other units a ticket of the unit reads
    std::vector<int32_t> sigma;         // [ntile + nreach] stage of every unit
    std::vector<int32_t> key;           // [max_sigma + 1] ticket key of a stage (ticket (unit, w) runs at key[sigma] + w)
    int max_sigma = 0, nshallow = 0, ndeep = 0, npair_cells = 0, nx_cells = 0, nrow_cells = 0, max_chain = 0;
};
// shallow_acc = largest flwacc routed inside the tile tickets.  Returns "" or "unsupported: ...".
// slack_units: stages with at least that many units start two ticks after the stage before (0: always one tick)
std::string build_tick_topo(const RouteGraph &g, int shallow_acc, TickTopoHost &out, int slack_units = 0);
// wunits[warp][k] = (unit, key of its stage), k < maxu, -1 = none; a warp's units are sorted by stage
void deal_tick_units(const TickTopoHost &t, int nwarp, std::vector<int32_t> &wunits, int &maxu);
// host replay of the ticket walk (tests): true = consistent and complete
bool replay_tick_schedule(const RouteGraph &rg, const TickTopoHost &tk, int nwarp, int nwin, std::vector<int32_t> &pub, long long &done,
                          int &maxu);

// ---- subtree engine (sub_kernels.cu): the engine owns the cell order -------------------------------------------------
// The drainage forest (pit pairs aside) is cut into connected subtrees ("components") of at most 32 cells and at most
// dmax + 1 cells of depth; components of the same level (1 + the largest level among the components that flow into them) are
// packed into tiles of 32 lanes.  Inside a tile a cell's inflows are lanes of the same warp (delay of a lane = delay of its
// parent - 1, so every in-tile edge spans exactly one micro-tick of the warp's wavefront); only the roots of the components
// hand their series to another tile, through exchange slots.  Tiles are numbered by level: a tile only reads lower tiles.
struct SubTopoHost {
    int ntile = 0, nslot = 0, dmax = 0, nlevel = 0, ncomp = 0, kmax = 0;
    std::vector<int32_t> cell;       // [ntile * 32] cell j (path order) of engine column j', -1 = empty lane
    std::vector<int32_t> jprime;     // [n] engine column of cell j
    std::vector<int32_t> rec;        // [ntile * 32] bit 0 valid, 1 pit cell (reservoirs only), 2 writes an exchange slot, 3 a pit cell
                                     // reads its row, 4 gauge, bits 8-11 delay, 12-15 in-tile inflows (they sit in consecutive lanes),
                                     // 16-19 exchange slots read, 20-22 position among the in-tile inflows of its parent, 23-27 lane of
                                     // the last in-tile inflow
    std::vector<uint32_t> child;     // [ntile * 32][2] lanes of the in-tile inflows, 5 bits each (6 in word 0, 2 in word 1),
                                     // reference summation order
    std::vector<int32_t> xout;       // [ntile * 32] exchange slot written by the lane, or -1
    std::vector<int32_t> extoff;     // [ntile * 32] first entry of the lane in extlist
    std::vector<int32_t> extlist;    // exchange slots read
    std::vector<uint8_t> tile_kmax;  // [ntile] largest number of in-tile inflows of a lane
    std::vector<uint8_t> tile_ext;   // [ntile] 1 = some lane reads an exchange slot
    std::vector<uint8_t> pair;       // [n] 1 = cell of a pit pair
    int64_t cells = 0, pair_cells = 0;
};
// Returns "" or "unsupported: ...".
std::string build_sub_topo(const RouteGraph &g, int dmax, SubTopoHost &out);

}  // namespace smash
