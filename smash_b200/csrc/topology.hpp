// topology.hpp -- host-side D8 mesh preprocessing for the B200 solver.
//
// Turns the reference's (flwdir, flwacc, active_cell, path) description
// (derived_type/mwd_mesh.f90:45-72) into the device ordering used by the kernels:
//   * the dependency edges that upstream_discharge (operator/md_routing_operator.f90:17-60) implies,
//     in its neighbour order i = 1..8, classified "same-step" (source earlier in `path`) or "lagged"
//     (source later in `path`: the reader sees the previous time step's value);
//   * a post-order numbering of the drainage forest (children before parents, heavy child last) cut
//     into blocks of B consecutive cells = one CTA each, so that most edges stay inside a block;
//   * per cell an in-block skew `off`: at tick d a cell works on time step t = d - off, every
//     in-block producer is exactly one tick ahead of its consumer (double-buffered exchange in
//     shared memory), cross-block producers are in lower-numbered blocks (progress flags in HBM).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace smash {

struct UpEntry {   // one inflow of a cell, kept in the reference's summation order
    int32_t a;     // >= 0: lane of the producer inside the same block ; < 0: -(index into ext)-1
    int32_t cur;   // 1: read this tick's value (late cell reading its pit partner), 0: previous tick's
};
struct ExtRef {    // producer (forward) or consumer-side downstream (reverse) living in another block
    int64_t base;  // element offset of (block, tick 0 + dtick, lane) inside the skewed q / w array
    int32_t blk;
    int32_t dtick; // tick of the other block = my tick + dtick
    int32_t lag;   // 1: the reader uses the value of time step t-1 (forward) / t+1 (reverse)
    int32_t pad_;
};

struct Topology {
    int nrow = 0, ncol = 0, ng = 0, T = 0;
    int B = 0;                 // lanes per block (CTA size)
    int nblocks = 0;
    int nactive = 0;           // cells computed (active_cell == 1 and local_active_cell == 1)
    int nslots = 0;            // nblocks * B
    int64_t total_ticks = 0;   // sum over blocks of (T + hmax)
    int max_skew = 0;
    int n_cross_edges = 0, n_pairs = 0;
    int64_t critical_ticks = 0; // longest dependency chain over blocks, in ticks (start delay + own ticks)
    int max_chain_blocks = 0;
    int n_clusters = 0, cluster_levels = 0;  // connected clusters of the forest partition and depth of the cluster tree   // longest chain of blocks linked by cross-block edges

    // per slot (size nslots); cell == -1 marks a padding lane
    std::vector<int32_t> cell;        // 0-based flat rect index row + col*nrow
    std::vector<int32_t> sparse_k;    // 0-based sparse forcing index (mw_sparse_storage.f90:12-49), -1 if none
    std::vector<int32_t> off;         // in-block skew
    std::vector<int32_t> flwacc;      // flow accumulation of the cell
    std::vector<uint8_t> late;        // forward: routing done in phase 2 (second member of a pit pair)
    std::vector<uint8_t> early;       // reverse: whole step done in phase 2 (first member of a pit pair)
    std::vector<int32_t> up_begin;    // nslots + 1, CSR into up
    std::vector<UpEntry> up;
    std::vector<ExtRef> ext;          // forward cross-block producers
    // reverse sweep: the single downstream consumer of each slot
    std::vector<int32_t> down_kind;   // 0 none, 1 in-block (previous reverse tick), 2 in-block same tick, 3 other block
    std::vector<int32_t> down_lane;   // lane (kind 1,2) or index into rext (kind 3)
    std::vector<ExtRef> rext;
    std::vector<int32_t> gauge_first; // per slot: first gauge sitting on this cell or -1
    std::vector<int32_t> gauge_next;  // per gauge: next gauge on the same cell or -1
    std::vector<int32_t> gauge_slot;  // per gauge: slot of its cell (-1: gauge on a non-computed cell)

    // per block
    std::vector<int32_t> hmax;        // ticks = T + hmax
    std::vector<int64_t> tick_base;   // first row of the block in the [block][tick][...][B] arrays (in rows)
    std::vector<uint8_t> flags;       // bit0 has in-block edges, bit1 has late cells, bit2 has cross-block consumers,
                                      // bit3 has cross-block producers (forward), bit4 reverse cross-block consumers
    // rect index -> slot (-1 if not computed)
    std::vector<int32_t> slot_of_cell;
    uint64_t mesh_hash = 0;
};

enum { BLK_INTRA = 1, BLK_LATE = 2, BLK_PUBLISH = 4, BLK_EXTDEP = 8, BLK_RPUBLISH = 16, BLK_REXTDEP = 32 };

// Builds the ordering.  Returns an empty string on success, an error message otherwise.
std::string build_topology(Topology &tp, int nrow, int ncol, int ng, int T, const int32_t *flwdir,
                           const int32_t *flwacc, const int32_t *active_cell, const int32_t *local_active_cell,
                           const int32_t *path, const int32_t *gauge_pos, int block_size);

uint64_t hash_mesh(int nrow, int ncol, int ng, const int32_t *flwdir, const int32_t *flwacc,
                   const int32_t *active_cell, const int32_t *local_active_cell, const int32_t *path,
                   const int32_t *gauge_pos);

}  // namespace smash
