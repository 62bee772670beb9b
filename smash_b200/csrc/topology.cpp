// topology.cpp -- see topology.hpp.
#include "topology.hpp"

#include <algorithm>
#include <cstring>

namespace smash {

// operator/md_routing_operator.f90:29-31: neighbour i (1..8) sits at (row + drow[i], col + dcol[i]) and flows
// into (row, col) iff its flwdir == i.  Hence a cell with flwdir == i drains to (row - drow[i], col - dcol[i]).
static const int DCOL[8] = {0, -1, -1, -1, 0, 1, 1, 1};
static const int DROW[8] = {1, 1, 0, -1, -1, -1, 0, 1};

uint64_t hash_mesh(int nrow, int ncol, int ng, const int32_t *flwdir, const int32_t *flwacc,
                   const int32_t *active_cell, const int32_t *local_active_cell, const int32_t *path,
                   const int32_t *gauge_pos) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ ((uint64_t)nrow << 32) ^ (uint64_t)ncol ^ ((uint64_t)ng << 48);
    auto mix = [&](const int32_t *p, size_t n) {
        if (!p) { h = (h ^ 0xABCDEFull) * 0x100000001B3ull; return; }
        size_t n2 = n / 2;
        const uint64_t *q = reinterpret_cast<const uint64_t *>(p);
        bool aligned = (reinterpret_cast<uintptr_t>(p) % 8) == 0;
        if (aligned) {
            for (size_t i = 0; i < n2; i++) { h ^= q[i]; h *= 0x9FB21C651E98DF25ull; h ^= h >> 29; }
            if (n & 1) { h ^= (uint32_t)p[n - 1]; h *= 0x9FB21C651E98DF25ull; }
        } else {
            for (size_t i = 0; i < n; i++) { h ^= (uint32_t)p[i]; h *= 0x9FB21C651E98DF25ull; h ^= h >> 29; }
        }
    };
    size_t nc = (size_t)nrow * ncol;
    mix(flwdir, nc); mix(flwacc, nc); mix(active_cell, nc); mix(local_active_cell, nc); mix(path, 2 * nc);
    mix(gauge_pos, (size_t)2 * ng);
    return h;
}

std::string build_topology(Topology &tp, int nrow, int ncol, int ng, int T, const int32_t *flwdir,
                           const int32_t *flwacc, const int32_t *active_cell, const int32_t *local_active_cell,
                           const int32_t *path, const int32_t *gauge_pos, int block_size) {
    const int ncell = nrow * ncol;
    if (nrow <= 0 || ncol <= 0 || T <= 0) return "mesh: nrow, ncol and ntime_step must be positive";
    if (block_size < 32 || block_size > 512 || block_size % 32) return "block size must be a multiple of 32 in [32,512]";
    tp = Topology();
    tp.nrow = nrow; tp.ncol = ncol; tp.ng = ng; tp.T = T; tp.B = block_size;
    const int B = block_size;

    // ---- rank in `path` (md_forward_structure.f90:82-92) and sparse index (mw_sparse_storage.f90:28-45)
    std::vector<int32_t> rank(ncell, -1), sparse_k(ncell, -1);
    std::vector<int32_t> computed;  // flat indices in path order
    computed.reserve(ncell);
    int ks = 0;
    for (int i = 0; i < ncell; i++) {
        int row = path[2 * i], col = path[2 * i + 1];
        if (!(row > 0 && col > 0)) continue;
        if (row > nrow || col > ncol) return "mesh.path holds an index outside the grid";
        int c = (row - 1) + (col - 1) * nrow;
        if (active_cell[c] == 1) {
            if (sparse_k[c] < 0) sparse_k[c] = ks++;
            if (!local_active_cell || local_active_cell[c] == 1) {
                if (rank[c] >= 0) return "mesh.path lists a cell twice";
                rank[c] = (int32_t)computed.size();
                computed.push_back(c);
            }
        }
    }
    const int n = (int)computed.size();
    tp.nactive = n;
    if (n == 0) return "mesh has no active cell";

    // ---- inflow edges of every computed cell, neighbour order i = 1..8 (md_routing_operator.f90:37-53)
    // node ids are ranks (path order).  in_begin/in_src: CSR; lagged flag per edge.
    std::vector<int32_t> in_begin(n + 1, 0), in_src;
    std::vector<uint8_t> in_lag;
    in_src.reserve(n);
    in_lag.reserve(n);
    std::vector<int32_t> parent(n, -1);      // tree parent (rank) or -1
    std::vector<uint8_t> parent_delta(n, 1); // skew difference to the parent: 1 normal, 0 lagged child / pit partner
    for (int r = 0; r < n; r++) {
        int c = computed[r];
        int row = c % nrow, col = c / nrow;  // 0-based
        in_begin[r] = (int32_t)in_src.size();
        if (flwacc[c] > 1) {
            for (int i = 0; i < 8; i++) {
                int rr = row + DROW[i], cc = col + DCOL[i];
                if (rr < 0 || rr >= nrow || cc < 0 || cc >= ncol) continue;
                int nb = rr + cc * nrow;
                if (flwdir[nb] != i + 1) continue;
                if (rank[nb] < 0) continue;  // never computed: its q stays 0 (oracle zero-initialises q)
                in_src.push_back(rank[nb]);
                in_lag.push_back(rank[nb] > r ? 1 : 0);
            }
        }
    }
    in_begin[n] = (int32_t)in_src.size();

    // ---- forest: every edge makes its source a child of its target, except the lagged half of a full
    // 2-cycle ("pit pair"), which would close a loop.
    std::vector<uint8_t> late(n, 0), early(n, 0);
    std::vector<int32_t> partner(n, -1);
    for (int r = 0; r < n; r++)
        for (int e = in_begin[r]; e < in_begin[r + 1]; e++) {
            int s = in_src[e];
            if (!in_lag[e]) {
                if (parent[s] >= 0 && parent[s] != r) return "mesh: a cell drains into two cells";
                parent[s] = r; parent_delta[s] = 1;
            }
        }
    for (int r = 0; r < n; r++)
        for (int e = in_begin[r]; e < in_begin[r + 1]; e++) {
            if (!in_lag[e]) continue;
            int s = in_src[e];            // s is later in path than r and drains into r
            if (parent[r] == s) {         // r also drains (same-step) into s: full pit pair, r early / s late
                late[s] = 1; early[r] = 1; partner[s] = r; partner[r] = s;
                parent_delta[r] = 0;
                tp.n_pairs++;
            } else {
                if (parent[s] >= 0) return "mesh: lagged inflow from a cell that also drains elsewhere";
                parent[s] = r; parent_delta[s] = 0;  // lagged child: same skew as its consumer
            }
        }
    for (int r = 0; r < n; r++)
        if (late[r] && early[r]) return "mesh: chained flow-direction cycles are not supported";

    // children lists (ascending flwacc, pit partner forced last)
    std::vector<int32_t> ch_begin(n + 1, 0), ch;
    {
        std::vector<int32_t> cnt(n, 0);
        for (int r = 0; r < n; r++) if (parent[r] >= 0) cnt[parent[r]]++;
        for (int r = 0; r < n; r++) ch_begin[r + 1] = ch_begin[r] + cnt[r];
        ch.resize(ch_begin[n]);
        std::vector<int32_t> pos(ch_begin.begin(), ch_begin.end() - 1);
        for (int r = 0; r < n; r++) if (parent[r] >= 0) ch[pos[parent[r]]++] = r;
        for (int r = 0; r < n; r++)
            std::sort(ch.begin() + ch_begin[r], ch.begin() + ch_begin[r + 1], [&](int a, int b) {
                bool pa = (partner[r] == a), pb = (partner[r] == b);
                if (pa != pb) return pb;  // partner last
                int fa = flwacc[computed[a]], fb = flwacc[computed[b]];
                if (fa != fb) return fa < fb;
                return a < b;
            });
    }
    // roots, largest basins first
    std::vector<int32_t> roots;
    for (int r = 0; r < n; r++) if (parent[r] < 0) roots.push_back(r);
    std::sort(roots.begin(), roots.end(), [&](int a, int b) {
        int fa = flwacc[computed[a]], fb = flwacc[computed[b]];
        if (fa != fb) return fa > fb;
        return a < b;
    });

    // ---- iterative post-order
    std::vector<int32_t> post;
    post.reserve(n);
    {
        std::vector<int32_t> stack_node, stack_it;
        for (int root : roots) {
            stack_node.push_back(root); stack_it.push_back(ch_begin[root]);
            while (!stack_node.empty()) {
                int u = stack_node.back();
                int &it = stack_it.back();
                if (it < ch_begin[u + 1]) {
                    int v = ch[it++];
                    stack_node.push_back(v); stack_it.push_back(ch_begin[v]);
                } else {
                    post.push_back(u);
                    stack_node.pop_back(); stack_it.pop_back();
                }
            }
        }
    }
    if ((int)post.size() != n) return "mesh: flow directions contain a cycle longer than two cells";

    // ---- partition the forest into connected clusters of at most B cells (bottom-up, Kundu-Misra style: when the
    // residual subtree of a cell exceeds the capacity its heaviest child subtrees are cut off), then pack clusters of
    // the same depth of the cluster tree into blocks.  Producers of a block are always one cluster-level below, so
    // blocks ordered by level only ever wait for lower-numbered blocks, and the chain of dependent blocks is as
    // short as the river network allows (a post-order cut into fixed chunks chains almost every block to its
    // predecessor).  A pit pair is never separated.
    std::vector<int32_t> slot_rank;  // slot -> rank or -1
    {
        std::vector<int32_t> w(n, 1);
        std::vector<uint8_t> cut(n, 0);
        std::vector<std::pair<int32_t, int32_t>> kids;
        for (int k = 0; k < n; k++) {
            const int v = post[k];
            int64_t tot = 1;
            for (int e = ch_begin[v]; e < ch_begin[v + 1]; e++) tot += w[ch[e]];
            const int cap = early[v] ? B - 1 : B;
            if (tot > cap) {
                kids.clear();
                for (int e = ch_begin[v]; e < ch_begin[v + 1]; e++)
                    if (partner[v] != ch[e]) kids.push_back({w[ch[e]], ch[e]});
                std::sort(kids.begin(), kids.end(), [](const std::pair<int32_t, int32_t> &x, const std::pair<int32_t, int32_t> &y) {
                    return x.first != y.first ? x.first > y.first : x.second < y.second;
                });
                for (size_t i = 0; i < kids.size() && tot > cap; i++) { cut[kids[i].second] = 1; tot -= kids[i].first; }
                if (tot > cap) return "internal: cluster capacity exceeded";
            }
            w[v] = (int32_t)tot;
        }
        // cluster ids, parents first
        std::vector<int32_t> cid(n, -1), csize, cparent, croot;
        for (int k = n - 1; k >= 0; k--) {
            const int v = post[k];
            if (parent[v] < 0 || cut[v]) {
                cid[v] = (int32_t)csize.size();
                csize.push_back(0);
                croot.push_back(v);
                cparent.push_back(parent[v] < 0 ? -1 : cid[parent[v]]);
            } else cid[v] = cid[parent[v]];
            csize[cid[v]]++;
        }
        const int nc = (int)csize.size();
        std::vector<int32_t> clevel(nc, 0);
        int maxlevel = 0;
        for (int c = nc - 1; c >= 0; c--) {   // children clusters have larger ids than their parent cluster
            if (cparent[c] >= 0) clevel[cparent[c]] = std::max(clevel[cparent[c]], clevel[c] + 1);
            maxlevel = std::max(maxlevel, clevel[c]);
        }
        // member lists in post-order
        std::vector<int32_t> cbegin(nc + 1, 0), cmem(n);
        for (int c = 0; c < nc; c++) cbegin[c + 1] = cbegin[c] + csize[c];
        {
            std::vector<int32_t> pos(cbegin.begin(), cbegin.end() - 1);
            for (int k = 0; k < n; k++) cmem[pos[cid[post[k]]]++] = post[k];
        }
        // pack level by level, best fit decreasing
        std::vector<std::vector<int32_t>> by_level(maxlevel + 1);
        for (int c = 0; c < nc; c++) by_level[clevel[c]].push_back(c);
        std::vector<std::vector<int32_t>> blocks;   // cluster ids per block
        for (int lv = 0; lv <= maxlevel; lv++) {
            auto &cl = by_level[lv];
            std::sort(cl.begin(), cl.end(), [&](int x, int y) { return csize[x] != csize[y] ? csize[x] > csize[y] : x < y; });
            std::vector<std::vector<int32_t>> open(B + 1);   // open blocks of this level by remaining capacity
            for (int c : cl) {
                int rem = -1;
                for (int r = csize[c]; r <= B; r++) if (!open[r].empty()) { rem = r; break; }
                int bi;
                if (rem < 0) { bi = (int)blocks.size(); blocks.emplace_back(); rem = B; }
                else { bi = open[rem].back(); open[rem].pop_back(); }
                blocks[bi].push_back(c);
                if (rem - csize[c] > 0) open[rem - csize[c]].push_back(bi);
            }
        }
        slot_rank.assign(blocks.size() * (size_t)B, -1);
        for (size_t bi = 0; bi < blocks.size(); bi++) {
            size_t lane = 0;
            for (int c : blocks[bi])
                for (int e = cbegin[c]; e < cbegin[c + 1]; e++) slot_rank[bi * B + lane++] = cmem[e];
        }
        tp.n_clusters = nc;
        tp.cluster_levels = maxlevel + 1;
    }
    tp.nslots = (int)slot_rank.size();
    tp.nblocks = tp.nslots / B;
    std::vector<int32_t> slot_of_rank(n, -1);
    for (int s = 0; s < tp.nslots; s++) if (slot_rank[s] >= 0) slot_of_rank[slot_rank[s]] = s;

    // ---- in-block skew
    // Every in-block tree is scheduled on its own: its root (the cell whose consumer lives in another block) runs
    // as early as its own height allows, its descendants one tick ahead per hop.  Offsets relative to the block-wide
    // maximum would delay the roots of shallow trees for nothing and that delay adds up along chains of blocks.
    std::vector<int32_t> h(tp.nslots, 0), root(tp.nslots, -1), height(tp.nslots, 0);
    tp.hmax.assign(tp.nblocks, 0);
    for (int s = tp.nslots - 1; s >= 0; s--) {
        int r = slot_rank[s];
        if (r < 0) continue;
        root[s] = s;
        int p = parent[r];
        if (p >= 0) {
            int ps = slot_of_rank[p];
            if (ps / B == s / B) {
                if (ps <= s) return "internal: parent does not follow child in post-order";
                h[s] = h[ps] + parent_delta[r];
                root[s] = root[ps];
            }
        }
        height[root[s]] = std::max(height[root[s]], h[s]);
        tp.hmax[s / B] = std::max(tp.hmax[s / B], h[s]);
    }
    tp.tick_base.assign(tp.nblocks + 1, 0);
    for (int b = 0; b < tp.nblocks; b++) {
        tp.tick_base[b + 1] = tp.tick_base[b] + (T + tp.hmax[b]);
        tp.max_skew = std::max(tp.max_skew, tp.hmax[b]);
    }
    tp.total_ticks = tp.tick_base[tp.nblocks];

    tp.cell.assign(tp.nslots, -1); tp.sparse_k.assign(tp.nslots, -1); tp.off.assign(tp.nslots, 0);
    tp.flwacc.assign(tp.nslots, 1); tp.late.assign(tp.nslots, 0); tp.early.assign(tp.nslots, 0);
    tp.flags.assign(tp.nblocks, 0);
    tp.slot_of_cell.assign(ncell, -1);
    for (int s = 0; s < tp.nslots; s++) {
        int r = slot_rank[s];
        if (r < 0) continue;
        int c = computed[r];
        tp.cell[s] = c; tp.sparse_k[s] = sparse_k[c]; tp.flwacc[s] = flwacc[c];
        tp.off[s] = height[root[s]] - h[s];
        tp.late[s] = late[r]; tp.early[s] = early[r];
        tp.slot_of_cell[c] = s;
        if (late[r]) tp.flags[s / B] |= BLK_LATE;
    }

    // ---- inflow entries per slot (reference summation order preserved)
    tp.up_begin.assign(tp.nslots + 1, 0);
    tp.down_kind.assign(tp.nslots, 0); tp.down_lane.assign(tp.nslots, -1);
    for (int s = 0; s < tp.nslots; s++) {
        tp.up_begin[s] = (int32_t)tp.up.size();
        int r = slot_rank[s];
        if (r < 0) continue;
        int blk = s / B;
        for (int e = in_begin[r]; e < in_begin[r + 1]; e++) {
            int sr = in_src[e], ss = slot_of_rank[sr];
            int lag = in_lag[e];
            UpEntry u;
            if (ss / B == blk) {
                bool cur = (late[r] && partner[r] == sr);  // late cell reads its early partner in the same tick
                // same-step producers run one tick ahead; lagged producers and pit partners share the tick offset
                int want = (cur || lag) ? tp.off[s] : tp.off[s] - 1;
                if (tp.off[ss] != want) return "internal: in-block skew is not one tick";
                u.a = ss % B; u.cur = cur ? 1 : 0;
                tp.flags[blk] |= BLK_INTRA;
                // reverse: the producer's downstream is this slot
                tp.down_kind[ss] = cur ? 2 : 1;
                tp.down_lane[ss] = s % B;
            } else {
                if (ss > s) return "internal: cross-block producer is not in an earlier block";
                ExtRef x;
                x.blk = ss / B;
                x.dtick = tp.off[ss] - lag - tp.off[s];
                x.lag = lag; x.pad_ = 0;
                x.base = (tp.tick_base[x.blk] + x.dtick) * (int64_t)B + (ss % B);
                u.a = -(int32_t)tp.ext.size() - 1; u.cur = 0;
                tp.ext.push_back(x);
                tp.flags[blk] |= BLK_EXTDEP;
                tp.flags[x.blk] |= BLK_PUBLISH;
                tp.n_cross_edges++;
                // reverse: producer ss reads w of slot s (in a later block)
                ExtRef y;
                y.blk = blk;
                y.dtick = tp.off[s] - tp.off[ss] + lag;
                y.lag = lag; y.pad_ = 0;
                y.base = (tp.tick_base[blk] + y.dtick) * (int64_t)B + (s % B);
                tp.down_kind[ss] = 3;
                tp.down_lane[ss] = (int32_t)tp.rext.size();
                tp.rext.push_back(y);
                tp.flags[blk] |= BLK_RPUBLISH;
                tp.flags[x.blk] |= BLK_REXTDEP;
            }
            tp.up.push_back(u);
        }
    }
    tp.up_begin[tp.nslots] = (int32_t)tp.up.size();

    // ---- critical path over blocks: block b may run tick d once every cross-block producer p has finished its
    // tick d + dtick, i.e. start[b] >= start[p] + dtick + 1 when every block advances one tick per time unit.
    {
        std::vector<int64_t> start(tp.nblocks, 0);
        std::vector<int32_t> depth(tp.nblocks, 1);
        for (int s = 0; s < tp.nslots; s++)
            for (int e = tp.up_begin[s]; e < tp.up_begin[s + 1]; e++) {
                if (tp.up[e].a >= 0) continue;
                const ExtRef &x = tp.ext[-tp.up[e].a - 1];
                const int b = s / B;
                start[b] = std::max(start[b], start[x.blk] + x.dtick + 1);
                depth[b] = std::max(depth[b], depth[x.blk] + 1);
            }
        for (int b = 0; b < tp.nblocks; b++) {
            tp.critical_ticks = std::max<int64_t>(tp.critical_ticks, start[b] + T + tp.hmax[b]);
            tp.max_chain_blocks = std::max(tp.max_chain_blocks, depth[b]);
        }
    }

    // ---- gauges (md_forward_structure.f90:206-210)
    tp.gauge_first.assign(tp.nslots, -1);
    tp.gauge_next.assign(ng > 0 ? ng : 0, -1);
    tp.gauge_slot.assign(ng > 0 ? ng : 0, -1);
    for (int g = ng - 1; g >= 0; g--) {
        int row = gauge_pos[g], col = gauge_pos[g + ng];
        if (row < 1 || row > nrow || col < 1 || col > ncol) return "mesh.gauge_pos outside the grid";
        int s = tp.slot_of_cell[(row - 1) + (col - 1) * nrow];
        tp.gauge_slot[g] = s;
        if (s >= 0) { tp.gauge_next[g] = tp.gauge_first[s]; tp.gauge_first[s] = g; }
    }
    return "";
}

}  // namespace smash
