// route_graph.cpp -- see route_graph.hpp.
#include "route_graph.hpp"

#include <algorithm>
#include <numeric>

namespace smash {

// operator/md_routing_operator.f90:29-31: neighbour i (1..8) sits at (row + drow[i], col + dcol[i]) and flows into
// (row, col) iff its flwdir == i.
static const int RG_DCOL[8] = {0, -1, -1, -1, 0, 1, 1, 1};
static const int RG_DROW[8] = {1, 1, 0, -1, -1, -1, 0, 1};

std::string build_route_graph(RouteGraph &g, int nrow, int ncol, int ng, const int32_t *flwdir, const int32_t *flwacc,
                              const int32_t *active_cell, const int32_t *local_active_cell, const int32_t *path,
                              const int32_t *gauge_pos, int ded_min, int ded_max, int reach, int order, const uint8_t *deep_mask) {
    const int ncell = nrow * ncol;
    if (nrow <= 0 || ncol <= 0) return "mesh: nrow and ncol must be positive";
    g = RouteGraph();
    g.nrow = nrow; g.ncol = ncol; g.ng = ng;

    // ---- cells in path order (md_forward_structure.f90:82-92), sparse index (mw_sparse_storage.f90:28-45)
    g.j_of_cell.assign(ncell, -1);
    std::vector<int32_t> sparse_of_cell(ncell, -1);
    int ks = 0;
    for (int i = 0; i < ncell; i++) {
        const int row = path[2 * i], col = path[2 * i + 1];
        if (!(row > 0 && col > 0)) continue;
        if (row > nrow || col > ncol) return "mesh.path holds an index outside the grid";
        const int c = (row - 1) + (col - 1) * nrow;
        if (active_cell[c] != 1) continue;
        if (sparse_of_cell[c] < 0) sparse_of_cell[c] = ks++;
        if (local_active_cell && local_active_cell[c] != 1) continue;
        if (g.j_of_cell[c] >= 0) return "mesh.path lists a cell twice";
        g.j_of_cell[c] = (int32_t)g.cell.size();
        g.cell.push_back(c);
    }
    const int n = (int)g.cell.size();
    if (n == 0) return "mesh has no active cell";
    g.n = n;
    g.npad = (n + 31) / 32 * 32;
    g.sparse_k.resize(n); g.flwacc.resize(n);
    g.direct = (ks == n);
    for (int j = 0; j < n; j++) {
        g.sparse_k[j] = sparse_of_cell[g.cell[j]];
        g.flwacc[j] = flwacc[g.cell[j]];
        if (g.sparse_k[j] != j) g.direct = false;
        if (g.flwacc[j] <= 1) g.nsrc++;
    }

    // ---- inflow lists, neighbour order i = 1..8 (md_routing_operator.f90:37-53); only cells with flwacc > 1 gather (:35)
    g.up_begin.assign(n + 1, 0);
    g.down.assign(n, -1);
    std::vector<uint8_t> lagged;   // per up entry: producer later in path (reader sees the previous time step)
    auto routed = [&](int j) { return deep_mask ? deep_mask[j] != 0 : g.flwacc[j] > 1; };
    for (int j = 0; j < n; j++) {
        g.up_begin[j] = (int32_t)g.up.size();
        if (g.flwacc[j] <= 1 || !routed(j)) continue;
        const int c = g.cell[j];
        const int row = c % nrow, col = c / nrow;
        for (int i = 0; i < 8; i++) {
            const int rr = row + RG_DROW[i], cc = col + RG_DCOL[i];
            if (rr < 0 || rr >= nrow || cc < 0 || cc >= ncol) continue;
            const int nb = rr + cc * nrow;
            if (flwdir[nb] != i + 1) continue;
            const int s = g.j_of_cell[nb];
            if (s < 0) continue;   // never computed: its q stays 0
            if (g.down[s] >= 0) return "mesh: a cell drains into two cells";
            g.down[s] = j;
            g.up.push_back({s, UP_NOWAIT});
            lagged.push_back(s > j ? 1 : 0);
        }
    }
    g.up_begin[n] = (int32_t)g.up.size();

    // ---- pit pairs: j (earlier in path) reads its partner's previous step, the partner reads j's current step
    std::vector<int32_t> partner(n, -1);
    std::vector<std::pair<int32_t, int32_t>> pairs;
    for (int j = 0; j < n; j++)
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) {
            if (!lagged[e]) continue;
            const int s = g.up[e].src;
            if (g.down[j] != s) return "unsupported: lagged inflow outside a pit pair";
            if (partner[j] >= 0 || partner[s] >= 0) return "mesh: chained flow-direction cycles are not supported";
            partner[j] = s; partner[s] = j;
            pairs.push_back({j, s});
        }
    for (int j = 0; j < n; j++)
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++)
            if (partner[j] == g.up[e].src) g.up[e].task = UP_PARTNER;

    // ---- heavy-path decomposition: the heavy child of a cell is its inflow with the largest flwacc (first on ties)
    std::vector<int32_t> heavy(n, -1), next(n, -1);
    for (int j = 0; j < n; j++) {
        if (partner[j] >= 0) continue;
        int best = -1, best_fa = -1;
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) {
            const int s = g.up[e].src;
            if (partner[s] >= 0) continue;   // cannot happen (pit cells only drain into each other); kept for safety
            if (g.flwacc[s] > best_fa) { best_fa = g.flwacc[s]; best = e; }
        }
        if (best >= 0) { heavy[j] = g.up[best].src; next[g.up[best].src] = j; g.up[best].task = UP_HEAVY; }
    }
    // cycle check + chains
    std::vector<std::vector<int32_t>> chains;
    std::vector<int32_t> chain_of(n, -1);
    for (int j = 0; j < n; j++) {
        if (heavy[j] >= 0 || partner[j] >= 0) continue;          // not a chain head
        if (next[j] < 0 && !routed(j)) continue;                 // lone source cell: final after the reservoir pass
        std::vector<int32_t> ch;
        for (int c = j; c >= 0; c = next[c]) {
            if (chain_of[c] >= 0) return "mesh: flow directions contain a cycle";
            chain_of[c] = (int32_t)chains.size();
            ch.push_back(c);
        }
        chains.push_back(std::move(ch));
    }
    for (int j = 0; j < n; j++)
        if (chain_of[j] < 0 && partner[j] < 0 && (heavy[j] >= 0 || routed(j)))
            return "mesh: flow directions contain a cycle longer than two cells";

    // ---- river reaches (reach > 0): the longest chains, as many as ded_max dedicated CTAs can take, are cut into reaches
    // of at most `reach` cells.  A reach is a chain of its own whose head gathers the tail of the previous reach like any
    // tributary; the forward routing pass runs the reaches as a pipeline (tick wavefront, split_kernels.cu).
    std::vector<uint8_t> ded_reach;
    if (reach > 0) {
        std::vector<int32_t> byl(chains.size());
        std::iota(byl.begin(), byl.end(), 0);
        std::stable_sort(byl.begin(), byl.end(), [&](int a, int b) { return chains[a].size() > chains[b].size(); });
        ded_reach.assign(chains.size(), 0);
        int used = 0;
        for (int ci : byl) {
            const int sz = (int)chains[ci].size();
            if (sz < ded_min) break;
            const int nseg = (sz + reach - 1) / reach;
            if (used + nseg > ded_max) break;
            used += nseg;
            ded_reach[ci] = 1;
            if (nseg == 1) continue;
            const std::vector<int32_t> whole = chains[ci];
            const int base = sz / nseg, rem = sz % nseg;
            int at = base + (rem > 0 ? 1 : 0);
            chains[ci].assign(whole.begin(), whole.begin() + at);
            for (int k = 1; k < nseg; k++) {
                const int len = base + (k < rem ? 1 : 0);
                const int c0 = whole[at], prev = whole[at - 1];
                for (int e = g.up_begin[c0]; e < g.up_begin[c0 + 1]; e++)
                    if (g.up[e].src == prev) g.up[e].task = UP_NOWAIT;       // no longer the chain predecessor
                heavy[c0] = -1; next[prev] = -1;
                std::vector<int32_t> part(whole.begin() + at, whole.begin() + at + len);
                for (int c : part) chain_of[c] = (int32_t)chains.size();
                chains.push_back(std::move(part));
                ded_reach.push_back(1);
                at += len;
            }
        }
    }

    // ---- dependency height of every chain (laterals are tails of other chains); processed in order of the tail's
    // position in path, which is a topological order for non-lagged edges (producer earlier in path than consumer)
    const int nch = (int)chains.size();
    std::vector<int32_t> height(nch, 0);
    std::vector<int64_t> finish(nch, 0);
    {
        std::vector<int32_t> order(nch);
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int a, int b) { return chains[a].back() < chains[b].back(); });
        for (int ci : order) {
            int h = 0;
            int64_t t = 0;
            for (int c : chains[ci]) {
                int64_t ready = 0;
                for (int e = g.up_begin[c]; e < g.up_begin[c + 1]; e++) {
                    const int s = g.up[e].src;
                    const int cs = chain_of[s];
                    if (cs < 0 || cs == ci) continue;
                    if (chains[cs].back() != s) return "internal: lateral inflow is not a chain tail";
                    h = std::max(h, height[cs] + 1);
                    ready = std::max(ready, finish[cs]);
                }
                t = std::max(t, ready) + 1;
            }
            height[ci] = h;
            finish[ci] = t;
            g.max_height = std::max(g.max_height, h);
            g.critical_cells = std::max(g.critical_cells, t);
            g.max_chain = std::max(g.max_chain, (int)chains[ci].size());
        }
    }
    // the reverse sweep publishes the cells a chain has finished in the low 16 bits of its progress word (split_kernels.cu)
    if (g.max_chain >= 0xfff0) return "unsupported: a heavy-path chain of 65520 cells or more";
    // Task order: by the topological level of the chain's last cell (level = cells on the longest path from a source to the
    // cell).  A chain's tributaries end at a lower level than the cell they join, so this is a dependency order, and it is
    // the order in which the serial walks down the rivers need their tributaries: what joins the upper reaches comes first.
    std::vector<int32_t> level(n, 1);
    for (int j = 0; j < n; j++)
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++)
            if (g.up[e].src < j) level[j] = std::max(level[j], level[g.up[e].src] + 1);
    std::vector<int32_t> torder(nch);
    std::iota(torder.begin(), torder.end(), 0);
    if (order == 1) {
        // Critical-path order: by the number of cells between the chain's last cell and the outlet of its basin, farthest
        // first.  A tributary joins its consumer above the consumer's last cell, so it is farther from the outlet: this
        // is a dependency order too, and the headwaters of the long rivers no longer queue behind every short coastal
        // chain of the domain.
        std::vector<int32_t> dist(n, 1);
        for (int j = n - 1; j >= 0; j--)
            if (g.down[j] > j) dist[j] = dist[g.down[j]] + 1;
        std::stable_sort(torder.begin(), torder.end(), [&](int a, int b) {
            const int da = dist[chains[a].back()], db = dist[chains[b].back()];
            if (da != db) return da > db;
            return chains[a].size() > chains[b].size();
        });
    } else if (order == 2) {
        // Basin by basin, the basins with the longest rivers first, each in level order.  The serial walk down a main river
        // can only start once the low levels of ITS basin are done; in a domain-wide level order that is when 90 % of the
        // whole domain is done.  Basins exchange nothing, so any basin order keeps the dependencies.
        std::vector<int32_t> root(n), crit(n, 0), rank_of(n, 0);
        for (int j = n - 1; j >= 0; j--) root[j] = (g.down[j] > j) ? root[g.down[j]] : j;
        for (int j = 0; j < n; j++) crit[root[j]] = std::max(crit[root[j]], level[j]);
        std::vector<int32_t> roots;
        for (int j = 0; j < n; j++) if (root[j] == j) roots.push_back(j);
        std::stable_sort(roots.begin(), roots.end(), [&](int a, int b) { return crit[a] > crit[b]; });
        for (size_t r = 0; r < roots.size(); r++) rank_of[roots[r]] = (int32_t)r;
        std::stable_sort(torder.begin(), torder.end(), [&](int a, int b) {
            const int ta = chains[a].back(), tb = chains[b].back();
            const int ra = rank_of[root[ta]], rb = rank_of[root[tb]];
            if (ra != rb) return ra < rb;
            if (level[ta] != level[tb]) return level[ta] < level[tb];
            return chains[a].size() > chains[b].size();
        });
    } else
    std::stable_sort(torder.begin(), torder.end(), [&](int a, int b) {
        const int la = level[chains[a].back()], lb = level[chains[b].back()];
        if (la != lb) return la < lb;
        return chains[a].size() > chains[b].size();
    });
    // the longest chains get warps of their own in the forward routing pass: they are moved to the end of the chain tasks
    {
        const int DED_MIN = ded_min, DED_MAX = ded_max;
        std::vector<int32_t> byl(torder);
        std::stable_sort(byl.begin(), byl.end(), [&](int a, int b) { return chains[a].size() > chains[b].size(); });
        std::vector<uint8_t> ded(nch, 0);
        if (reach > 0) {
            for (int c = 0; c < nch; c++) if (ded_reach[c]) { ded[c] = 1; g.nded++; }
        } else
            for (int i = 0; i < nch && i < DED_MAX && (int)chains[byl[i]].size() >= DED_MIN; i++) { ded[byl[i]] = 1; g.nded++; }
        std::stable_partition(torder.begin(), torder.end(), [&](int c) { return !ded[c]; });
    }
    std::vector<int32_t> task_of_chain(nch);
    for (int t = 0; t < nch; t++) task_of_chain[torder[t]] = t;

    g.nchain = nch; g.npair = (int)pairs.size(); g.ntask = g.nchain + g.npair;
    g.task_begin.assign(g.ntask + 1, 0);
    g.cell_task.assign(n, -1);
    for (int t = 0; t < nch; t++) {
        g.task_begin[t] = (int32_t)g.task_cells.size();
        for (int c : chains[torder[t]]) { g.task_cells.push_back(c); g.cell_task[c] = t; }
    }
    for (int p = 0; p < g.npair; p++) {
        g.task_begin[nch + p] = (int32_t)g.task_cells.size();
        g.task_cells.push_back(pairs[p].first); g.task_cells.push_back(pairs[p].second);
        g.cell_task[pairs[p].first] = g.cell_task[pairs[p].second] = nch + p;
    }
    g.task_begin[g.ntask] = (int32_t)g.task_cells.size();
    // lateral entries wait for the producing task (source cells outside every task need no wait)
    for (int j = 0; j < n; j++)
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++)
            if (g.up[e].task == UP_NOWAIT) {
                const int s = g.up[e].src;
                if (g.cell_task[s] >= 0) {
                    const bool s_ded = g.cell_task[s] >= g.nchain - g.nded && g.cell_task[s] < g.nchain;
                    if (g.cell_task[s] >= g.cell_task[j] && !s_ded) return "internal: task order violates a dependency";
                    g.up[e].task = g.cell_task[s];
                }
            }
    g.down_task.assign(n, -1);
    for (int j = 0; j < n; j++) if (g.down[j] >= 0) g.down_task[j] = g.cell_task[g.down[j]];

    // ---- gauges (md_forward_structure.f90:206-210)
    g.gauge_first.assign(n, -1);
    g.gauge_next.assign(ng > 0 ? ng : 0, -1);
    g.gauge_cell.assign(ng > 0 ? ng : 0, -1);
    for (int k = ng - 1; k >= 0; k--) {
        const int row = gauge_pos[k], col = gauge_pos[k + ng];
        if (row < 1 || row > nrow || col < 1 || col > ncol) return "mesh.gauge_pos outside the grid";
        const int j = g.j_of_cell[(row - 1) + (col - 1) * nrow];
        g.gauge_cell[k] = j;
        if (j >= 0) { g.gauge_next[k] = g.gauge_first[j]; g.gauge_first[j] = k; }
    }
    // ---- compact per-task records: everything a routing warp needs about its cells in two coalesced loads
    g.tcell.resize(g.task_cells.size());
    for (size_t i = 0; i < g.task_cells.size(); i++) {
        const int j = g.task_cells[i];
        TaskCell tc;
        tc.j = j;
        const int nup = g.up_begin[j + 1] - g.up_begin[j];
        tc.meta = (routed(j) ? 1 : 0) | (g.gauge_first[j] >= 0 ? 2 : 0) | (nup << 8);
        tc.up_off = (int32_t)g.tup.size();
        tc.pad_ = 0;
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) g.tup.push_back(g.up[e]);
        g.tcell[i] = tc;
    }
    g.first_routed = n;
    for (int j = 0; j < n; j++) if (routed(j) && (!deep_mask || partner[j] < 0)) { g.first_routed = j; break; }
    return "";
}

// ------------------------------------------------------------------------------------------------
// Topology of the tick pass (tick_kernels.cu) from the full route graph.
//
// Cells are visited in path order, which is a topological order of every non-lagged edge.  Classes: S source (flwacc == 1),
// R shallow routed cell (flwacc <= shallow_acc, no inflow from a deep / pit cell, at most 8 dependency rounds inside its
// tile), P cell of a pit pair, D every other gathering cell.  The D cells are cut into heavy-path chains (heavy inflow = the
// D inflow with the largest flwacc) and the chains into reaches of at most 32 cells.
//
// Units: tile b (cells 32 b .. 32 b + 31) and reach r (unit ntile + r).  Stage sigma of a unit = 1 + the largest stage among
// the units that produce one of its inputs; ticket (unit, window w) has key sigma + w and only reads what tickets with a
// smaller key wrote.
// ------------------------------------------------------------------------------------------------
std::string build_tick_topo(const RouteGraph &g, int shallow_acc, TickTopoHost &out, int slack_units) {
    const int n = g.n, npad = g.npad, ntile = npad / 32;
    out = TickTopoHost();
    out.ntile = ntile;
    out.meta.assign(npad, 0); out.upoff.assign(npad, 0); out.tile_rounds.assign(ntile, 0); out.pair.assign(n, 0);
    out.cons1.assign(npad, -1); out.cons2.assign(npad, -1);
    std::vector<uint8_t> cls(n, 0), round(n, 0);
    for (int j = 0; j < n; j++) {
        if (g.flwacc[j] <= 1) { cls[j] = 0; continue; }
        // flwacc > 1 without a computed inflow: still routed (qup = 0, md_forward_structure.f90:146-156)
        bool pair = false, deep = g.flwacc[j] > shallow_acc;
        int r = 0;
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) {
            const int s = g.up[e].src;
            if (g.up[e].task == UP_PARTNER) { pair = true; break; }
            if (s >= j) return "unsupported: lagged inflow outside a pit pair";
            if (cls[s] >= 2) deep = true;
            else if (cls[s] == 1 && (s >> 5) == (j >> 5)) r = std::max(r, round[s] + 1);
        }
        if (r > 7) deep = true;
        cls[j] = pair ? 2 : deep ? 3 : 1;
        round[j] = cls[j] == 1 ? (uint8_t)r : 0;
        if (pair) { out.pair[j] = 1; out.npair_cells++; }
    }
    for (int j = 0; j < n; j++)
        if (cls[j] == 2 && g.down[j] >= 0 && cls[g.down[j]] != 2) return "unsupported: a pit pair drains into another cell";

    // ---- heavy-path chains of the D cells, cut into reaches
    std::vector<int32_t> heavy(n, -1), next(n, -1), reach_of(n, -1), lane_of(n, -1);
    for (int j = 0; j < n; j++) {
        if (cls[j] != 3) continue;
        int best = -1, best_fa = -1;
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) {
            const int s = g.up[e].src;
            if (cls[s] == 3 && g.flwacc[s] > best_fa) { best_fa = g.flwacc[s]; best = s; }
        }
        if (best >= 0) { heavy[j] = best; next[best] = j; }
    }
    for (int j = 0; j < n; j++) {
        if (cls[j] != 3 || heavy[j] >= 0) continue;                      // chain heads
        int len = 0;
        for (int c = j; c >= 0; c = next[c], len++) {
            if (len % 32 == 0) { out.reach_cells.resize(out.reach_cells.size() + 32, -1); out.nreach++; }
            reach_of[c] = out.nreach - 1; lane_of[c] = len % 32;
            out.reach_cells[(size_t)(out.nreach - 1) * 32 + len % 32] = c;
            out.ndeep++;
        }
        out.max_chain = std::max(out.max_chain, len);
    }
    auto owner = [&](int s) { return cls[s] == 3 ? ntile + reach_of[s] : (s >> 5); };

    // ---- inflow lists (reference summation order) and per-cell records
    for (int j = 0; j < n; j++) {
        const int d = g.down[j];
        out.upoff[j] = (int32_t)out.ups.size();
        int nup = 0;
        if (cls[j] == 1 || cls[j] == 3)
            for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) {
                const int s = g.up[e].src;
                // the previous cell of the same reach hands its discharge over inside the warp
                if (cls[j] == 3 && s == heavy[j] && reach_of[s] == reach_of[j]) continue;
                out.ups.push_back(s);
                nup++;
            }
        if (nup > 15) return "unsupported: more than 15 inflows";
        const bool chained = cls[j] == 3 && heavy[j] >= 0 && reach_of[heavy[j]] == reach_of[j];
        // who reads this cell's discharge from memory: a shallow cell or a reach reads the exchange blocks, a pit cell reads rows
        const bool handed = d >= 0 && cls[j] == 3 && cls[d] == 3 && heavy[d] == j && reach_of[d] == reach_of[j];
        const bool want_x = d >= 0 && !handed && (cls[d] == 1 || cls[d] == 3);
        const bool want_row = d >= 0 && cls[d] == 2 && cls[j] != 2;
        if (g.flwacc[j] - 1 >= (1 << 19)) return "unsupported: flow accumulation above 2^19";
        out.meta[j] = cls[j] | (want_x ? 4 : 0) | (want_row ? 8 : 0) | (g.gauge_first[j] >= 0 ? 16 : 0) | (round[j] << 5) | (nup << 8) |
                      (chained ? 4096 : 0) | (std::max(0, g.flwacc[j] - 1) << 13);
        // the unit that reads the block a tile ticket writes for this cell (a D cell: its own reach reads the runoff), and the
        // unit that reads the block its reach ticket writes
        if (cls[j] == 3) {
            out.cons1[j] = ntile + reach_of[j];
            if (want_x) out.cons2[j] = owner(d);
        } else if (want_x && owner(d) != (j >> 5)) {
            out.cons1[j] = owner(d);
        }
        if (cls[j] == 1) {
            out.nshallow++;
            out.tile_rounds[j >> 5] = std::max<uint8_t>(out.tile_rounds[j >> 5], (uint8_t)(round[j] + 1));
        }
        if (want_x) out.nx_cells++;
        if (want_row) out.nrow_cells++;
    }
    if (out.ups.empty()) out.ups.push_back(0);

    // ---- stages
    const int nunit = ntile + out.nreach;
    out.sigma.assign(nunit, 0);
    for (int j = 0; j < n; j++) {                                         // tiles: path order visits every producer tile first
        if (cls[j] != 1) continue;
        int &sg = out.sigma[j >> 5];
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) {
            const int s = g.up[e].src;
            if ((s >> 5) != (j >> 5)) sg = std::max(sg, out.sigma[s >> 5] + 1);
        }
    }
    // reaches by their last cell (the largest index of the reach: flwacc grows downstream), so every producer reach comes first
    std::vector<int32_t> order(out.nreach);
    std::vector<int32_t> last(out.nreach, -1);
    for (int r = 0; r < out.nreach; r++) {
        order[r] = r;
        for (int l = 0; l < 32; l++) last[r] = std::max(last[r], out.reach_cells[(size_t)r * 32 + l]);
    }
    std::sort(order.begin(), order.end(), [&](int x, int y) { return last[x] < last[y]; });
    for (int r : order) {
        int sg = 0;
        for (int l = 0; l < 32; l++) {
            const int c = out.reach_cells[(size_t)r * 32 + l];
            if (c < 0) continue;
            sg = std::max(sg, out.sigma[c >> 5] + 1);                   // its own runoff block
            for (int e = g.up_begin[c]; e < g.up_begin[c + 1]; e++) {
                const int s = g.up[e].src;
                if (cls[s] == 3 && reach_of[s] == r) continue;
                const int o = owner(s);
                if (o >= ntile && last[o - ntile] >= last[r]) return "internal: reach order violates a dependency";
                sg = std::max(sg, out.sigma[o] + 1);
            }
        }
        out.sigma[ntile + r] = sg;
    }
    for (int u = 0; u < nunit; u++) out.max_sigma = std::max(out.max_sigma, out.sigma[u]);
    // blocks a unit waits for per window = arrivals counted by its producers
    out.need.assign(nunit, 0);
    for (int j = 0; j < n; j++) {
        if (out.cons1[j] >= 0) out.need[out.cons1[j]]++;
        if (out.cons2[j] >= 0) out.need[out.cons2[j]]++;
    }
    // ticket key of a stage: consumers of a crowded stage start two ticks after their producers (a whole tick of slack, so
    // that nobody waits for a producer that is still being worked on); the sparse deep stages follow one tick apart
    std::vector<int32_t> count(out.max_sigma + 1, 0);
    for (int u = 0; u < nunit; u++) count[out.sigma[u]]++;
    out.key.assign(out.max_sigma + 1, 0);
    for (int sg = 1; sg <= out.max_sigma; sg++) out.key[sg] = out.key[sg - 1] + ((slack_units > 0 && count[sg] >= slack_units) ? 2 : 1);
    // blocks read `far_ticks` or more ticks after they were written are not worth keeping in L2 (bit 30 of the entries)
    auto reader = [&](int j) { return cls[j] == 3 ? ntile + reach_of[j] : (j >> 5); };
    const int far_ticks = 4;
    for (int j = 0; j < n; j++) {
        if (cls[j] != 1 && cls[j] != 3) continue;
        const int kj = out.key[out.sigma[reader(j)]];
        const int nup = out.meta[j] >> 8 & 15;
        for (int e = 0; e < nup; e++) {
            const int s = out.ups[out.upoff[j] + e];
            if (kj - out.key[out.sigma[owner(s)]] >= far_ticks) out.ups[out.upoff[j] + e] = s | (1 << 30);
        }
    }
    for (int j = 0; j < n; j++) {
        if (out.cons1[j] >= 0 && out.key[out.sigma[out.cons1[j]]] - out.key[out.sigma[j >> 5]] >= far_ticks) out.cons1[j] |= 1 << 30;
        if (out.cons2[j] >= 0 && out.key[out.sigma[out.cons2[j]]] - out.key[out.sigma[reader(j)]] >= far_ticks) out.cons2[j] |= 1 << 30;
    }
    return "";
}

// Units dealt to the warps of a resident grid: sorted by stage, round-robin, so every warp holds the same mix of stages.
void deal_tick_units(const TickTopoHost &t, int nwarp, std::vector<int32_t> &wunits, int &maxu) {
    const int nunit = t.ntile + t.nreach;
    std::vector<int32_t> order(nunit);
    for (int u = 0; u < nunit; u++) order[u] = u;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return t.sigma[x] < t.sigma[y]; });
    maxu = (nunit + nwarp - 1) / nwarp;
    wunits.assign((size_t)nwarp * maxu * 2, -1);
    for (int i = 0; i < nunit; i++) {
        const int g = i % nwarp, k = i / nwarp;
        wunits[((size_t)g * maxu + k) * 2] = order[i];
        wunits[((size_t)g * maxu + k) * 2 + 1] = t.key[t.sigma[order[i]]];
    }
}

// Host replay of the tick pass (tests): every warp walks its tickets in key order, a ticket runs only when everything it reads
// has been published.  pub[j] = unit that publishes cell j's discharge.  Returns true when every dependency points to a
// smaller stage and the replay completes; done = tickets replayed.
bool replay_tick_schedule(const RouteGraph &rg, const TickTopoHost &tk, int nwarp, int nwin, std::vector<int32_t> &pub, long long &done,
                          int &maxu) {
    std::vector<int32_t> wunits;
    deal_tick_units(tk, nwarp, wunits, maxu);
    const int ntile = tk.ntile, nunit = ntile + tk.nreach;
    std::vector<std::vector<int32_t>> need(nunit);
    pub.assign(rg.npad, -1);
    bool ok = true;
    for (int r = 0; r < tk.nreach; r++)
        for (int l = 0; l < 32; l++) {
            const int c = tk.reach_cells[(size_t)r * 32 + l];
            if (c >= 0) pub[c] = ntile + r;
        }
    for (int j = 0; j < rg.n; j++) {
        const int cls = tk.meta[j] & 3;
        if (cls <= 1) pub[j] = j >> 5;
        if (cls == 3 && pub[j] < 0) return false;
    }
    for (int j = 0; j < rg.n; j++) {
        const int cls = tk.meta[j] & 3;
        if (cls != 1 && cls != 3) continue;
        const int unit = pub[j];
        if (cls == 3) need[unit].push_back(j >> 5);
        for (int e = 0; e < (tk.meta[j] >> 8 & 15); e++) {
            const int src = tk.ups[tk.upoff[j] + e] & ~(1 << 30);
            if (src < 0 || src >= rg.n || pub[src] < 0) return false;
            const int own = pub[src];
            if (own != unit) { need[unit].push_back(own); if (tk.key[tk.sigma[own]] >= tk.key[tk.sigma[unit]]) ok = false; }
        }
    }
    // the arrival counts the device waits for must equal the blocks each unit reads from other units
    for (int u = 0; u < nunit; u++)
        if ((int)need[u].size() != tk.need[u]) ok = false;
    // own[u]: windows the owner has finished; seen[u]: windows it has published (a warp publishes when it leaves a tick)
    std::vector<int32_t> own(nunit, 0), seen(nunit, 0), cur_k(nwarp, 0), cur_i(nwarp, 0), nu(nwarp, 0);
    for (int g = 0; g < nwarp; g++) {
        while (nu[g] < maxu && wunits[((size_t)g * maxu + nu[g]) * 2] >= 0) nu[g]++;
        cur_k[g] = nu[g] ? wunits[(size_t)g * maxu * 2 + 1] : 0;
    }
    done = 0;
    const long long total = (long long)nunit * nwin;
    bool moved = true;
    while (moved && done < total) {
        moved = false;
        for (int g = 0; g < nwarp; g++) {
            if (!nu[g]) continue;
            const int kmax = wunits[((size_t)g * maxu + nu[g] - 1) * 2 + 1] + nwin - 1;
            while (cur_k[g] <= kmax) {
                const int u = wunits[((size_t)g * maxu + cur_i[g]) * 2], ky = wunits[((size_t)g * maxu + cur_i[g]) * 2 + 1];
                const int w = cur_k[g] - ky;
                if (w >= 0 && w < nwin) {
                    bool ready = own[u] == w;
                    for (int p_ : need[u]) if (seen[p_] <= w) { ready = false; break; }
                    if (!ready) break;
                    own[u] = w + 1; done++; moved = true;
                }
                if (++cur_i[g] == nu[g]) {
                    cur_i[g] = 0; cur_k[g]++;
                    for (int q = 0; q < nu[g]; q++) seen[wunits[((size_t)g * maxu + q) * 2]] = own[wunits[((size_t)g * maxu + q) * 2]];
                }
            }
        }
    }
    return ok && done == total;
}

// ------------------------------------------------------------------------------------------------
// Subtree tiles (route_graph.hpp).  Cells are visited in path order (inflows first): a cell keeps the open subtrees of its
// inflows, heaviest first, while the union holds at most 32 cells and dmax + 1 cells of depth; an inflow that does not fit is
// closed and becomes the root of a component.
// ------------------------------------------------------------------------------------------------
std::string build_sub_topo(const RouteGraph &g, int dmax, SubTopoHost &out) {
    const int n = g.n;
    out = SubTopoHost();
    if (dmax < 1 || dmax > 15) return "internal: dmax out of range";
    out.dmax = dmax;
    out.pair.assign(n, 0);
    out.cells = n;
    // ---- pit cells; tree over the other cells
    for (int j = 0; j < n; j++)
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) {
            if (g.up[e].task == UP_PARTNER) out.pair[j] = 1;
            else if (g.up[e].src >= j) return "unsupported: lagged inflow outside a pit pair";
        }
    std::vector<int32_t> parent(n, -1);
    for (int j = 0; j < n; j++) {
        if (out.pair[j]) { out.pair_cells++; continue; }
        const int d = g.down[j];
        if (d >= 0 && !out.pair[d]) parent[j] = d;
    }
    for (int j = 0; j < n; j++)
        if (out.pair[j] && g.down[j] >= 0 && !out.pair[g.down[j]]) return "unsupported: a pit pair drains into another cell";
    // ---- open subtrees
    std::vector<int32_t> osize(n, 1), oheight(n, 1);
    std::vector<uint8_t> closed(n, 0);
    std::vector<int32_t> kids;
    for (int j = 0; j < n; j++) {
        if (out.pair[j]) { closed[j] = 1; continue; }
        kids.clear();
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) kids.push_back(g.up[e].src);
        if (kids.size() > 8) return "unsupported: more than 8 inflows";
        std::stable_sort(kids.begin(), kids.end(), [&](int x, int y) { return g.flwacc[x] > g.flwacc[y]; });
        int size = 1, height = 1;
        for (int c : kids) {
            if (size + osize[c] <= 32 && oheight[c] + 1 <= dmax + 1) { size += osize[c]; height = std::max(height, oheight[c] + 1); }
            else closed[c] = 1;
        }
        osize[j] = size; oheight[j] = height;
        if (parent[j] < 0) closed[j] = 1;
    }
    // ---- components (root = closed cell), their levels
    std::vector<int32_t> comp(n, -1);
    for (int j = n - 1; j >= 0; j--) comp[j] = closed[j] ? j : comp[parent[j]];
    std::vector<int32_t> level(n, 0);                                     // per component root
    for (int c = 0; c < n; c++)                                           // ascending: a component's root is its largest cell
        if (closed[c] && !out.pair[c] && parent[c] >= 0) {
            const int cr = comp[parent[c]];
            level[cr] = std::max(level[cr], level[c] + 1);
        }
    // ---- tiles: first fit decreasing inside a level
    std::vector<int32_t> roots;
    for (int c = 0; c < n; c++) if (closed[c]) roots.push_back(c);
    out.ncomp = (int)roots.size();
    std::stable_sort(roots.begin(), roots.end(), [&](int x, int y) {
        if (level[x] != level[y]) return level[x] < level[y];
        return osize[x] > osize[y];
    });
    std::vector<int32_t> tile_of_comp(n, -1), tile_fill;
    {
        size_t i = 0;
        while (i < roots.size()) {
            size_t k = i;
            while (k < roots.size() && level[roots[k]] == level[roots[i]]) k++;
            const int first_tile = (int)tile_fill.size();
            // bins with free room, indexed by free lanes: a component goes to the fullest bin it fits
            std::vector<std::vector<int32_t>> by_free(33);
            for (size_t r = i; r < k; r++) {
                const int sz = osize[roots[r]];
                int t = -1;
                for (int f = sz; f <= 32 && t < 0; f++)
                    if (!by_free[f].empty()) { t = by_free[f].back(); by_free[f].pop_back(); }
                if (t < 0) { t = (int)tile_fill.size(); tile_fill.push_back(0); }
                tile_of_comp[roots[r]] = t;
                tile_fill[t] += sz;
                if (tile_fill[t] < 32) by_free[32 - tile_fill[t]].push_back(t);
            }
            (void)first_tile;
            out.nlevel = std::max(out.nlevel, level[roots[i]] + 1);
            i = k;
        }
    }
    out.ntile = (int)tile_fill.size();
    {
        // ---- dispatch order of the tiles: a topological order of the tile graph that starts the longest pipelines first
        // (list scheduling by the length of the longest chain of tiles downstream), so that the river tiles run under the
        // bulk of the domain instead of forming a tail.  A tile still only reads tiles with a smaller number.
        const int nt = out.ntile;
        std::vector<std::vector<int32_t>> cons(nt);
        std::vector<int32_t> indeg(nt, 0), dl(nt, 1), tlevel(nt, 0);
        for (int c = 0; c < n; c++)
            if (closed[c] && !out.pair[c] && parent[c] >= 0) {
                const int a = tile_of_comp[c], b2 = tile_of_comp[comp[parent[c]]];
                if (a == b2) return "internal: a tile reads itself";
                cons[a].push_back(b2);
            }
        for (int r : roots) tlevel[tile_of_comp[r]] = level[r];           // all components of a tile share the level
        for (int t = 0; t < nt; t++) {
            std::sort(cons[t].begin(), cons[t].end());
            cons[t].erase(std::unique(cons[t].begin(), cons[t].end()), cons[t].end());
            for (int b2 : cons[t]) indeg[b2]++;
        }
        std::vector<int32_t> bylevel(nt);
        for (int t = 0; t < nt; t++) bylevel[t] = t;
        std::stable_sort(bylevel.begin(), bylevel.end(), [&](int x, int y) { return tlevel[x] > tlevel[y]; });
        for (int t : bylevel)                                             // consumers (higher level) first
            for (int b2 : cons[t]) dl[t] = std::max(dl[t], dl[b2] + 1);
        std::vector<std::pair<int32_t, int32_t>> heap;                    // (downstream length, -tile): longest first, then the old order
        for (int t = 0; t < nt; t++) if (indeg[t] == 0) heap.push_back({dl[t], -t});
        std::make_heap(heap.begin(), heap.end());
        std::vector<int32_t> newid(nt, -1);
        int next_id = 0;
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end());
            const int t = -heap.back().second;
            heap.pop_back();
            newid[t] = next_id++;
            for (int b2 : cons[t])
                if (--indeg[b2] == 0) { heap.push_back({dl[b2], -b2}); std::push_heap(heap.begin(), heap.end()); }
        }
        if (next_id != nt) return "internal: the tile graph has a cycle";
        for (int r : roots) tile_of_comp[r] = newid[tile_of_comp[r]];
    }
    const size_t np2 = (size_t)out.ntile * 32;
    out.cell.assign(np2, -1); out.jprime.assign(n, -1); out.rec.assign(np2, 0); out.child.assign(np2 * 2, 0u);
    out.xout.assign(np2, -1); out.extoff.assign(np2, 0); out.tile_kmax.assign(out.ntile, 0); out.tile_ext.assign(out.ntile, 0);
    // ---- lanes: components in the order of `roots`, cells of a component in path order (descending so that parents come first
    // is not needed: a lane only needs its delay and the lanes of its inflows)
    // breadth first from the root of every component, so that the in-tile inflows of a cell sit in consecutive lanes, in the
    // reference's summation order: the warp sums them with a segmented scan
    std::vector<int32_t> next_lane(out.ntile, 0), delay(n, 0), segpos(n, 0), lastc(n, -1);
    {
        std::vector<int32_t> queue;
        for (int r : roots) {
            const int t = tile_of_comp[r];
            queue.clear();
            queue.push_back(r);
            for (size_t q = 0; q < queue.size(); q++) {
                const int j = queue[q];
                const int lane = next_lane[t]++;
                out.jprime[j] = t * 32 + lane;
                out.cell[(size_t)t * 32 + lane] = j;
                if (out.pair[j]) continue;
                int pos = 0;
                for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) {
                    const int c = g.up[e].src;
                    if (closed[c]) continue;
                    segpos[c] = pos++;
                    lastc[j] = c;
                    queue.push_back(c);
                }
            }
        }
    }
    for (int j = n - 1; j >= 0; j--) delay[j] = closed[j] ? oheight[j] - 1 : delay[parent[j]] - 1;
    // ---- exchange slots: one per closed root that flows into another component
    std::vector<int32_t> slot(n, -1);
    for (int c = 0; c < n; c++)
        if (closed[c] && !out.pair[c] && parent[c] >= 0) slot[c] = out.nslot++;
    for (int j = 0; j < n; j++) {
        const size_t jp = (size_t)out.jprime[j];
        const int t = (int)(jp >> 5);
        int rec = 1, nch = 0, next = 0;
        out.extoff[jp] = (int32_t)out.extlist.size();
        if (out.pair[j]) rec |= 2;
        else {
            if (delay[j] < 0 || delay[j] > dmax) return "internal: lane delay out of range";
            for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) {
                const int c = g.up[e].src;
                if (closed[c]) { out.extlist.push_back(slot[c]); next++; if (slot[c] < 0) return "internal: inflow without a slot"; }
                else {
                    const int cl = out.jprime[c] & 31;
                    if ((out.jprime[c] >> 5) != t || delay[c] != delay[j] - 1) return "internal: in-tile inflow off the wavefront";
                    if (nch < 6) out.child[jp * 2] |= (uint32_t)cl << (5 * nch);
                    else out.child[jp * 2 + 1] |= (uint32_t)cl << (5 * (nch - 6));
                    nch++;
                }
            }
            if (slot[j] >= 0) { rec |= 4; out.xout[jp] = slot[j]; }
            const int d = g.down[j];
            if (d >= 0 && out.pair[d]) rec |= 8;
            rec |= delay[j] << 8;
        }
        if (g.gauge_first[j] >= 0) rec |= 16;
        rec |= (nch << 12) | (next << 16) | (segpos[j] << 20) | ((lastc[j] >= 0 ? (out.jprime[lastc[j]] & 31) : 0) << 23);
        if (lastc[j] >= 0 && (out.jprime[lastc[j]] >> 5) != t) return "internal: inflow lanes outside the tile";
        out.rec[jp] = rec;
        out.tile_kmax[t] = std::max<uint8_t>(out.tile_kmax[t], (uint8_t)nch);
        if (next) out.tile_ext[t] = 1;
        out.kmax = std::max(out.kmax, nch);
    }
    if (out.extlist.empty()) out.extlist.push_back(0);
    // a tile only reads slots written by lower tiles
    for (int j = 0; j < n; j++)
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1]; e++) {
            const int c = g.up[e].src;
            if (!out.pair[j] && closed[c] && (out.jprime[c] >> 5) >= (out.jprime[j] >> 5)) return "internal: tile order violates a dependency";
        }
    return "";
}

}  // namespace smash
