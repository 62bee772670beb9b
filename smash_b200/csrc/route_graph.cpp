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
// Classes of the window pass (window_kernels.cu) from the full route graph.  Cells are visited in path order, which is a
// topological order of every non-lagged edge: a cell is deep when its flow accumulation exceeds shallow_acc, when it sits in
// a pit pair, or when one of its inflows is deep or lagged; every other gathering cell is shallow and routed inside the
// window pass.
// ------------------------------------------------------------------------------------------------
std::string build_window_topo(const RouteGraph &g, int shallow_acc, WindowTopoHost &out) {
    const int n = g.n, npad = g.npad, ntile = npad / 32;
    out = WindowTopoHost();
    out.meta.assign(npad, 0); out.upoff.assign(npad, 0); out.deep.assign(npad, 0); out.tile_rounds.assign(ntile, 0);
    std::vector<uint8_t> cls(n, 0), round(n, 0);
    for (int j = 0; j < n; j++) {
        const int nup = g.up_begin[j + 1] - g.up_begin[j];
        if (g.flwacc[j] <= 1) { cls[j] = 0; continue; }
        // flwacc > 1 without a computed inflow: still routed (qup = 0, md_forward_structure.f90:146-156)
        bool deep = g.flwacc[j] > shallow_acc || nup > 8 || (g.flwacc[j] - 1) >= (1 << 19);
        int r = 0;
        for (int e = g.up_begin[j]; e < g.up_begin[j + 1] && !deep; e++) {
            const int s = g.up[e].src;
            if (g.up[e].task == UP_PARTNER || s >= j || cls[s] == 2) { deep = true; break; }
            if (cls[s] == 1 && (s >> 5) == (j >> 5)) r = std::max(r, round[s] + 1);
        }
        if (!deep && r > 7) deep = true;
        cls[j] = deep ? 2 : 1;
        round[j] = deep ? 0 : (uint8_t)r;
    }
    for (int j = 0; j < n; j++) {
        const int nup = (cls[j] == 1) ? g.up_begin[j + 1] - g.up_begin[j] : 0;
        const int d = g.down[j];
        const bool want_x = d >= 0 && cls[d] == 1, want_row = d >= 0 && cls[d] == 2 && cls[j] != 2;
        out.upoff[j] = (int32_t)out.ups.size();
        for (int e = 0; e < nup; e++) out.ups.push_back(g.up[g.up_begin[j] + e].src);
        out.meta[j] = cls[j] | (want_x ? 4 : 0) | (want_row ? 8 : 0) | (g.gauge_first[j] >= 0 ? 16 : 0) | (round[j] << 5) | (nup << 8) |
                      (std::min(g.flwacc[j] - 1, (1 << 19) - 1) << 12);
        out.deep[j] = cls[j] == 2;
        if (cls[j] == 1) {
            out.nshallow++;
            out.tile_rounds[j >> 5] = std::max<uint8_t>(out.tile_rounds[j >> 5], (uint8_t)(round[j] + 1));
            out.max_round = std::max(out.max_round, (int)round[j]);
        }
        if (cls[j] == 2) out.ndeep++;
        if (want_row) out.nrow++;
        if (want_x) out.nx_cells++;
    }
    if (out.ups.empty()) out.ups.push_back(-1);
    return "";
}

}  // namespace smash
