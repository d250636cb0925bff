// Graph partitioner of the cell-face dual graph (new: the reference is single-process; north_star: "the mesh is graph-partitioned across
// the GPUs").  Multilevel recursive bisection, written from the published scheme (Karypis & Kumar 1998):
//   coarsening       heavy-edge matching, vertices visited by ascending degree, ties to the lighter partner, then a fixed scramble of
//                    the ids; two-hop matching of what is left
//   initial cut      greedy graph growing from ~128 seeds of the coarsest graph (<= COARSEN_TO vertices), each followed by a
//                    Fiduccia-Mattheyses refinement; the N_CANDIDATES best are uncoarsened to a level of >= SELECT_AT vertices, where
//                    the smallest cut wins (the cut of a few hundred blobs says little about the length of the cut it becomes)
//   uncoarsening     projection + boundary Fiduccia-Mattheyses passes per level (gain = external - internal edge weight, ties to the
//                    smaller id, roll-back to the best balanced prefix)
//   finest level     the two sides are EXACTLY n * (np/2) / np and the rest (the sizes recursive coordinate bisection produces, so the
//                    two partitioners are interchangeable as far as load balance goes)
// No coordinates are used.  Deterministic: no random numbers, every tie is broken by vertex id, and the two halves of a cut are
// independent subproblems (OpenMP tasks), so the partition does not depend on the number of threads that compute it - every rank
// of a job can compute it for itself and all agree (what mlb_create_partitioned / mlb_create_local require).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <numeric>
#include <set>
#include <stdexcept>
#include <utility>
#include <vector>

#include "mlb_internal.h"

namespace mlb {
namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;
double g_t[4] = {0, 0, 0, 0};      // MLB_PARTITION_DEBUG: seconds in coarsening / initial bisections / refinement / subgraph extraction (single-thread runs)
struct Tick { int k; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(); explicit Tick(int k_) : k(k_) {}
              ~Tick() { g_t[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); } };
constexpr uint32_t COARSEN_TO = 512;

struct Graph {
    uint32_t n = 0;
    std::vector<uint64_t> xadj;     // [n + 1]
    std::vector<uint32_t> adj, ew;  // neighbours, edge weights
    std::vector<uint32_t> vw;       // vertex weights
    uint64_t weight() const { return std::accumulate(vw.begin(), vw.end(), (uint64_t)0); }
};

// tie-breaker of the matching: a fixed scramble of the pair's ids.  "The smaller id" would give every tie of a regularly numbered mesh
// the same direction and the coarse vertices the shape of needles; a scramble keeps them round, and is as deterministic.
inline uint64_t mix(uint32_t a, uint32_t b) {
    uint64_t x = ((uint64_t)a << 32 | b) * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    return x;
}

// ---- coarsening --------------------------------------------------------------------------------------------------------
bool coarsen(const Graph & g, Graph & c, std::vector<uint32_t> & cmap, uint64_t max_vw) {
    std::vector<uint32_t> match(g.n, NONE);
    cmap.assign(g.n, NONE);
    // visiting order: ascending degree, then id (counting sort) - low-degree vertices have the fewest partners and choose first
    std::vector<uint32_t> order(g.n);
    {
        uint64_t max_deg = 0;
        for (uint32_t v = 0; v < g.n; v++) max_deg = std::max(max_deg, g.xadj[v + 1] - g.xadj[v]);
        const uint64_t cap = std::min<uint64_t>(max_deg, 4096);
        std::vector<uint32_t> count(cap + 2, 0);
        for (uint32_t v = 0; v < g.n; v++) count[std::min(g.xadj[v + 1] - g.xadj[v], cap) + 1]++;
        for (uint64_t d = 0; d <= cap; d++) count[d + 1] += count[d];
        for (uint32_t v = 0; v < g.n; v++) order[count[std::min(g.xadj[v + 1] - g.xadj[v], cap)]++] = v;
    }
    for (uint32_t k = 0; k < g.n; k++) {       // heavy-edge matching; equal weights: the lighter partner, then a fixed scramble of the ids
        const uint32_t v = order[k];
        if (match[v] != NONE) continue;
        uint32_t best = NONE, bw = 0;
        for (uint64_t e = g.xadj[v]; e < g.xadj[v + 1]; e++) {
            const uint32_t u = g.adj[e];
            if (u == v || match[u] != NONE || (uint64_t)g.vw[v] + g.vw[u] > max_vw) continue;
            if (best == NONE || g.ew[e] > bw || (g.ew[e] == bw && (g.vw[u] < g.vw[best] || (g.vw[u] == g.vw[best] && mix(u, v) < mix(best, v))))) { best = u; bw = g.ew[e]; }
        }
        if (best != NONE) { match[v] = best; match[best] = v; }
    }
    for (uint32_t k = 0; k < g.n; k++) {       // two-hop matching of what is left: two unmatched vertices with a common (matched) neighbour
        const uint32_t v = order[k];
        if (match[v] != NONE) continue;
        uint32_t best = NONE;
        for (uint64_t e = g.xadj[v]; e < g.xadj[v + 1] && best == NONE; e++) {
            const uint32_t u = g.adj[e];
            if (g.xadj[u + 1] - g.xadj[u] > 64) continue;
            for (uint64_t e2 = g.xadj[u]; e2 < g.xadj[u + 1]; e2++) {
                const uint32_t w = g.adj[e2];
                if (w != v && match[w] == NONE && (uint64_t)g.vw[v] + g.vw[w] <= max_vw && (best == NONE || w < best)) best = w;
            }
        }
        if (best != NONE) { match[v] = best; match[best] = v; }
        else match[v] = v;
    }
    std::vector<uint32_t> first, second;
    for (uint32_t v = 0; v < g.n; v++) {       // coarse ids in the order of the smaller member's id
        if (cmap[v] != NONE) continue;
        const uint32_t cv = (uint32_t)first.size();
        const uint32_t u = match[v];
        first.push_back(v);
        second.push_back(u == v ? NONE : u);
        cmap[v] = cv;
        if (u != v) cmap[u] = cv;
    }
    const uint32_t nc = (uint32_t)first.size();
    if ((double)nc > 0.95 * g.n) return false;
    c.n = nc;
    c.vw.assign(nc, 0);
    c.xadj.assign((size_t)nc + 1, 0);
    // rows of the coarse graph: the members' neighbours mapped to coarse ids, duplicates merged by a linear look through the row so
    // far (rows of a mesh graph hold a handful of entries).  Two passes - rows at an upper-bound offset, then compacted - so that the
    // rows are independent: taskloops, executed by whatever threads of the enclosing team are idle; the result does not depend on them.
    std::vector<uint64_t> ub((size_t)nc + 1, 0);
    for (uint32_t cv = 0; cv < nc; cv++) {
        uint64_t d = g.xadj[first[cv] + 1] - g.xadj[first[cv]];
        if (second[cv] != NONE) d += g.xadj[second[cv] + 1] - g.xadj[second[cv]];
        ub[cv + 1] = ub[cv] + d;
    }
    std::vector<uint32_t> tadj(ub[nc]), tew(ub[nc]), cnt(nc, 0);
#pragma omp taskloop grainsize(16384) shared(g, c, cmap, first, second, ub, tadj, tew, cnt)
    for (uint32_t cv = 0; cv < nc; cv++) {
        const uint64_t row = ub[cv];
        uint32_t k = 0, w = 0;
        const uint32_t mem[2] = {first[cv], second[cv]};
        for (int m = 0; m < 2; m++) {
            const uint32_t v = mem[m];
            if (v == NONE) continue;
            w += g.vw[v];
            for (uint64_t e = g.xadj[v]; e < g.xadj[v + 1]; e++) {
                const uint32_t cu = cmap[g.adj[e]];
                if (cu == cv) continue;
                uint32_t j = 0;
                while (j < k && tadj[row + j] != cu) j++;
                if (j < k) tew[row + j] += g.ew[e];
                else { tadj[row + k] = cu; tew[row + k] = g.ew[e]; k++; }
            }
        }
        c.vw[cv] = w;
        cnt[cv] = k;
    }
    for (uint32_t cv = 0; cv < nc; cv++) c.xadj[cv + 1] = c.xadj[cv] + cnt[cv];
    c.adj.resize(c.xadj[nc]); c.ew.resize(c.xadj[nc]);
#pragma omp taskloop grainsize(16384) shared(c, ub, tadj, tew, cnt)
    for (uint32_t cv = 0; cv < nc; cv++)
        for (uint32_t j = 0; j < cnt[cv]; j++) { c.adj[c.xadj[cv] + j] = tadj[ub[cv] + j]; c.ew[c.xadj[cv] + j] = tew[ub[cv] + j]; }
    return true;
}

// ---- two-way Fiduccia-Mattheyses refinement -------------------------------------------------------------------------------
// where[v] in {0, 1}; side 0 should weigh target0.  move_tol: how far side 0 may stray from target0 during a pass;
// accept_tol: how far it may be off in a state the pass may end in.  Returns the cut.
struct Refiner {
    const Graph & g;
    std::vector<uint8_t> & where;
    int64_t target0, move_tol, accept_tol;
    std::vector<int64_t> ed, id;         // external / internal edge weight of every vertex
    std::vector<uint8_t> locked;
    typedef std::pair<int64_t, uint32_t> Key;   // (-gain, vertex): begin() = largest gain, smallest id
    std::set<Key> bucket[2];
    int64_t w0 = 0, cut = 0;

    Refiner(const Graph & g_, std::vector<uint8_t> & w, int64_t t0, int64_t mt, int64_t at) : g(g_), where(w), target0(t0), move_tol(mt), accept_tol(at) {
        ed.assign(g.n, 0); id.assign(g.n, 0); locked.assign(g.n, 0);
        const uint32_t n = g.n;
        int64_t * edp = ed.data(), * idp = id.data();
        const Graph * gp = &g;
        const uint8_t * wp = where.data();
#pragma omp taskloop grainsize(32768) firstprivate(edp, idp, gp, wp)
        for (uint32_t v = 0; v < n; v++)
            for (uint64_t e = gp->xadj[v]; e < gp->xadj[v + 1]; e++) (wp[gp->adj[e]] == wp[v] ? idp[v] : edp[v]) += gp->ew[e];
        for (uint32_t v = 0; v < g.n; v++) {
            if (where[v] == 0) w0 += g.vw[v];
            cut += ed[v];
        }
        cut /= 2;
    }
    Key key(uint32_t v) const { return Key(-(ed[v] - id[v]), v); }
    int64_t off() const { return w0 > target0 ? w0 - target0 : target0 - w0; }

    void move(uint32_t v) {              // flips v, keeps ed / id / buckets of the unlocked neighbours current
        const uint8_t from = where[v];
        cut -= ed[v] - id[v];
        w0 += from == 0 ? -(int64_t)g.vw[v] : (int64_t)g.vw[v];
        where[v] = from ^ 1;
        std::swap(ed[v], id[v]);
        for (uint64_t e = g.xadj[v]; e < g.xadj[v + 1]; e++) {
            const uint32_t u = g.adj[e];
            if (u == v) continue;
            const bool tracked = !locked[u];
            if (tracked) bucket[where[u]].erase(key(u));
            if (where[u] == from) { id[u] -= g.ew[e]; ed[u] += g.ew[e]; }     // u stayed behind: the edge is now cut
            else { ed[u] -= g.ew[e]; id[u] += g.ew[e]; }
            if (tracked && ed[u] > 0) bucket[where[u]].insert(key(u));
        }
    }

    // one pass; returns true if the state improved
    bool pass() {
        bucket[0].clear(); bucket[1].clear();
        std::fill(locked.begin(), locked.end(), 0);
        for (uint32_t v = 0; v < g.n; v++) if (ed[v] > 0) bucket[where[v]].insert(key(v));
        std::vector<uint32_t> moves;
        const bool feas0 = off() <= accept_tol;
        bool best_feas = feas0;
        int64_t best_cut = cut, best_off = off();
        size_t best_at = 0;
        const size_t patience = std::min<size_t>(2000, std::max<size_t>(64, g.n / 100));
        bool all_in[2] = {false, false};
        while (true) {
            int from = -1;
            if (w0 > target0 + move_tol) from = 0;
            else if (w0 < target0 - move_tol) from = 1;
            const bool forced = from >= 0;
            if (forced && bucket[from].empty() && !all_in[from]) {     // (an isolated heavy side: any unlocked vertex may go)
                for (uint32_t v = 0; v < g.n; v++) if (where[v] == from && !locked[v]) bucket[from].insert(key(v));
                all_in[from] = true;
            }
            uint32_t v = NONE;
            if (forced) { if (!bucket[from].empty()) v = bucket[from].begin()->second; }
            else {
                // the better of the two sides' best candidates among those the balance admits
                Key cand[2]; bool has[2] = {false, false};
                int scanned[2] = {0, 0};
                for (int s = 0; s < 2; s++)
                    for (auto it = bucket[s].begin(); it != bucket[s].end(); ++it) {
                        const int64_t nw0 = w0 + (s == 0 ? -(int64_t)g.vw[it->second] : (int64_t)g.vw[it->second]);
                        if (nw0 >= target0 - move_tol && nw0 <= target0 + move_tol) { cand[s] = *it; has[s] = true; break; }
                        if (g.vw[it->second] == 1 || ++scanned[s] >= 8) break;   // unit weights: if the best does not fit, none does; else a short look
                    }
                if (has[0] && has[1]) {
                    // equal gains: move from the heavier side (towards balance), then the smaller id
                    if (cand[0].first != cand[1].first) from = cand[0].first < cand[1].first ? 0 : 1;
                    else from = w0 >= target0 ? 0 : 1;
                } else if (has[0]) from = 0;
                else if (has[1]) from = 1;
                if (from >= 0) v = cand[from].second;
            }
            if (v == NONE) break;
            bucket[where[v]].erase(key(v));
            locked[v] = 1;
            move(v);
            moves.push_back(v);
            const bool feas = off() <= accept_tol;
            const bool better = (feas && !best_feas) || (feas == best_feas && (cut < best_cut || (cut == best_cut && off() < best_off)));
            if (better) { best_feas = feas; best_cut = cut; best_off = off(); best_at = moves.size(); }
            else if (moves.size() - best_at > patience && best_feas) break;
        }
        for (size_t i = moves.size(); i > best_at; i--) {         // roll back (locked vertices: buckets are rebuilt by the next pass)
            const uint32_t v = moves[i - 1];
            locked[v] = 1;
            move(v);
        }
        return best_at > 0;
    }

    int64_t run(int max_passes) {
        for (int p = 0; p < max_passes; p++) {
            const int64_t c0 = cut, o0 = off();
            if (!pass()) break;
            if (cut == c0 && off() == o0) break;
        }
        return cut;
    }
};

// ---- initial bisection of the coarsest graph: greedy graph growing ---------------------------------------------------------
void grow_from(const Graph & g, uint32_t seed, int64_t target0, std::vector<uint8_t> & where) {
    where.assign(g.n, 1);
    std::vector<int64_t> gain(g.n, 0);          // (weight to side 0) - (weight to side 1) of a side-1 vertex
    std::vector<uint8_t> in_front(g.n, 0);
    std::set<std::pair<int64_t, uint32_t>> front;
    for (uint32_t v = 0; v < g.n; v++) for (uint64_t e = g.xadj[v]; e < g.xadj[v + 1]; e++) gain[v] -= g.ew[e];
    int64_t w0 = 0;
    uint32_t next_seed = seed, scanned = 0;
    while (w0 < target0) {
        uint32_t v = NONE;
        if (!front.empty()) {
            // best gain whose weight does not overshoot by more than it undershoots without it
            for (auto it = front.begin(); it != front.end(); ++it)
                if (w0 + (int64_t)g.vw[it->second] - target0 <= target0 - w0) { v = it->second; break; }
            if (v == NONE) break;
            front.erase(std::make_pair(-gain[v], v));
        } else {
            while (scanned < g.n && where[next_seed] == 0) { next_seed = (next_seed + 1) % g.n; scanned++; }
            if (scanned >= g.n) break;
            v = next_seed;
            if (w0 + (int64_t)g.vw[v] - target0 > target0 - w0 && w0 > 0) break;
        }
        where[v] = 0; in_front[v] = 0;
        w0 += g.vw[v];
        for (uint64_t e = g.xadj[v]; e < g.xadj[v + 1]; e++) {
            const uint32_t u = g.adj[e];
            if (u == v || where[u] == 0) continue;
            if (in_front[u]) front.erase(std::make_pair(-gain[u], u));
            gain[u] += 2 * (int64_t)g.ew[e];
            front.insert(std::make_pair(-gain[u], u));
            in_front[u] = 1;
        }
    }
}

// candidates: the N_CANDIDATES best (feasible first, then by cut, then by seed) of the grown-and-refined bisections
constexpr size_t N_CANDIDATES = 8;
constexpr uint32_t SELECT_AT = 16384;      // candidates are carried down to the first level with this many vertices; the smallest cut there goes on

struct Candidate { bool feas; int64_t cut, off; uint32_t seed; std::vector<uint8_t> where; };
bool better(const Candidate & a, const Candidate & b) {
    if (a.feas != b.feas) return a.feas;
    if (a.cut != b.cut) return a.cut < b.cut;
    if (a.off != b.off) return a.off < b.off;
    return a.seed < b.seed;
}

void initial_bisections(const Graph & g, int64_t target0, int64_t tol, std::vector<Candidate> & out) {
    out.clear();
    const uint32_t stride = std::max<uint32_t>(1, g.n / 128);      // (a graph that could not be coarsened further may still be large)
    for (uint32_t seed = 0; seed < g.n; seed += stride) {
        Candidate c;
        grow_from(g, seed, target0, c.where);
        Refiner r(g, c.where, target0, tol, tol);
        r.run(4);
        c.feas = r.off() <= tol; c.cut = r.cut; c.off = r.off(); c.seed = seed;
        bool duplicate = false;
        for (const Candidate & o : out) if (o.cut == c.cut && o.where == c.where) { duplicate = true; break; }
        if (duplicate) continue;
        out.push_back(std::move(c));
        std::sort(out.begin(), out.end(), better);
        if (out.size() > N_CANDIDATES) out.pop_back();
    }
}

// ---- multilevel bisection: where[v] = 0 for exactly n0 vertices of the (unit-weight) graph g ----------------------------------
void bisect(const Graph & g, uint32_t n0, std::vector<uint8_t> & where) {
    if (n0 == 0 || n0 >= g.n) { where.assign(g.n, n0 == 0 ? 1 : 0); return; }
    std::vector<Graph> levels;
    std::vector<std::vector<uint32_t>> cmaps;
    const Graph * cur = &g;
    const uint64_t total = g.n;
    const uint64_t max_vw = std::max<uint64_t>(1, (3 * total) / (2 * COARSEN_TO));
    while (cur->n > COARSEN_TO) {
        Graph c;
        std::vector<uint32_t> cmap;
        Tick tick(0);
        if (!coarsen(*cur, c, cmap, max_vw)) break;
        levels.push_back(std::move(c));
        cmaps.push_back(std::move(cmap));
        cur = &levels.back();
    }
    const double frac = (double)n0 / (double)g.n;
    const int64_t target_w = (int64_t)(frac * (double)total + 0.5);
    auto level_tol = [&](const Graph & lg) {   // the heaviest vertex, or 0.2 % of the graph: a coarse cut need not be exact, the finest is
        const uint32_t mx = lg.vw.empty() ? 1 : *std::max_element(lg.vw.begin(), lg.vw.end());
        return std::max<int64_t>((int64_t)mx, (int64_t)(0.002 * (double)total));
    };
    // one uncoarsening step: w on levels[l] -> the graph it was coarsened from, refined there; returns the cut
    auto step = [&](size_t l, std::vector<uint8_t> & w) {
        Tick tick(2);
        const Graph & fine = l == 0 ? g : levels[l - 1];
        std::vector<uint8_t> wf(fine.n);
        {
            uint8_t * wfp = wf.data();
            const uint8_t * wc = w.data();
            const uint32_t * cm = cmaps[l].data();
            const uint32_t nf = fine.n;
#pragma omp taskloop grainsize(65536) firstprivate(wfp, wc, cm)
            for (uint32_t v = 0; v < nf; v++) wfp[v] = wc[cm[v]];
        }
        w.swap(wf);
        const bool finest = l == 0;
        const int64_t tol = level_tol(fine);
        Refiner r(fine, w, finest ? (int64_t)n0 : target_w, finest ? std::max<int64_t>(2, tol / 4) : tol, finest ? 0 : tol);
        r.run(finest ? 10 : 6);
        if (getenv("MLB_PARTITION_DEBUG")) fprintf(stderr, "[mlb partition] level %zu: n %u cut %lld off %lld\n", l, fine.n, (long long)r.cut, (long long)r.off());
        return r.cut;
    };
    std::vector<Candidate> cands;
    { Tick tick(1);
    initial_bisections(*cur, target_w, cur == &g ? 0 : level_tol(*cur), cands); }
    std::vector<uint8_t> w;
    size_t l = levels.size();                  // levels still to be undone
    if (cur == &g) {                           // (a graph small enough not to be coarsened: exact sizes here and now)
        w.swap(cands.front().where);
        Refiner r(g, w, (int64_t)n0, std::max<int64_t>(1, (int64_t)(0.01 * g.n)), 0);
        r.run(8);
    } else {
        // every candidate is uncoarsened to the first level with >= SELECT_AT vertices (cheap: the levels above it hold a few per cent of
        // the vertices); the cut of a 100-vertex graph says little about the length of the cut it becomes, the cut at that level does
        size_t stop = l;
        while (stop > 0) { stop--; if ((stop == 0 ? g.n : levels[stop - 1].n) >= SELECT_AT) break; }
        int64_t best = -1;
        for (Candidate & c : cands) {
            int64_t cut = c.cut;
            for (size_t k = l; k-- > stop;) cut = step(k, c.where);
            if (best < 0 || cut < best) { best = cut; w = c.where; }
        }
        l = stop;
    }
    while (l-- > 0) step(l, w);
    where.swap(w);
    // the sizes are a contract (load balance): whatever the refinement left, make them exact with the cheapest moves
    int64_t n_side0 = 0;
    for (uint32_t v = 0; v < g.n; v++) n_side0 += where[v] == 0;
    if (n_side0 != (int64_t)n0) {
        Refiner r(g, where, (int64_t)n0, 0, 0);
        r.run(2);
        n_side0 = 0;
        for (uint32_t v = 0; v < g.n; v++) n_side0 += where[v] == 0;
        for (uint32_t v = 0; v < g.n && n_side0 != (int64_t)n0; v++) {      // last resort (cannot happen on a connected unit-weight graph)
            if (n_side0 > (int64_t)n0 && where[v] == 0) { where[v] = 1; n_side0--; }
            else if (n_side0 < (int64_t)n0 && where[v] == 1) { where[v] = 0; n_side0++; }
        }
    }
}

// ---- recursion -------------------------------------------------------------------------------------------------------------
void recurse(const Graph & g, const std::vector<uint32_t> & ids, int32_t p0, int32_t np, int32_t * part_out) {
    if (np == 1 || g.n == 0) { for (uint32_t v = 0; v < g.n; v++) part_out[ids[v]] = p0; return; }
    const int32_t npl = np / 2;
    const uint32_t n0 = (uint32_t)((double)g.n * npl / np);        // the sizes recursive coordinate bisection cuts (api.cu: rcb)
    std::vector<uint8_t> where;
    bisect(g, n0, where);
    Graph sub[2];
    std::vector<uint32_t> sub_ids[2], local(g.n);
    { Tick tick(3);
    for (uint32_t v = 0; v < g.n; v++) { local[v] = (uint32_t)sub_ids[where[v]].size(); sub_ids[where[v]].push_back(ids[v]); }
    for (int s = 0; s < 2; s++) { sub[s].n = (uint32_t)sub_ids[s].size(); sub[s].xadj.assign((size_t)sub[s].n + 1, 0); sub[s].vw.assign(sub[s].n, 1); }
    for (uint32_t v = 0; v < g.n; v++) {
        Graph & s = sub[where[v]];
        for (uint64_t e = g.xadj[v]; e < g.xadj[v + 1]; e++)
            if (where[g.adj[e]] == where[v] && g.adj[e] != v) { s.adj.push_back(local[g.adj[e]]); s.ew.push_back(1); }
        s.xadj[local[v] + 1] = s.adj.size();
    }
    }
#pragma omp task shared(sub, sub_ids) if (g.n > 100000)
    recurse(sub[0], sub_ids[0], p0, npl, part_out);
    recurse(sub[1], sub_ids[1], p0 + npl, np - npl, part_out);
#pragma omp taskwait
}

}  // namespace

void graph_partition(uint32_t n, const uint64_t * xadj, const uint32_t * adj, int32_t n_parts, int32_t * part_out) {
    Graph g;
    g.n = n;
    g.xadj.assign(xadj, xadj + (size_t)n + 1);
    g.adj.assign(adj, adj + xadj[n]);
    for (uint64_t e = 0; e < xadj[n]; e++) if (g.adj[e] >= n) throw std::runtime_error("mlb_partition_graph: neighbour id out of range");
    g.ew.assign(g.adj.size(), 1);
    g.vw.assign(n, 1);
    std::vector<uint32_t> ids(n);
    std::iota(ids.begin(), ids.end(), 0u);
#pragma omp parallel
#pragma omp single
    recurse(g, ids, 0, n_parts, part_out);
    if (getenv("MLB_PARTITION_DEBUG"))
        fprintf(stderr, "[mlb partition] seconds (summed over threads): coarsening %.2f, initial bisections %.2f, refinement %.2f, subgraphs %.2f\n", g_t[0], g_t[1], g_t[2], g_t[3]);
}

// dual graph of a mesh: cells are adjacent through their common faces (cells_of_face), in face order
void dual_graph(uint32_t n_cells, uint32_t n_faces, const int32_t * cells_of_face, std::vector<uint64_t> & xadj, std::vector<uint32_t> & adj) {
    xadj.assign((size_t)n_cells + 1, 0);
    for (uint32_t f = 0; f < n_faces; f++) {
        const int32_t a = cells_of_face[2 * (size_t)f], b = cells_of_face[2 * (size_t)f + 1];
        if (a >= 0 && b >= 0 && a != b) {
            if ((uint32_t)a >= n_cells || (uint32_t)b >= n_cells) throw std::runtime_error("mlb_partition_graph: cells_of_face out of range");
            xadj[(size_t)a + 1]++; xadj[(size_t)b + 1]++;
        }
    }
    for (uint32_t c = 0; c < n_cells; c++) xadj[c + 1] += xadj[c];
    adj.assign(xadj[n_cells], 0);
    std::vector<uint64_t> at(xadj.begin(), xadj.end() - 1);
    for (uint32_t f = 0; f < n_faces; f++) {
        const int32_t a = cells_of_face[2 * (size_t)f], b = cells_of_face[2 * (size_t)f + 1];
        if (a >= 0 && b >= 0 && a != b) { adj[at[a]++] = (uint32_t)b; adj[at[b]++] = (uint32_t)a; }
    }
}

}  // namespace mlb
