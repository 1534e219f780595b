// Decomposer::RecursiveAssembly, host side: the recombination plan of one connected component
// (fiksi/src/analyze/graph/recursive_assembly.rs:165-645, the Modified Frontier Algorithm as the reference runs it:
// breadth-first search for the first unblocked connected subgraph, frontier / core split, contraction of the core).
//
// The reference keeps its vertex, edge, subgraph and frontier sets in `hashbrown` hash sets and iterates them
// (:219,228,267,399,441,586,606); that order depends on the hasher's per-process seed.  Here every set is a bit
// set over element / constraint ids and is walked in ascending id order -- one of the orders the reference can
// take, and the one the oracle restates.  Steps carry the cluster bookkeeping as it was BEFORE the step (the
// reference clones its three maps into every step, :237-246,:314-321).
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <map>
#include <stdexcept>
#include <utility>
#include <vector>

namespace fk {

struct RaBits {
    std::vector<uint64_t> w;
    RaBits() = default;
    explicit RaBits(size_t n) : w((n + 63) / 64, 0) {}
    void grow(size_t n) { if (w.size() * 64 < n) w.resize((n + 63) / 64, 0); }
    bool test(uint32_t i) const { return (size_t)(i >> 6) < w.size() && ((w[i >> 6] >> (i & 63)) & 1u); }
    void set(uint32_t i) { grow((size_t)i + 1); w[i >> 6] |= 1ull << (i & 63); }
    void reset(uint32_t i) { if ((size_t)(i >> 6) < w.size()) w[i >> 6] &= ~(1ull << (i & 63)); }
    size_t count() const { size_t c = 0; for (uint64_t x : w) c += (size_t)__builtin_popcountll(x); return c; }
    bool same(const RaBits& o) const {
        const size_t n = std::max(w.size(), o.w.size());
        for (size_t k = 0; k < n; k++)
            if ((k < w.size() ? w[k] : 0) != (k < o.w.size() ? o.w[k] : 0)) return false;
        return true;
    }
    template <class F>
    void each(F&& f) const {  // ascending
        for (size_t k = 0; k < w.size(); k++)
            for (uint64_t x = w[k]; x; x &= x - 1) f((uint32_t)(k * 64 + (size_t)__builtin_ctzll(x)));
    }
};

// cluster bookkeeping: key -> list, insertion order of the lists preserved, keys walked in ascending order
using RaLists = std::map<uint32_t, std::vector<uint32_t>>;

struct RaStep {
    std::vector<uint32_t> constraints, elements, free_elements;
    RaLists on_frontiers;       // element -> clusters whose frontier it is on
    RaLists owned_elements;     // cluster -> elements it owns
    RaLists frontier_elements;  // cluster -> elements on its frontier
};

// The weighted constraint graph of graph.rs:98-117: vertices with degrees of freedom, hyper-edges with a valency.
struct RaGraph {
    std::vector<int> dof, valency;
    std::vector<std::vector<uint32_t>> vertex_edges, edge_vertices;
    uint32_t add_vertex(int d) {
        dof.push_back(d);
        vertex_edges.emplace_back();
        return (uint32_t)dof.size() - 1;
    }
    uint32_t add_edge(int val, const uint32_t* v, int n) {
        const uint32_t id = (uint32_t)valency.size();
        valency.push_back(val);
        edge_vertices.emplace_back(v, v + n);
        for (int k = 0; k < n; k++) vertex_edges[v[k]].push_back(id);
        return id;
    }
};

class RecursiveAssemblyPlanner {
public:
    explicit RecursiveAssemblyPlanner(RaGraph g) : g_(std::move(g)), real_vertices_((uint32_t)g_.dof.size()), real_edges_((uint32_t)g_.valency.size()) {}

    std::vector<RaStep> plan(const std::vector<uint32_t>& vertices, const std::vector<uint32_t>& edges, int D = 3) {
        RaBits live_v, live_e, done_v, done_e;
        for (uint32_t v : vertices) live_v.set(v);
        for (uint32_t e : edges) live_e.set(e);
        RaLists on_frontiers, owned, frontier_of;
        std::map<uint32_t, uint32_t> owner;
        std::vector<RaBits> blocked;
        std::vector<RaStep> out;
        std::vector<uint32_t> pending_constraints, pending_free;
        auto snapshot = [&](RaStep& st) {
            st.on_frontiers = on_frontiers;
            st.owned_elements = owned;
            st.frontier_elements = frontier_of;
        };
        for (uint32_t key = 0;; key++) {
            RaBits sub;
            if (!first_open_subgraph(live_v, live_e, blocked, D, sub)) {
                // nothing left to contract: whatever has not been solved yet goes into one last step (:209-251)
                RaStep st;
                live_e.each([&](uint32_t e) { if (e < real_edges_ && !done_e.test(e)) st.constraints.push_back(e); });
                live_v.each([&](uint32_t v) { if (v < real_vertices_ && !done_v.test(v)) st.free_elements.push_back(v); });
                if (!st.constraints.empty()) {
                    live_v.each([&](uint32_t v) { if (v < real_vertices_) st.elements.push_back(v); });
                    snapshot(st);
                    out.push_back(std::move(st));
                }
                break;
            }
            // frontier = vertices with a live edge leaving the subgraph; the rest is the core (:263-310)
            std::vector<uint32_t> core, real;
            RaBits frontier, core_bits;
            sub.each([&](uint32_t v) {
                if (v < real_vertices_) real.push_back(v);
                if (v < real_vertices_ && !done_v.test(v)) {
                    pending_free.push_back(v);
                    done_v.set(v);
                    owner[v] = key;
                }
                bool leaves = false;
                for (uint32_t e : g_.vertex_edges[v]) {
                    if (!live_e.test(e)) continue;
                    if (inside(e, sub)) {
                        if (e < real_edges_ && !done_e.test(e)) {
                            pending_constraints.push_back(e);
                            done_e.set(e);
                        }
                    } else {
                        leaves = true;
                    }
                }
                if (leaves) frontier.set(v);
                else { core.push_back(v); core_bits.set(v); }
            });
            if (!pending_constraints.empty()) {  // the step is emitted before the contraction (:312-322)
                RaStep st;
                st.constraints.swap(pending_constraints);
                st.elements = real;
                st.free_elements = pending_free;
                snapshot(st);
                out.push_back(std::move(st));
            }
            if (!core.empty() || !pending_free.empty()) {  // (:324-337)
                owned[key] = std::move(pending_free);
                pending_free.clear();
            }
            for (uint32_t v : core) {  // (:340-387)
                if (v < real_vertices_)
                    for (uint32_t e : g_.vertex_edges[v])
                        if (inside(e, core_bits)) live_e.reset(e);
                const uint32_t old = owner.at(v);
                owner[v] = key;
                if (old != key) {  // the cluster that owned v is merged into this one
                    std::vector<uint32_t> theirs = std::move(owned.at(old));
                    owned.erase(old);
                    for (uint32_t x : theirs) owner[x] = key;
                    std::vector<uint32_t>& mine = owned.at(key);
                    mine.insert(mine.end(), theirs.begin(), theirs.end());
                    std::vector<uint32_t> fr = std::move(frontier_of.at(old));
                    frontier_of.erase(old);
                    for (uint32_t x : fr) {
                        auto it = on_frontiers.find(x);
                        if (it == on_frontiers.end()) continue;
                        std::vector<uint32_t>& l = it->second;
                        const size_t at = (size_t)(std::find(l.begin(), l.end(), old) - l.begin());
                        if (at == l.size()) throw std::out_of_range("recursive assembly: frontier bookkeeping out of step");  // the reference panics (:377-381)
                        l[at] = l.back();  // Vec::swap_remove
                        l.pop_back();
                    }
                }
                on_frontiers.erase(v);
            }
            frontier.each([&](uint32_t v) {  // (:388-397)
                on_frontiers[v].push_back(key);
                if (v < real_vertices_) frontier_of[key].push_back(v);
            });
            if (sub.count() - frontier.count() <= 1) {  // nothing to contract: never offer this subgraph again (:409-419)
                blocked.push_back(sub);
                continue;
            }
            // contraction: the core becomes one vertex, its edges to a frontier vertex one edge (:421-480)
            for (uint32_t v : core) live_v.reset(v);
            const uint32_t cv = g_.add_vertex(0);
            owner[cv] = key;
            live_v.set(cv);
            int frontier_dof = 0, incoming = 0;
            frontier.each([&](uint32_t v) {
                frontier_dof += g_.dof[v];
                int binary = 0;
                const std::vector<uint32_t> edges_of_v = g_.vertex_edges[v];
                for (uint32_t e : edges_of_v) {
                    if (!live_e.test(e) || !inside(e, sub)) continue;
                    std::vector<uint32_t> merged;  // graph.rs:58-93: non-frontier ends collapse into cv
                    bool collapsed = false;
                    for (uint32_t x : g_.edge_vertices[e]) {
                        if (frontier.test(x)) merged.push_back(x);
                        else if (!collapsed) { merged.push_back(cv); collapsed = true; }
                    }
                    if (merged.size() == 2) {
                        binary += g_.valency[e];
                        live_e.reset(e);
                    } else {
                        g_.edge_vertices[e] = std::move(merged);
                    }
                }
                if (binary > 0) {
                    const uint32_t ends[2] = {v, cv};
                    live_e.set(g_.add_edge(binary, ends, 2));
                    incoming += binary;
                }
            });
            if (incoming > 0) g_.dof[cv] = frontier_dof - incoming - D;
            else live_v.reset(cv);
        }
        return out;
    }

private:
    RaGraph g_;
    uint32_t real_vertices_, real_edges_;

    bool inside(uint32_t e, const RaBits& set) const {
        for (uint32_t x : g_.edge_vertices[e])
            if (!set.test(x)) return false;
        return true;
    }
    // recursive_assembly.rs:499-645: breadth-first over connected vertex sets (grown one adjacent vertex at a time, seeds and
    // extensions in ascending order); the first set of two or more vertices with dof > -(D + 1) that is not blocked.
    bool first_open_subgraph(const RaBits& live_v, const RaBits& live_e, const std::vector<RaBits>& blocked, int D, RaBits& found) const {
        struct State { RaBits set, next; int dof; };
        auto neighbours = [&](uint32_t v, const RaBits& set, RaBits& next) {
            for (uint32_t e : g_.vertex_edges[v]) {
                if (!live_e.test(e)) continue;
                for (uint32_t x : g_.edge_vertices[e])
                    if (live_v.test(x) && !set.test(x)) next.set(x);
            }
        };
        std::deque<State> queue;
        live_v.each([&](uint32_t v) {
            State s;
            s.set.set(v);
            neighbours(v, s.set, s.next);
            s.dof = g_.dof[v];
            queue.push_back(std::move(s));
        });
        bool hit = false;
        while (!queue.empty() && !hit) {
            State cur = std::move(queue.front());
            queue.pop_front();
            cur.next.each([&](uint32_t v) {
                if (hit) return;
                RaBits grown = cur.set;
                grown.set(v);
                int closed = 0;  // valency of the live edges that v closes inside the grown set
                for (uint32_t e : g_.vertex_edges[v])
                    if (live_e.test(e) && inside(e, grown)) closed += g_.valency[e];
                const int dof = cur.dof + g_.dof[v] - closed;
                bool is_blocked = false;
                for (const RaBits& b : blocked) is_blocked = is_blocked || b.same(grown);
                if (!is_blocked && dof > -(D + 1)) {
                    found = std::move(grown);
                    hit = true;
                    return;
                }
                State nx;
                nx.next = cur.next;
                nx.next.reset(v);
                neighbours(v, grown, nx.next);
                nx.set = std::move(grown);
                nx.dof = dof;
                queue.push_back(std::move(nx));
            });
        }
        return hit;
    }
};

}  // namespace fk
