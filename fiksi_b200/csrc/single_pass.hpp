// Decomposer::SinglePass on the host side of the solve boundary: the equation graph
// of fiksi/src/lib.rs:262,395,434, a maximum matching of free variables to expressions
// (Hopcroft-Karp, fiksi/src/analyze/graph/equations.rs:297-403), and the strongly connected sets of
// expressions of the matched graph (Pearce's variant of Tarjan's algorithm, :458-567) in the order
// assemble::solve consumes them (:186-221, assemble/mod.rs:169-210).
//
// Dense-id implementation: variables and expressions are small consecutive integers, so the
// reference's IndexMaps become arrays plus one insertion-order list (the only place where the
// reference's iteration order is observable: Tarjan's roots are tried in the order expressions were
// first matched).  One deliberate difference: the reference hands out an SCC's free variables in
// HashSet order (unspecified); here they are ascending, which is also what fk_problem requires.
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <vector>

namespace fk {

struct SinglePassStep {
    std::vector<uint32_t> free_variables;  // ascending
    std::vector<uint32_t> expressions;     // in the order the SCC search emitted them (row order)
};

class SinglePassPlanner {
public:
    // var_exprs[v]: expressions of variable v in creation order; expr_vars[e]: the slots of expression e.
    SinglePassPlanner(const std::vector<std::vector<uint32_t>>& var_exprs, const std::vector<std::vector<uint32_t>>& expr_vars)
        : ve_(var_exprs), ev_(expr_vars) {}

    // `free_sorted`: ascending free variables of one connected component.
    std::vector<SinglePassStep> plan(const std::vector<uint32_t>& free_sorted) {
        const uint32_t nv = (uint32_t)ve_.size(), ne = (uint32_t)ev_.size();
        is_free_.assign(nv, 0);
        for (uint32_t v : free_sorted) is_free_[v] = 1;
        var_match_.assign(nv, kNone);
        expr_match_.assign(ne, kNone);
        matched_order_.clear();
        dist_.assign(nv, kInf);
        // Hopcroft-Karp: phases of one BFS (layering from the unmatched variables) and one DFS per
        // still unmatched variable, in ascending variable order
        while (bfs(free_sorted))
            for (uint32_t a : free_sorted)
                if (var_match_[a] == kNone) dfs(a);
        // strongly connected expressions; roots in first-matched order
        index_ = 1;
        comp_ = (uint32_t)matched_order_.size() - 1u;  // wraps for an empty matching, like the reference
        root_.assign(ne, 0);
        seen_.assign(ne, 0);
        stack_.clear();
        sccs_.clear();
        for (uint32_t e : matched_order_)
            if (!seen_[e]) visit(e);
        std::vector<SinglePassStep> out;
        for (size_t k = sccs_.size(); k-- > 0;) {  // reverse: topological order of the condensation
            SinglePassStep st;
            st.expressions = sccs_[k];
            for (uint32_t e : st.expressions) {
                const uint32_t mv = expr_match_[e];
                for (uint32_t var : ev_[e])
                    if (var == mv || (var_match_[var] == kNone && is_free_[var])) st.free_variables.push_back(var);
            }
            std::sort(st.free_variables.begin(), st.free_variables.end());
            st.free_variables.erase(std::unique(st.free_variables.begin(), st.free_variables.end()), st.free_variables.end());
            out.push_back(std::move(st));
        }
        return out;
    }

private:
    static constexpr uint32_t kNone = 0xFFFFFFFFu, kInf = 0xFFFFFFFFu;
    const std::vector<std::vector<uint32_t>>& ve_;
    const std::vector<std::vector<uint32_t>>& ev_;
    std::vector<uint8_t> is_free_, seen_;
    std::vector<uint32_t> var_match_, expr_match_, matched_order_, dist_, root_, stack_;
    std::vector<std::vector<uint32_t>> sccs_;
    uint32_t dummy_ = kInf, index_ = 1, comp_ = 0;

    static uint32_t inc(uint32_t d) { return d == kInf ? d : d + 1; }

    void match(uint32_t a, uint32_t b) {
        var_match_[a] = b;
        if (expr_match_[b] == kNone) matched_order_.push_back(b);  // first insertion fixes the position
        expr_match_[b] = a;
    }
    bool bfs(const std::vector<uint32_t>& free_sorted) {
        std::deque<uint32_t> queue;
        for (uint32_t a : free_sorted) {
            if (var_match_[a] != kNone) dist_[a] = kInf;
            else { dist_[a] = 0; queue.push_back(a); }
        }
        dummy_ = kInf;
        while (!queue.empty()) {
            const uint32_t a = queue.front();
            queue.pop_front();
            if (dist_[a] >= dummy_) continue;
            const uint32_t nd = inc(dist_[a]);
            for (uint32_t b : ve_[a]) {
                const uint32_t ma = expr_match_[b];
                if (ma == kNone) {
                    if (dummy_ == kInf) dummy_ = nd;
                } else if (dist_[ma] == kInf) {
                    dist_[ma] = nd;
                    queue.push_back(ma);
                }
            }
        }
        return dummy_ != kInf;
    }
    bool dfs(uint32_t a) {
        const uint32_t want = inc(dist_[a]);
        for (uint32_t b : ve_[a]) {
            const uint32_t ma = expr_match_[b];
            if (ma == kNone) {
                if (dummy_ == want) { match(a, b); return true; }
            } else if (dist_[ma] == want && dfs(ma)) {
                match(a, b);
                return true;
            }
        }
        dist_[a] = kInf;
        return false;
    }
    template <class F>
    void for_each_neighbor(uint32_t e, F&& f) const {
        const uint32_t mv = expr_match_[e];
        for (uint32_t a : ev_[e]) {
            if (!is_free_[a]) continue;
            if (!(a == mv || var_match_[a] == kNone)) continue;
            for (uint32_t b : ve_[a])
                if (b != e && expr_match_[b] != kNone) f(b);
        }
    }
    void visit(uint32_t v) {
        bool is_root = true;
        uint32_t vi = index_;
        root_[v] = vi;
        seen_[v] = 1;
        index_ += 1;
        for_each_neighbor(v, [&](uint32_t nb) {
            if (!seen_[nb]) visit(nb);
            if (root_[nb] < vi) {
                vi = root_[nb];
                root_[v] = vi;
                is_root = false;
            }
        });
        if (is_root) {
            std::vector<uint32_t> scc{v};
            index_ -= 1;
            while (!stack_.empty() && !(vi > root_[stack_.back()])) {
                const uint32_t w = stack_.back();
                stack_.pop_back();
                scc.push_back(w);
                root_[w] = comp_;
                index_ -= 1;
            }
            root_[v] = comp_;
            comp_ -= 1;
            sccs_.push_back(std::move(scc));
        } else {
            stack_.push_back(v);
        }
    }
};

}  // namespace fk
