// fk_system: the host-side mirror of fiksi::System (fiksi/src/lib.rs:252-467) above the solve
// boundary — element / constraint creation, the incidence graph with its incremental connected
// components, fix / unfix, and assemble::solve's scale, perturbation and write-back
// (fiksi/src/assemble/mod.rs:32-167).  All numerics (residuals, Jacobians, ordering,
// factorisation, LM) happen behind fk_lm_solve_batch / the evaluation kernels on the GPU.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fiksi_b200.h"
#include "recursive_assembly.hpp"
#include "single_pass.hpp"
#include "symbolic.hpp"

namespace {

enum ElemTag : uint8_t { kLength = 0, kPoint = 1, kLine = 2, kCircle = 3 };

struct Elem {
    uint8_t tag;
    uint32_t a, b;  // Length/Point: a = first variable; Line: a, b = point variables; Circle: a = centre var, b = radius var
};
struct Constr {
    uint8_t tag;         // ConstraintTag order (constraints/mod.rs:893-905)
    uint32_t first_expr; // index of its first expression
    uint8_t n_expr;      // valency: 2 for PointPointCoincidence, else 1 (constraints/mod.rs:992-1034)
    uint8_t n_inc;       // incident primitive elements (graph.rs:112-117), in the order the reference lists them
    uint32_t inc[4];
};

// Knuth/Lewis LCG of fiksi/src/rand.rs:24-39.
struct Lcg {
    uint32_t s;
    double next() {
        s = s * 1664525u + 1013904223u;
        return (1.0 / 4294967295.0) * (double)s;
    }
};

}  // namespace

struct fk_system {
    std::vector<Elem> elems;
    std::vector<double> vars;
    std::vector<uint32_t> var_owner;  // variable -> primitive element (lib.rs variable_to_primitive)
    std::vector<uint8_t> fixed;       // per variable
    std::vector<Constr> constrs;
    std::vector<uint8_t> kind;
    std::vector<uint32_t> idx;        // 4 per expression
    std::vector<double> param;
    // incidence graph: incremental connected components exactly as graph.rs:178-225 builds them —
    // only the elements incident to the new constraint are re-pointed on a merge, members of a
    // merged-away component keep their old (now empty) slot.
    std::vector<uint32_t> comp_of;    // per element, 0 = none, else 1-based slot
    struct Comp { std::set<uint32_t> elements, constraints; };
    std::vector<Comp> comps;
    std::string error;
    std::unique_ptr<fk_topology, void (*)(fk_topology*)> eval_topo{nullptr, fk_topology_destroy};

    uint32_t new_element(uint8_t tag, uint32_t a, uint32_t b, const double* v, int nv) {
        uint32_t id = (uint32_t)elems.size();
        uint32_t first = (uint32_t)vars.size();
        for (int k = 0; k < nv; k++) {
            vars.push_back(v[k]);
            var_owner.push_back(id);
            fixed.push_back(0);
        }
        elems.push_back({tag, nv ? first : a, b});
        comp_of.push_back(0);
        eval_topo.reset();
        return id;
    }
    int element_vars(uint32_t e, uint32_t out[4]) const {
        const Elem& el = elems[e];
        switch (el.tag) {
            case kLength: out[0] = el.a; return 1;
            case kPoint: out[0] = el.a; out[1] = el.a + 1; return 2;
            case kLine: out[0] = el.a; out[1] = el.a + 1; out[2] = el.b; out[3] = el.b + 1; return 4;
            default: out[0] = el.a; out[1] = el.a + 1; out[2] = el.b; return 3;
        }
    }
    void connect(uint32_t constraint, const uint32_t* inc, int n) {
        uint32_t target = 0;
        size_t best = 0;
        for (int k = 0; k < n; k++) {
            uint32_t c = comp_of[inc[k]];
            if (c && comps[c - 1].elements.size() > best) {
                best = comps[c - 1].elements.size();
                target = c;
            }
        }
        if (!target) {
            comps.emplace_back();
            target = (uint32_t)comps.size();
        }
        Comp merged;
        merged.elements.swap(comps[target - 1].elements);
        merged.constraints.swap(comps[target - 1].constraints);
        for (int k = 0; k < n; k++) {
            uint32_t c = comp_of[inc[k]];
            if (c) {
                Comp& src = comps[c - 1];  // already emptied if it is the target (or a stale slot)
                merged.elements.insert(src.elements.begin(), src.elements.end());
                merged.constraints.insert(src.constraints.begin(), src.constraints.end());
                src.elements.clear();
                src.constraints.clear();
            } else {
                merged.elements.insert(inc[k]);
            }
            comp_of[inc[k]] = target;
        }
        merged.constraints.insert(constraint);
        comps[target - 1].elements.swap(merged.elements);
        comps[target - 1].constraints.swap(merged.constraints);
    }
    void push_expr(uint8_t k, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3, double p) {
        kind.push_back(k);
        idx.push_back(i0); idx.push_back(i1); idx.push_back(i2); idx.push_back(i3);
        param.push_back(p);
    }
};

extern "C" {

int fk_system_create(fk_system** out) {
    if (!out) return FK_ERR_INVALID;
    *out = new (std::nothrow) fk_system();
    return *out ? FK_OK : FK_ERR_OOM;
}
void fk_system_destroy(fk_system* s) { delete s; }

// elements/mod.rs:280-454
uint32_t fk_system_add_length(fk_system* s, double length) { return s->new_element(kLength, 0, 0, &length, 1); }
uint32_t fk_system_add_point(fk_system* s, double x, double y) {
    double v[2] = {x, y};
    return s->new_element(kPoint, 0, 0, v, 2);
}
uint32_t fk_system_add_line(fk_system* s, uint32_t p1, uint32_t p2) {
    if (p1 >= s->elems.size() || p2 >= s->elems.size() || s->elems[p1].tag != kPoint || s->elems[p2].tag != kPoint) return UINT32_MAX;
    return s->new_element(kLine, s->elems[p1].a, s->elems[p2].a, nullptr, 0);
}
uint32_t fk_system_add_circle(fk_system* s, uint32_t center, uint32_t radius) {
    if (center >= s->elems.size() || radius >= s->elems.size() || s->elems[center].tag != kPoint || s->elems[radius].tag != kLength) return UINT32_MAX;
    return s->new_element(kCircle, s->elems[center].a, s->elems[radius].a, nullptr, 0);
}

// elements/mod.rs:60-86
int fk_system_fix(fk_system* s, uint32_t element, int fix) {
    if (!s || element >= s->elems.size()) return FK_ERR_INVALID;
    uint32_t v[4];
    int n = s->element_vars(element, v);
    for (int k = 0; k < n; k++) s->fixed[v[k]] = fix ? 1 : 0;
    return FK_OK;
}

// constraints/mod.rs:317-891.  `tag` in ConstraintTag order; `elements` are the handles the
// reference's constructor takes (e.g. {point, line} for PointLineIncidence).
uint32_t fk_system_add_constraint(fk_system* s, int tag, const uint32_t* el, uint32_t n_el, double p) {
    static const uint8_t want[11][4] = {
        {kPoint, kPoint, 255, 255}, {kPoint, kPoint, 255, 255}, {kPoint, kPoint, kPoint, 255}, {kPoint, kLine, 255, 255},
        {kPoint, kLine, 255, 255},  {kPoint, kCircle, 255, 255}, {kPoint, kPoint, kPoint, kPoint}, {kLine, kLine, 255, 255},
        {kLine, kLine, 255, 255},   {kLine, kLine, 255, 255},   {kLine, kCircle, 255, 255}};
    if (!s || tag < 0 || tag > 10 || !el) return UINT32_MAX;
    uint32_t need = 0;
    while (need < 4 && want[tag][need] != 255) need++;
    if (n_el != need) return UINT32_MAX;
    for (uint32_t k = 0; k < need; k++)
        if (el[k] >= s->elems.size() || s->elems[el[k]].tag != want[tag][k]) return UINT32_MAX;
    auto var = [&](uint32_t e) { return s->elems[e].a; };
    auto own = [&](uint32_t v) { return s->var_owner[v]; };
    const uint32_t id = (uint32_t)s->constrs.size();
    const uint32_t first = (uint32_t)s->kind.size();
    uint32_t inc[4];
    int n_inc = 0;
    uint8_t n_expr = 1;
    switch (tag) {
        case 0: {  // PointPointCoincidence: x and y equalities
            inc[0] = el[0]; inc[1] = el[1]; n_inc = 2; n_expr = 2;
            s->push_expr(FK_VARIABLE_VARIABLE_EQUALITY, var(el[0]), var(el[1]), 0, 0, 0.0);
            s->push_expr(FK_VARIABLE_VARIABLE_EQUALITY, var(el[0]) + 1, var(el[1]) + 1, 0, 0, 0.0);
        } break;
        case 1:
            inc[0] = el[0]; inc[1] = el[1]; n_inc = 2;
            s->push_expr(FK_POINT_POINT_DISTANCE, var(el[0]), var(el[1]), 0, 0, p);
            break;
        case 2:
            inc[0] = el[0]; inc[1] = el[1]; inc[2] = el[2]; n_inc = 3;
            s->push_expr(FK_POINT_POINT_POINT_ANGLE, var(el[0]), var(el[1]), var(el[2]), 0, p);
            break;
        case 3: case 4: {
            const Elem& l = s->elems[el[1]];
            inc[0] = el[0]; inc[1] = own(l.a); inc[2] = own(l.b); n_inc = 3;
            s->push_expr(tag == 3 ? FK_POINT_LINE_INCIDENCE : FK_POINT_LINE_DISTANCE, var(el[0]), l.a, l.b, 0, tag == 4 ? p : 0.0);
        } break;
        case 5: {
            const Elem& c = s->elems[el[1]];
            inc[0] = el[0]; inc[1] = own(c.a); inc[2] = own(c.b); n_inc = 3;
            s->push_expr(FK_POINT_CIRCLE_INCIDENCE, var(el[0]), c.a, c.b, 0, 0.0);
        } break;
        case 6:
            for (int k = 0; k < 4; k++) inc[k] = own(var(el[k]));
            n_inc = 4;
            s->push_expr(FK_SEGMENT_SEGMENT_LENGTH_EQUALITY, var(el[0]), var(el[1]), var(el[2]), var(el[3]), 0.0);
            break;
        case 7: case 8: case 9: {
            const Elem& l1 = s->elems[el[0]];
            const Elem& l2 = s->elems[el[1]];
            inc[0] = own(l1.a); inc[1] = own(l1.b); inc[2] = own(l2.a); inc[3] = own(l2.b); n_inc = 4;
            const uint8_t k = tag == 7 ? FK_LINE_LINE_ANGLE : (tag == 8 ? FK_LINE_LINE_PARALLELISM : FK_LINE_LINE_PERPENDICULARITY);
            s->push_expr(k, l1.a, l1.b, l2.a, l2.b, tag == 7 ? p : 0.0);
        } break;
        default: {  // 10 LineCircleTangency
            const Elem& l = s->elems[el[0]];
            const Elem& c = s->elems[el[1]];
            inc[0] = own(l.a); inc[1] = own(l.b); inc[2] = own(c.a); inc[3] = own(c.b); n_inc = 4;
            s->push_expr(FK_LINE_CIRCLE_TANGENCY, l.a, l.b, c.a, c.b, 0.0);
        } break;
    }
    s->connect(id, inc, n_inc);
    Constr rec{(uint8_t)tag, first, n_expr, (uint8_t)n_inc, {0, 0, 0, 0}};
    for (int k = 0; k < n_inc; k++) rec.inc[k] = inc[k];
    s->constrs.push_back(rec);
    s->eval_topo.reset();
    return id;
}

uint32_t fk_system_num_variables(const fk_system* s) { return s ? (uint32_t)s->vars.size() : 0; }
uint32_t fk_system_num_constraints(const fk_system* s) { return s ? (uint32_t)s->constrs.size() : 0; }
uint32_t fk_system_element_variable(const fk_system* s, uint32_t e) { return (s && e < s->elems.size()) ? s->elems[e].a : UINT32_MAX; }
int fk_system_get_variables(const fk_system* s, double* out) {
    if (!s || !out) return FK_ERR_INVALID;
    std::memcpy(out, s->vars.data(), s->vars.size() * sizeof(double));
    return FK_OK;
}
// elements/mod.rs:558-579 (update_value) for any variable
int fk_system_set_variable(fk_system* s, uint32_t var, double value) {
    if (!s || var >= s->vars.size()) return FK_ERR_INVALID;
    s->vars[var] = value;
    return FK_OK;
}
// constraints/mod.rs:992-1046 (update_parameter)
int fk_system_set_parameter(fk_system* s, uint32_t constraint, double value) {
    if (!s || constraint >= s->constrs.size()) return FK_ERR_INVALID;
    s->param[s->constrs[constraint].first_expr] = value;
    return FK_OK;
}

// == System::solve(SolvingOptions { optimizer: LevenbergMarquardt, decomposer: None, perturb })
// (lib.rs:464, assemble/mod.rs:46-167).  reports: one per solved component in component order, up
// to `cap`; *n_solved receives the number of components solved.
int fk_system_solve(fk_system* s, int perturb, fk_report* reports, uint32_t cap, uint32_t* n_solved) {
    if (!s) return FK_ERR_INVALID;
    if (n_solved) *n_solved = 0;
    const size_t nv = s->vars.size();
    // assemble/mod.rs:32-44,58-79: RMS of all variables and of the distance parameters (sequential sums)
    double sum = 0.0;
    size_t cnt = 0;
    for (double v : s->vars) { sum += v * v; cnt++; }
    for (size_t e = 0; e < s->kind.size(); e++)
        if (s->kind[e] == FK_POINT_POINT_DISTANCE || s->kind[e] == FK_POINT_LINE_DISTANCE) { sum += s->param[e] * s->param[e]; cnt++; }
    const double scale = std::sqrt(sum / (double)cnt);
    const double recip = 1.0 / scale;
    std::vector<double> vt(nv);
    for (size_t i = 0; i < nv; i++) vt[i] = s->vars[i] * recip;
    std::vector<double> pt(s->param);
    for (size_t e = 0; e < pt.size(); e++)
        if (s->kind[e] == FK_POINT_POINT_DISTANCE || s->kind[e] == FK_POINT_LINE_DISTANCE) pt[e] = recip * s->param[e];

    // One compact problem per non-empty component: the referenced variables are renumbered
    // locally (ascending, so x/y of a point stay adjacent), which also lets components of the same
    // shape share one topology inside fk_lm_solve_batch.
    struct Local {
        std::vector<double> vars, param, x;
        std::vector<uint8_t> kind;
        std::vector<uint32_t> idx, free_local, rows, free_global;
        fk_problem prob;
    };
    std::vector<std::unique_ptr<Local>> locals;
    Lcg rng{42};  // assemble/mod.rs:47: one generator per solve, consumed by the components in order
    for (const fk_system::Comp& c : s->comps) {
        if (c.elements.empty()) continue;  // assemble/mod.rs:87-89
        std::set<uint32_t> free_set;
        for (uint32_t e : c.elements) {
            uint32_t v[4];
            int n = s->element_vars(e, v);
            for (int k = 0; k < n; k++)
                if (!s->fixed[v[k]]) free_set.insert(v[k]);
        }
        if (perturb) {  // assemble/mod.rs:113-124
            for (uint32_t fv : free_set) {
                const double r1 = rng.next();
                const double r2 = rng.next();
                vt[fv] += vt[fv] * (1.0 / 8196.0) * r1 + (1.0 / 65568.0) * r2;
            }
        }
        std::unique_ptr<Local> L(new Local());
        std::vector<uint32_t> exprs;
        for (uint32_t ci : c.constraints)
            for (uint8_t k = 0; k < s->constrs[ci].n_expr; k++) exprs.push_back(s->constrs[ci].first_expr + k);
        std::set<uint32_t> used(free_set.begin(), free_set.end());
        for (uint32_t e : exprs) {
            uint32_t sv[8];
            int a = fk::expand_slots(s->kind[e], &s->idx[4 * (size_t)e], sv);
            for (int k = 0; k < a; k++) used.insert(sv[k]);
        }
        std::map<uint32_t, uint32_t> local;
        for (uint32_t g : used) {
            local.emplace(g, (uint32_t)L->vars.size());
            L->vars.push_back(vt[g]);  // values as of now: later components have not been perturbed yet
        }
        for (uint32_t g : free_set) {
            L->free_global.push_back(g);
            L->free_local.push_back(local[g]);
            L->x.push_back(vt[g]);
        }
        static const int stored[FK_NUM_KINDS] = {2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 3, 3};
        for (uint32_t e : exprs) {
            L->rows.push_back((uint32_t)L->kind.size());
            L->kind.push_back(s->kind[e]);
            L->param.push_back(pt[e]);
            for (int q = 0; q < 4; q++) L->idx.push_back(q < stored[s->kind[e]] ? local[s->idx[4 * (size_t)e + q]] : 0);
        }
        fk_problem& p = L->prob;
        p.n_vars = (uint32_t)L->vars.size(); p.vars = L->vars.data();
        p.n_expr = (uint32_t)L->kind.size(); p.kind = L->kind.data(); p.idx = L->idx.data(); p.param = L->param.data();
        p.n_free = (uint32_t)L->free_local.size(); p.free_vars = L->free_local.data();
        p.n_rows = (uint32_t)L->rows.size(); p.rows = L->rows.data();
        locals.push_back(std::move(L));
    }
    if (locals.empty()) return FK_OK;
    std::vector<const fk_problem*> probs;
    std::vector<double*> xs;
    for (auto& L : locals) {
        probs.push_back(&L->prob);
        xs.push_back(L->x.data());
    }
    std::vector<fk_report> reps(locals.size());
    int rc = fk_lm_solve_batch((uint32_t)locals.size(), probs.data(), xs.data(), reps.data(), 1);
    if (rc != FK_OK) return rc;
    // assemble/mod.rs:161-166
    for (auto& L : locals)
        for (size_t k = 0; k < L->free_global.size(); k++) s->vars[L->free_global[k]] = scale * L->x[k];
    for (size_t k = 0; k < locals.size() && reports && k < cap; k++) reports[k] = reps[k];
    if (n_solved) *n_solved = (uint32_t)locals.size();
    return FK_OK;
}

// ---- Decomposer::SinglePass ------------------------------------------------------------
namespace {

// The equation graph of lib.rs:262,395,434.
void equation_graph(const fk_system* s, std::vector<std::vector<uint32_t>>& var_exprs, std::vector<std::vector<uint32_t>>& expr_vars) {
    var_exprs.assign(s->vars.size(), {});
    expr_vars.assign(s->kind.size(), {});
    for (uint32_t e = 0; e < s->kind.size(); e++) {
        uint32_t sv[8];
        const int a = fk::expand_slots(s->kind[e], &s->idx[4 * (size_t)e], sv);
        expr_vars[e].assign(sv, sv + a);
        for (int k = 0; k < a; k++) var_exprs[sv[k]].push_back(e);
    }
}

std::vector<uint32_t> component_free_variables(const fk_system* s, const fk_system::Comp& c) {
    std::set<uint32_t> free_set;
    for (uint32_t e : c.elements) {
        uint32_t v[4];
        const int n = s->element_vars(e, v);
        for (int k = 0; k < n; k++)
            if (!s->fixed[v[k]]) free_set.insert(v[k]);
    }
    return std::vector<uint32_t>(free_set.begin(), free_set.end());
}

// One compact fk_problem: `free_sorted` free, `exprs` as rows (in that order), every other referenced
// variable fixed at its current value in `vt`.
struct LocalProblem {
    std::vector<double> vars, param, x;
    std::vector<uint8_t> kind;
    std::vector<uint32_t> idx, free_local, rows, free_global;
    fk_problem prob;
};
void build_local(const fk_system* s, const std::vector<double>& vt, const std::vector<double>& pt, const std::vector<uint32_t>& free_sorted,
                 const std::vector<uint32_t>& exprs, LocalProblem& L) {
    std::set<uint32_t> used(free_sorted.begin(), free_sorted.end());
    for (uint32_t e : exprs) {
        uint32_t sv[8];
        const int a = fk::expand_slots(s->kind[e], &s->idx[4 * (size_t)e], sv);
        for (int k = 0; k < a; k++) used.insert(sv[k]);
    }
    std::map<uint32_t, uint32_t> local;
    for (uint32_t g : used) {
        local.emplace(g, (uint32_t)L.vars.size());
        L.vars.push_back(vt[g]);
    }
    for (uint32_t g : free_sorted) {
        L.free_global.push_back(g);
        L.free_local.push_back(local[g]);
        L.x.push_back(vt[g]);
    }
    static const int stored[FK_NUM_KINDS] = {2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 3, 3};
    for (uint32_t e : exprs) {
        L.rows.push_back((uint32_t)L.kind.size());
        L.kind.push_back(s->kind[e]);
        L.param.push_back(pt[e]);
        for (int q = 0; q < 4; q++) L.idx.push_back(q < stored[s->kind[e]] ? local[s->idx[4 * (size_t)e + q]] : 0);
    }
    fk_problem& p = L.prob;
    p.n_vars = (uint32_t)L.vars.size(); p.vars = L.vars.data();
    p.n_expr = (uint32_t)L.kind.size(); p.kind = L.kind.data(); p.idx = L.idx.data(); p.param = L.param.data();
    p.n_free = (uint32_t)L.free_local.size(); p.free_vars = L.free_local.data();
    p.n_rows = (uint32_t)L.rows.size(); p.rows = L.rows.data();
}

}  // namespace

// The sequence of sub-problems SinglePass solves, all components in order (host only, no device):
// call with NULL arrays for sizes3 = {steps, total free variables, total expressions}, then with
// free_ptr[steps+1], free_vars, expr_ptr[steps+1], exprs.
int fk_system_single_pass_plan(const fk_system* s, uint32_t* sizes3, uint32_t* free_ptr, uint32_t* free_vars, uint32_t* expr_ptr,
                               uint32_t* exprs) {
    if (!s) return FK_ERR_INVALID;
    std::vector<std::vector<uint32_t>> var_exprs, expr_vars;
    equation_graph(s, var_exprs, expr_vars);
    fk::SinglePassPlanner planner(var_exprs, expr_vars);
    std::vector<fk::SinglePassStep> all;
    for (const fk_system::Comp& c : s->comps) {
        if (c.elements.empty()) continue;
        for (fk::SinglePassStep& st : planner.plan(component_free_variables(s, c))) all.push_back(std::move(st));
    }
    uint32_t nf = 0, ne = 0;
    for (const auto& st : all) { nf += (uint32_t)st.free_variables.size(); ne += (uint32_t)st.expressions.size(); }
    if (sizes3) { sizes3[0] = (uint32_t)all.size(); sizes3[1] = nf; sizes3[2] = ne; }
    if (!free_ptr || !free_vars || !expr_ptr || !exprs) return FK_OK;
    uint32_t af = 0, ae = 0;
    for (size_t k = 0; k < all.size(); k++) {
        free_ptr[k] = af; expr_ptr[k] = ae;
        for (uint32_t v : all[k].free_variables) free_vars[af++] = v;
        for (uint32_t e : all[k].expressions) exprs[ae++] = e;
    }
    free_ptr[all.size()] = af; expr_ptr[all.size()] = ae;
    return FK_OK;
}


// ---- Decomposer::RecursiveAssembly -------------------------------------------------------
namespace {

// graph.rs:98-117 as the planner wants it: dof = variables an element adds (lib.rs:372-403), valency = expressions
// of the constraint (constraints/mod.rs:331-334 and the like).
fk::RaGraph constraint_graph(const fk_system* s) {
    fk::RaGraph g;
    for (const Elem& e : s->elems) g.add_vertex(e.tag == kLength ? 1 : (e.tag == kPoint ? 2 : 0));
    for (const Constr& c : s->constrs) g.add_edge(c.n_expr, c.inc, c.n_inc);
    return g;
}

std::vector<fk::RaStep> component_plan(const fk_system* s, const fk_system::Comp& c) {
    fk::RecursiveAssemblyPlanner planner(constraint_graph(s));
    return planner.plan(std::vector<uint32_t>(c.elements.begin(), c.elements.end()),
                        std::vector<uint32_t>(c.constraints.begin(), c.constraints.end()));
}

const std::vector<uint32_t>* find_list(const fk::RaLists& m, uint32_t key) {
    auto it = m.find(key);
    return it == m.end() ? nullptr : &it->second;
}

// One step of the plan as a flattened problem (ClusteredSystem, assemble/mod.rs:282-590): free variables = the poses
// of the clusters the step moves (three each, starting at 0) followed by the variables of the step's elements; rows =
// two FK_POSE_POINT rows per (cluster, frontier point), then the step's expressions with their variables renumbered.
// The points' positions before the step enter as fixed variables behind the free ones.
struct ClusteredProblem {
    std::vector<uint32_t> elements;                                        // step_plus_frontier_elements
    std::vector<std::pair<uint32_t, std::vector<uint32_t>>> clusters;      // cluster key -> frontier points, in first-seen order
    std::map<uint32_t, uint32_t> slot;                                     // system variable -> free variable of the step
    std::vector<double> vars, param, x;
    std::vector<uint8_t> kind;
    std::vector<uint32_t> idx, free_vars, rows;
    fk_problem prob{};
    std::string error;

    bool build(const fk_system* s, const fk::RaStep& step, const std::vector<double>& vt, const std::vector<double>& pt) {
        auto has = [](const std::vector<uint32_t>& v, uint32_t x) { return std::find(v.begin(), v.end(), x) != v.end(); };
        elements = step.elements;
        // clusters reachable through shared frontier points (:355-398)
        std::vector<uint32_t> reach;
        auto add_clusters_of = [&](uint32_t el) -> size_t {
            const std::vector<uint32_t>* l = find_list(step.on_frontiers, el);
            if (!l) return 0;
            for (uint32_t c : *l)
                if (!has(reach, c)) reach.push_back(c);
            return l->size();
        };
        for (uint32_t el : step.elements)
            if (s->elems[el].tag == kPoint) add_clusters_of(el);
        for (size_t i = 0; i < reach.size(); i++) {
            const std::vector<uint32_t>* fr = find_list(step.frontier_elements, reach[i]);
            if (!fr) { error = "recursive assembly: cluster without a frontier list"; return false; }
            for (uint32_t el : *fr) {
                if (s->elems[el].tag != kPoint) continue;
                const size_t n_frontiers = add_clusters_of(el);
                if (!has(elements, el) && n_frontiers > 1) elements.push_back(el);
            }
        }
        // pose rows: one pair per (cluster, point on its frontier) (:401-429)
        for (uint32_t el : elements) {
            const std::vector<uint32_t>* l = find_list(step.on_frontiers, el);
            if (!l || s->elems[el].tag != kPoint) continue;
            for (uint32_t c : *l) {
                auto it = std::find_if(clusters.begin(), clusters.end(), [&](const std::pair<uint32_t, std::vector<uint32_t>>& p) { return p.first == c; });
                if (it == clusters.end()) {
                    clusters.emplace_back(c, std::vector<uint32_t>());
                    it = clusters.end() - 1;
                }
                it->second.push_back(el);
            }
        }
        // free variables (:432-477)
        vars.assign(clusters.size() * 3, 0.0);
        for (uint32_t el : elements) {
            const Elem& e = s->elems[el];
            const int n = e.tag == kLength ? 1 : (e.tag == kPoint ? 2 : 0);
            for (int k = 0; k < n; k++) {
                slot[e.a + k] = (uint32_t)vars.size();
                vars.push_back(vt[e.a + k]);
            }
        }
        const uint32_t n_free = (uint32_t)vars.size();
        x = vars;
        for (uint32_t k = 0; k < n_free; k++) free_vars.push_back(k);
        // rows
        for (size_t ci = 0; ci < clusters.size(); ci++)
            for (uint32_t point : clusters[ci].second) {
                const uint32_t pv = s->elems[point].a;
                const uint32_t at = (uint32_t)vars.size();
                vars.push_back(vt[pv]);      // the point before the step: constant of the two rows
                vars.push_back(vt[pv + 1]);
                for (int y = 0; y < 2; y++) {
                    rows.push_back((uint32_t)kind.size());
                    kind.push_back(y ? FK_POSE_POINT_Y : FK_POSE_POINT_X);
                    param.push_back(0.0);
                    idx.push_back((uint32_t)(3 * ci)); idx.push_back(slot.at(pv) + (uint32_t)y); idx.push_back(at); idx.push_back(0);
                }
            }
        static const int stored[FK_NUM_KINDS] = {2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 3, 3};
        for (uint32_t c : step.constraints)
            for (uint8_t k = 0; k < s->constrs[c].n_expr; k++) {
                const uint32_t e = s->constrs[c].first_expr + k;
                rows.push_back((uint32_t)kind.size());
                kind.push_back(s->kind[e]);
                param.push_back(pt[e]);
                for (int q = 0; q < 4; q++) {
                    uint32_t v = 0;
                    if (q < stored[s->kind[e]]) {
                        auto it = slot.find(s->idx[4 * (size_t)e + q]);
                        if (it == slot.end()) { error = "recursive assembly: an expression reads a variable outside its step"; return false; }
                        v = it->second;
                    }
                    idx.push_back(v);
                }
            }
        prob.n_vars = (uint32_t)vars.size(); prob.vars = vars.data();
        prob.n_expr = (uint32_t)kind.size(); prob.kind = kind.data(); prob.idx = idx.data(); prob.param = param.data();
        prob.n_free = n_free; prob.free_vars = free_vars.data();
        prob.n_rows = (uint32_t)rows.size(); prob.rows = rows.data();
        return true;
    }
};

// == System::solve(SolvingOptions { optimizer: LevenbergMarquardt, decomposer: RecursiveAssembly, perturb })
// (assemble/mod.rs:212-277): every step of every component's plan is one LM solve on the GPU; afterwards the points a
// moved cluster owns (and the step did not solve for) follow the cluster's pose on the host.  As in the reference,
// `fix` has no effect on this decomposer (ClusteredSystem takes every variable of the step's elements).
int solve_recursive_assembly(fk_system* s, int perturb, fk_report* reports, uint32_t cap, uint32_t* n_solved) {
    if (n_solved) *n_solved = 0;
    const size_t nv = s->vars.size();
    double sum = 0.0;
    size_t cnt = 0;
    for (double v : s->vars) { sum += v * v; cnt++; }
    for (size_t e = 0; e < s->kind.size(); e++)
        if (s->kind[e] == FK_POINT_POINT_DISTANCE || s->kind[e] == FK_POINT_LINE_DISTANCE) { sum += s->param[e] * s->param[e]; cnt++; }
    const double scale = std::sqrt(sum / (double)cnt);
    const double recip = 1.0 / scale;
    std::vector<double> vt(nv);
    for (size_t i = 0; i < nv; i++) vt[i] = s->vars[i] * recip;
    std::vector<double> pt(s->param);
    for (size_t e = 0; e < pt.size(); e++)
        if (s->kind[e] == FK_POINT_POINT_DISTANCE || s->kind[e] == FK_POINT_LINE_DISTANCE) pt[e] = recip * s->param[e];
    Lcg rng{42};
    uint32_t solved = 0;
    for (const fk_system::Comp& c : s->comps) {
        if (c.elements.empty()) continue;
        if (perturb) {
            for (uint32_t fv : component_free_variables(s, c)) {
                const double r1 = rng.next();
                const double r2 = rng.next();
                vt[fv] += vt[fv] * (1.0 / 8196.0) * r1 + (1.0 / 65568.0) * r2;
            }
        }
        std::vector<fk::RaStep> steps;
        try {
            steps = component_plan(s, c);
        } catch (const std::out_of_range&) {  // the reference panics here: unwrap on a missing cluster entry (recursive_assembly.rs:352-373)
            s->error = "recursive assembly: the decomposition reaches a cluster without bookkeeping (the reference panics on this system)";
            return FK_ERR_INVALID;
        }
        for (const fk::RaStep& step : steps) {
            ClusteredProblem P;
            if (!P.build(s, step, vt, pt)) {
                s->error = P.error;
                return FK_ERR_INVALID;
            }
            fk_report rep{};
            const int rc = fk_lm_solve(&P.prob, P.x.data(), &rep);
            if (rc != FK_OK) return rc;
            for (const auto& kv : P.slot) {  // :226-233
                vt[kv.first] = P.x[kv.second];
                s->vars[kv.first] = scale * P.x[kv.second];
            }
            for (size_t ci = 0; ci < P.clusters.size(); ci++) {  // :235-273
                const std::vector<uint32_t>* owned = find_list(step.owned_elements, P.clusters[ci].first);
                if (!owned) continue;
                const double rot = P.x[3 * ci], tx = P.x[3 * ci + 1], ty = P.x[3 * ci + 2];
                const double sn = std::sin(rot), cs = std::cos(rot);  // Pose2D::transform_point, expressions.rs:1122-1136
                for (uint32_t el : *owned) {
                    if (std::find(P.elements.begin(), P.elements.end(), el) != P.elements.end() || s->elems[el].tag != kPoint) continue;
                    const uint32_t pv = s->elems[el].a;
                    const double u = vt[pv], w = vt[pv + 1];
                    const double uc = u * cs, us = u * sn, vc = w * cs, vs = w * sn;
                    const double nx = tx + uc - vs, ny = ty + us + vc;
                    vt[pv] = nx; vt[pv + 1] = ny;
                    s->vars[pv] = scale * nx; s->vars[pv + 1] = scale * ny;
                }
            }
            if (reports && solved < cap) reports[solved] = rep;
            solved++;
        }
    }
    if (n_solved) *n_solved = solved;
    return FK_OK;
}

}  // namespace

// == System::solve(SolvingOptions { optimizer: LevenbergMarquardt, decomposer, perturb }).
// decomposer 0: None (== fk_system_solve); 1: SinglePass (assemble/mod.rs:169-210): the strongly
// connected expression sets of every component are solved one after the other on the GPU, each seeing
// the variables solved before it as fixed values.  reports: one per solved sub-problem.
int fk_system_solve_opts(fk_system* s, int decomposer, int perturb, fk_report* reports, uint32_t cap, uint32_t* n_solved) {
    if (!s) return FK_ERR_INVALID;
    if (decomposer == 0) return fk_system_solve(s, perturb, reports, cap, n_solved);
    if (decomposer == 2) return solve_recursive_assembly(s, perturb, reports, cap, n_solved);
    if (decomposer != 1) return FK_ERR_INVALID;
    if (n_solved) *n_solved = 0;
    const size_t nv = s->vars.size();
    double sum = 0.0;
    size_t cnt = 0;
    for (double v : s->vars) { sum += v * v; cnt++; }
    for (size_t e = 0; e < s->kind.size(); e++)
        if (s->kind[e] == FK_POINT_POINT_DISTANCE || s->kind[e] == FK_POINT_LINE_DISTANCE) { sum += s->param[e] * s->param[e]; cnt++; }
    const double scale = std::sqrt(sum / (double)cnt);
    const double recip = 1.0 / scale;
    std::vector<double> vt(nv);
    for (size_t i = 0; i < nv; i++) vt[i] = s->vars[i] * recip;
    std::vector<double> pt(s->param);
    for (size_t e = 0; e < pt.size(); e++)
        if (s->kind[e] == FK_POINT_POINT_DISTANCE || s->kind[e] == FK_POINT_LINE_DISTANCE) pt[e] = recip * s->param[e];
    std::vector<std::vector<uint32_t>> var_exprs, expr_vars;
    equation_graph(s, var_exprs, expr_vars);
    fk::SinglePassPlanner planner(var_exprs, expr_vars);
    Lcg rng{42};
    uint32_t solved = 0;
    for (const fk_system::Comp& c : s->comps) {
        if (c.elements.empty()) continue;
        const std::vector<uint32_t> free_sorted = component_free_variables(s, c);
        if (perturb) {
            for (uint32_t fv : free_sorted) {
                const double r1 = rng.next();
                const double r2 = rng.next();
                vt[fv] += vt[fv] * (1.0 / 8196.0) * r1 + (1.0 / 65568.0) * r2;
            }
        }
        for (const fk::SinglePassStep& st : planner.plan(free_sorted)) {
            LocalProblem L;
            build_local(s, vt, pt, st.free_variables, st.expressions, L);
            fk_report rep{};
            const int rc = fk_lm_solve(&L.prob, L.x.data(), &rep);
            if (rc != FK_OK) return rc;
            for (size_t k = 0; k < L.free_global.size(); k++) {  // assemble/mod.rs:201-208
                vt[L.free_global[k]] = L.x[k];
                s->vars[L.free_global[k]] = scale * L.x[k];
            }
            if (reports && solved < cap) reports[solved] = rep;
            solved++;
        }
    }
    if (n_solved) *n_solved = solved;
    return FK_OK;
}

// The recombination plan of all components as one stream of 32-bit words (host only, no device): per step
// n_constraints, constraints..., n_elements, elements..., n_free, free elements..., then on_frontiers, owned_elements,
// frontier_elements, each as n_keys and per key (ascending)  key, n, values....  Returns FK_OK; *n_words receives the
// stream's length (at most `cap` words are written), *n_steps the number of steps.
int fk_system_recursive_assembly_plan(const fk_system* s, uint32_t* out, uint32_t cap, uint32_t* n_words, uint32_t* n_steps) {
    if (!s) return FK_ERR_INVALID;
    std::vector<uint32_t> w;
    auto list = [&](const std::vector<uint32_t>& v) {
        w.push_back((uint32_t)v.size());
        w.insert(w.end(), v.begin(), v.end());
    };
    auto lists = [&](const fk::RaLists& m) {
        w.push_back((uint32_t)m.size());
        for (const auto& kv : m) {
            w.push_back(kv.first);
            list(kv.second);
        }
    };
    uint32_t steps = 0;
    for (const fk_system::Comp& c : s->comps) {
        if (c.elements.empty()) continue;
        std::vector<fk::RaStep> plan;
        try {
            plan = component_plan(s, c);
        } catch (const std::out_of_range&) {  // (the reference panics on this system, see solve_recursive_assembly)
            return FK_ERR_INVALID;
        }
        for (const fk::RaStep& st : plan) {
            list(st.constraints); list(st.elements); list(st.free_elements);
            lists(st.on_frontiers); lists(st.owned_elements); lists(st.frontier_elements);
            steps++;
        }
    }
    if (n_words) *n_words = (uint32_t)w.size();
    if (n_steps) *n_steps = steps;
    for (size_t k = 0; k < w.size() && k < cap && out; k++) out[k] = w[k];
    return FK_OK;
}

// == System::analyze (lib.rs:454-458, analyze/numerical/mod.rs:123-163): the dense Jacobian + the
// incremental Gauss-Jordan run on the GPU (fk_batch_analyze with a batch of one); the host maps
// dependent expressions back to their constraints (:149-160).
int fk_system_analyze(fk_system* s, uint32_t* constraints, uint32_t cap, uint32_t* n_found) {
    if (!s) return FK_ERR_INVALID;
    if (n_found) *n_found = 0;
    const uint32_t ne = (uint32_t)s->kind.size();
    if (ne == 0) return FK_OK;
    std::vector<uint32_t> all_vars(s->vars.size()), all_rows(ne);
    for (uint32_t i = 0; i < all_vars.size(); i++) all_vars[i] = i;
    for (uint32_t i = 0; i < ne; i++) all_rows[i] = i;
    fk_problem p{};
    p.n_vars = (uint32_t)s->vars.size(); p.vars = s->vars.data();
    p.n_expr = ne; p.kind = s->kind.data(); p.idx = s->idx.data(); p.param = s->param.data();
    p.n_free = p.n_vars; p.free_vars = all_vars.data();
    p.n_rows = ne; p.rows = all_rows.data();
    fk_topology* t = nullptr;
    int rc = fk_topology_create(&p, &t);
    if (rc != FK_OK) return rc;
    std::vector<uint8_t> independent(ne, 0);
    rc = fk_batch_analyze(t, 0, 1, s->vars.data(), s->param.data(), independent.data());
    fk_topology_destroy(t);
    if (rc != FK_OK) return rc;
    std::vector<uint32_t> expr_to_constraint(ne, 0);
    for (uint32_t c = 0; c < s->constrs.size(); c++)
        for (uint8_t k = 0; k < s->constrs[c].n_expr; k++) expr_to_constraint[s->constrs[c].first_expr + k] = c;
    uint32_t found = 0;
    for (uint32_t e = 0; e < ne; e++)
        if (!independent[e]) {
            if (constraints && found < cap) constraints[found] = expr_to_constraint[e];
            found++;
        }
    if (n_found) *n_found = found;
    return FK_OK;
}

// Residual of every constraint at the current (unscaled) variables == ConstraintHandle::
// calculate_residual (constraints/mod.rs:88-110): the expression residual, or the 2-norm of the two
// equalities of a coincidence.  Evaluated by the K2 kernel.
int fk_system_residuals(fk_system* s, double* out) {
    if (!s || !out) return FK_ERR_INVALID;
    const uint32_t ne = (uint32_t)s->kind.size();
    if (ne == 0) return FK_OK;
    std::vector<double> r(ne);
    std::vector<uint32_t> all_vars(s->vars.size()), all_rows(ne);
    for (uint32_t i = 0; i < all_vars.size(); i++) all_vars[i] = i;
    for (uint32_t i = 0; i < ne; i++) all_rows[i] = i;
    if (!s->eval_topo) {
        fk_problem p{};
        p.n_vars = (uint32_t)s->vars.size(); p.vars = s->vars.data();
        p.n_expr = ne; p.kind = s->kind.data(); p.idx = s->idx.data(); p.param = s->param.data();
        p.n_free = p.n_vars; p.free_vars = all_vars.data();
        p.n_rows = ne; p.rows = all_rows.data();
        fk_topology* t = nullptr;
        int rc = fk_topology_create(&p, &t);
        if (rc != FK_OK) return rc;
        s->eval_topo.reset(t);
    }
    fk_topology_info info;
    fk_topology_info_get(s->eval_topo.get(), &info);
    int rc;
    if (info.path == 2) {
        rc = fk_topology_eval(s->eval_topo.get(), s->vars.data(), s->param.data(), s->vars.data(), r.data(), nullptr, 0, nullptr);
    } else {
        fk_batch_plan* plan = nullptr;
        int dev = 0;
        rc = fk_batch_plan_create(s->eval_topo.get(), 1, dev, &plan);
        if (rc == FK_OK) rc = fk_batch_plan_upload(plan, 1, s->vars.data(), s->param.data(), nullptr);
        if (rc == FK_OK) rc = fk_batch_plan_eval(plan, 1, nullptr);
        if (rc == FK_OK) rc = fk_batch_plan_eval_download(plan, r.data(), nullptr, nullptr);
        if (rc == FK_OK) rc = fk_batch_plan_sync(plan);
        fk_batch_plan_destroy(plan);
    }
    if (rc != FK_OK) return rc;
    for (size_t c = 0; c < s->constrs.size(); c++) {
        const Constr& k = s->constrs[c];
        if (k.n_expr > 1) {
            double q = 0.0;
            for (uint8_t e = 0; e < k.n_expr; e++) q += r[k.first_expr + e] * r[k.first_expr + e];
            out[c] = std::sqrt(q);
        } else {
            out[c] = r[k.first_expr];
        }
    }
    return FK_OK;
}

// Connected components as the reference's graph holds them (including element-less slots).
uint32_t fk_system_num_components(const fk_system* s) { return s ? (uint32_t)s->comps.size() : 0; }
int fk_system_component(const fk_system* s, uint32_t ci, uint32_t* n_elements, uint32_t* elements, uint32_t* n_constraints, uint32_t* constraints) {
    if (!s || ci >= s->comps.size()) return FK_ERR_INVALID;
    const fk_system::Comp& c = s->comps[ci];
    if (n_elements) *n_elements = (uint32_t)c.elements.size();
    if (n_constraints) *n_constraints = (uint32_t)c.constraints.size();
    if (elements) std::copy(c.elements.begin(), c.elements.end(), elements);
    if (constraints) std::copy(c.constraints.begin(), c.constraints.end(), constraints);
    return FK_OK;
}

}  // extern "C"
