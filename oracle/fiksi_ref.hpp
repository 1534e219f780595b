// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of Fiksi's numeric solve path: the 11 residual/gradient expressions, the
// `Subsystem` problem (free/fixed variable map, COO triplets for free columns), the
// Levenberg–Marquardt driver on the augmented system [J; sqrt(lambda) I] with sparse Householder
// QR under COLAMD ordering, the connected-component graph (including its stale-index behaviour),
// and `assemble::solve`'s scale / perturb / write-back.  Every routine cites the reference lines
// it follows (paths relative to /root/reference).
//
// Parity status: the linear algebra underneath is pinned by the reference's known-answer tests
// (see solvi_ref.hpp / colamd_ref.hpp); the LCG by fiksi/src/rand.rs:49-63; the expression
// gradients by the reference's finite-difference property (expressions.rs:1196-1510); the
// end-to-end scenarios by the reference's residual thresholds (fiksi/src/tests/*.rs).  The
// reference cannot be compiled in this environment (no Rust toolchain), so CONVERGED COORDINATES
// ARE PARITY-UNPINNED against a reference binary; SURVEY.md App. E hand-derived checkpoints are
// checked in tests/test_oracle_fiksi.py.
//
// One extension: the reference's inner damping loop is unbounded and spins forever once lambda
// overflows to +inf (every later step is NaN; lm.rs:115-191).  The oracle stops there and reports
// exit_reason 4 ("reference would hang").
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "solvi_ref.hpp"

namespace orc {
namespace fiksi {

// fiksi/src/rand.rs:13-40
struct Rng {
    uint32_t state;
    explicit Rng(uint32_t seed) : state(seed) {}
    uint32_t next_u32() {
        state = state * 1664525u + 1013904223u;
        return state;
    }
    double next_f64() {
        uint32_t val = next_u32();
        return (1. / (double)UINT32_MAX) * (double)val;
    }
};

// fiksi/src/utils.rs:12-33 (sequential left-to-right sums)
inline double sum_squares(const double* v, size_t n) {
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += v[i] * v[i];
    return s;
}

// fiksi/src/constraints/expressions.rs:28-40 — kinds in enum order.
enum Kind : uint8_t {
    VariableVariableEquality = 0,
    PointPointDistance = 1,
    PointPointPointAngle = 2,
    PointLineIncidence = 3,
    PointLineDistance = 4,
    PointCircleIncidence = 5,
    SegmentSegmentLengthEquality = 6,
    LineLineAngle = 7,
    LineLineParallelism = 8,
    LineLinePerpendicularity = 9,
    LineCircleTangency = 10,
    NUM_KINDS = 11,
};

struct Expression {
    uint8_t kind = 0;
    uint32_t idx[4] = {0, 0, 0, 0};  // base indices exactly as stored in the reference structs
    double param = 0.0;              // distance / angle, 0 if none
};

// expressions.rs:48-182.  Returns the number of variable slots.
inline int variable_indices(const Expression& e, uint32_t out[8]) {
    const uint32_t* i = e.idx;
    switch (e.kind) {
        case VariableVariableEquality:
            out[0] = i[0]; out[1] = i[1];
            return 2;
        case PointPointDistance:
            out[0] = i[0]; out[1] = i[0] + 1; out[2] = i[1]; out[3] = i[1] + 1;
            return 4;
        case PointPointPointAngle:
        case PointLineIncidence:
        case PointLineDistance:
            out[0] = i[0]; out[1] = i[0] + 1; out[2] = i[1]; out[3] = i[1] + 1;
            out[4] = i[2]; out[5] = i[2] + 1;
            return 6;
        case PointCircleIncidence:
            out[0] = i[0]; out[1] = i[0] + 1; out[2] = i[1]; out[3] = i[1] + 1; out[4] = i[2];
            return 5;
        case SegmentSegmentLengthEquality:
        case LineLineAngle:
        case LineLineParallelism:
        case LineLinePerpendicularity:
            out[0] = i[0]; out[1] = i[0] + 1; out[2] = i[1]; out[3] = i[1] + 1;
            out[4] = i[2]; out[5] = i[2] + 1; out[6] = i[3]; out[7] = i[3] + 1;
            return 8;
        case LineCircleTangency:
            out[0] = i[0]; out[1] = i[0] + 1; out[2] = i[1]; out[3] = i[1] + 1;
            out[4] = i[2]; out[5] = i[2] + 1; out[6] = i[3];
            return 7;
    }
    return 0;
}

// expressions.rs:195-211
inline Expression transform(const Expression& e, double length_scale_recip) {
    Expression t = e;
    if (e.kind == PointPointDistance || e.kind == PointLineDistance)
        t.param = length_scale_recip * e.param;
    return t;
}

static const double PI = 3.14159265358979323846264338327950288;  // core::f64::consts::PI

// expressions.rs:319-353
inline double ppd_eval(const double* v, double param_distance, double g[4]) {
    double p1x = v[0], p1y = v[1], p2x = v[2], p2y = v[3];
    double dx = p1x - p2x, dy = p1y - p2y;
    double distance = std::sqrt(dx * dx + dy * dy);
    double residual = distance - param_distance;
    double distance_recip = 1. / distance;
    g[0] = (p1x - p2x) * distance_recip;
    g[1] = (p1y - p2y) * distance_recip;
    g[2] = -(p1x - p2x) * distance_recip;
    g[3] = -(p1y - p2y) * distance_recip;
    return residual;
}

inline double wrap_angle(double angle) {
    // expressions.rs:393-399,665-671
    if (angle > PI) return angle - 2.0 * PI;
    if (angle < -PI) return angle + 2.0 * PI;
    return angle;
}

// expressions.rs:214-276 + the per-kind kernels :291-874.  `v` holds the slot values in
// `variable_indices` order; `g` receives one entry per slot.  Returns the residual.
inline double compute_residual_and_gradient(const Expression& e, const double v[8], double g[8]) {
    switch (e.kind) {
        case VariableVariableEquality: {  // :291-301
            g[0] = -1.;
            g[1] = 1.;
            return v[1] - v[0];
        }
        case PointPointDistance:
            return ppd_eval(v, e.param, g);
        case PointPointPointAngle: {  // :372-425
            double ux = v[0] - v[2], uy = v[1] - v[3];
            double vx = v[4] - v[2], vy = v[5] - v[3];
            double angle = wrap_angle(std::atan2(vy, vx) - std::atan2(uy, ux));
            double residual = angle - e.param;
            double ur = 1. / (ux * ux + uy * uy);
            double vr = 1. / (vx * vx + vy * vy);
            double d1x = uy * ur, d1y = -ux * ur;
            double d3x = -vy * vr, d3y = vx * vr;
            double d2x = -d1x - d3x, d2y = -d1y - d3y;
            g[0] = d1x; g[1] = d1y; g[2] = d2x; g[3] = d2y; g[4] = d3x; g[5] = d3y;
            return residual;
        }
        case PointLineIncidence: {  // :445-477
            double px = v[0], py = v[1], l1x = v[2], l1y = v[3], l2x = v[4], l2y = v[5];
            double ux = l2x - l1x, uy = l2y - l1y;
            double wx = px - l1x, wy = py - l1y;
            double residual = ux * wy - uy * wx;
            g[0] = -uy; g[1] = ux; g[2] = -py + l2y; g[3] = px - l2x; g[4] = wy; g[5] = -wx;
            return residual;
        }
        case PointLineDistance: {  // :500-544
            double px = v[0], py = v[1], l1x = v[2], l1y = v[3], l2x = v[4], l2y = v[5];
            double ux = l2x - l1x, uy = l2y - l1y;
            double wx = px - l1x, wy = py - l1y;
            double cross = ux * wy - uy * wx;
            double len2 = ux * ux + uy * uy;
            double len = std::sqrt(len2);
            double lr = 1. / len;
            double a = cross / len2;
            double b = -a * ux;
            double c = px + a * uy;
            double residual = lr * cross - e.param;
            g[0] = -lr * uy;
            g[1] = lr * ux;
            g[2] = -lr * (b - l2y + py);
            g[3] = -lr * (l2x - c);
            g[4] = lr * (b + wy);
            g[5] = -lr * (c - l1x);
            return residual;
        }
        case PointCircleIncidence: {  // :560-576
            double r = ppd_eval(v, v[4], g);
            g[4] = -1.;
            return r;
        }
        case SegmentSegmentLengthEquality: {  // :593-620
            double g1[4], g2[4];
            double r1 = ppd_eval(v, 0., g1);
            double r2 = ppd_eval(v + 4, 0., g2);
            g[0] = -g1[0]; g[1] = -g1[1]; g[2] = -g1[2]; g[3] = -g1[3];
            g[4] = g2[0]; g[5] = g2[1]; g[6] = g2[2]; g[7] = g2[3];
            return r2 - r1;
        }
        case LineLineAngle: {  // :640-696
            double ux = v[2] - v[0], uy = v[3] - v[1];
            double vx = v[6] - v[4], vy = v[7] - v[5];
            double angle = wrap_angle(std::atan2(vy, vx) - std::atan2(uy, ux));
            double residual = angle - e.param;
            double ur = 1. / (ux * ux + uy * uy);
            double vr = 1. / (vx * vx + vy * vy);
            double a = -uy * ur, b = ux * ur, c = vy * vr, d = -vx * vr;
            g[0] = a; g[1] = b; g[2] = -a; g[3] = -b; g[4] = c; g[5] = d; g[6] = -c; g[7] = -d;
            return residual;
        }
        case LineLineParallelism: {  // :713-752
            double ux = v[2] - v[0], uy = v[3] - v[1];
            double vx = v[6] - v[4], vy = v[7] - v[5];
            double residual = vx * uy - vy * ux;
            g[0] = vy; g[1] = -vx; g[2] = -vy; g[3] = vx; g[4] = -uy; g[5] = ux; g[6] = uy; g[7] = -ux;
            return residual;
        }
        case LineLinePerpendicularity: {  // :769-799
            double ux = v[2] - v[0], uy = v[3] - v[1];
            double vx = v[6] - v[4], vy = v[7] - v[5];
            double residual = vx * ux + vy * uy;
            g[0] = -vx; g[1] = -vy; g[2] = vx; g[3] = vy; g[4] = -ux; g[5] = -uy; g[6] = ux; g[7] = uy;
            return residual;
        }
        case LineCircleTangency: {  // :816-874
            double l1x = v[0], l1y = v[1], l2x = v[2], l2y = v[3], cx = v[4], cy = v[5], rad = v[6];
            double ddx = l1x - l2x, ddy = l1y - l2y;
            double length2 = ddx * ddx + ddy * ddy;
            double length = std::sqrt(length2);
            if (length == 0.) {
                for (int k = 0; k < 7; k++) g[k] = 0.;
                return 0.;
            }
            double length_recip = 1. / length;
            double signed_area = l1x * (l2y - cy) + l2x * (cy - l1y) + cx * (l1y - l2y);
            double residual = length_recip * std::fabs(signed_area) - rad;
            // f64::signum: 1.0 for +0.0 and positives, -1.0 for -0.0 and negatives, NaN for NaN
            double sign = std::isnan(signed_area) ? signed_area : std::copysign(1.0, signed_area);
            double length3_recip = 1. / (length2 * length);
            g[0] = sign * length3_recip * (length2 * (l2y - cy) + signed_area * (l2x - l1x));
            g[1] = sign * length3_recip * (length2 * (-l2x + cx) + signed_area * (l2y - l1y));
            g[2] = sign * length3_recip * (length2 * (cy - l1y) - signed_area * (l2x - l1x));
            g[3] = sign * length3_recip * (length2 * (l1x - cx) - signed_area * (l2y - l1y));
            g[4] = sign * length_recip * (l1y - l2y);
            g[5] = sign * length_recip * (-l1x + l2x);
            g[6] = -1.;
            return residual;
        }
    }
    return 0.;
}

// fiksi/src/subsystem.rs:9-167 + variable_map.rs:43-73.  The IndexSet lookup becomes a dense
// var -> free index table (same mapping: index == insertion order of `free_variables`).
struct Subsystem {
    const double* system_variables;      // scaled system variables (fixed values are read here)
    const Expression* all_expressions;   // scaled expressions
    std::vector<uint32_t> expressions;   // rows: expression ids
    std::vector<uint32_t> free_variables;  // insertion order == free index
    std::vector<int32_t> var_to_free;    // -1 if fixed / not in this subsystem

    Subsystem(const double* vars, size_t n_vars, const Expression* exprs,
              const std::vector<uint32_t>& free_vars, const std::vector<uint32_t>& rows)
        : system_variables(vars), all_expressions(exprs), expressions(rows), free_variables(free_vars),
          var_to_free(n_vars, -1) {
        for (size_t k = 0; k < free_vars.size(); k++)
            if (var_to_free[free_vars[k]] < 0) var_to_free[free_vars[k]] = (int32_t)k;
    }
    uint32_t num_variables() const { return (uint32_t)free_variables.size(); }
    uint32_t num_residuals() const { return (uint32_t)expressions.size(); }

    double value_of(uint32_t var, const double* free_values) const {
        int32_t f = var_to_free[var];
        return f >= 0 ? free_values[f] : system_variables[var];
    }
    // subsystem.rs:93-104
    void calculate_residuals(const double* variables, double* residuals) const {
        uint32_t vi[8];
        double vals[8] = {0, 0, 0, 0, 0, 0, 0, 0}, grad[8];
        for (size_t row = 0; row < expressions.size(); row++) {
            const Expression& e = all_expressions[expressions[row]];
            int a = variable_indices(e, vi);
            for (int k = 0; k < a; k++) vals[k] = value_of(vi[k], variables);
            residuals[row] = compute_residual_and_gradient(e, vals, grad);
        }
    }
    // subsystem.rs:106-124 + expressions.rs:962-1090: dense row-major Jacobian (rows x free variables).
    // Entries are ASSIGNED per slot (expressions.rs:1003-1007) and the buffer is never cleared, so a
    // variable that fills two slots of a row keeps the later slot's value.
    void calculate_residuals_and_jacobian(const double* variables, double* residuals, double* jacobian) const {
        uint32_t vi[8];
        double vals[8] = {0, 0, 0, 0, 0, 0, 0, 0}, grad[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const size_t n = free_variables.size();
        for (size_t row = 0; row < expressions.size(); row++) {
            const Expression& e = all_expressions[expressions[row]];
            int a = variable_indices(e, vi);
            for (int k = 0; k < a; k++) vals[k] = value_of(vi[k], variables);
            residuals[row] = compute_residual_and_gradient(e, vals, grad);
            for (int k = 0; k < a; k++) {
                int32_t f = var_to_free[vi[k]];
                if (f >= 0) jacobian[row * n + (size_t)f] = grad[k];
            }
        }
    }
    // subsystem.rs:126-166
    void calculate_residuals_and_sparse_jacobian(const double* variables, double* residuals,
                                                 solvi::TripletMat& jac) const {
        uint32_t vi[8];
        double vals[8] = {0, 0, 0, 0, 0, 0, 0, 0}, grad[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (size_t row = 0; row < expressions.size(); row++) {
            const Expression& e = all_expressions[expressions[row]];
            int a = variable_indices(e, vi);
            for (int k = 0; k < a; k++) vals[k] = value_of(vi[k], variables);
            residuals[row] = compute_residual_and_gradient(e, vals, grad);
            for (int k = 0; k < a; k++) {
                int32_t f = var_to_free[vi[k]];
                if (f >= 0) jac.push_triplet(row, (size_t)f, grad[k]);
            }
        }
    }
};

enum ExitReason : uint32_t {
    EXIT_CONVERGED_RESIDUAL = 0,  // lm.rs:110-112
    EXIT_SMALL_STEP = 1,          // lm.rs:139-142
    EXIT_STALLED = 2,             // lm.rs:164-168
    EXIT_MAX_OUTER = 3,           // lm.rs:109
    EXIT_LAMBDA_OVERFLOW = 4,     // reference would spin forever
};

struct LmReport {
    uint32_t exit_reason = 0;
    uint32_t outer_iters = 0;     // outer iterations that entered the damping loop
    uint32_t factorizations = 0;  // inner iterations (one factorize + solve each)
    uint32_t accepted = 0;        // accepted steps
    double ssr = 0.0;             // sum of squared residuals at the returned variables
    double lambda = 0.0;          // final damping
    uint64_t trace_hash = 0;      // rolling base-3 hash of decisions: 1 accept, 2 reject, 0 unsolved
    std::string trace;            // 'A' accept, 'R' reject, 'U' unsolved (R diagonal exactly zero)
    // symbolic artefacts of the (only) SymbolicQr::build call, for the parity probes
    solvi::Structure jacobian_structure;
    solvi::SymbolicQr symbolic;
    std::vector<double> initial_residuals;  // before negation
};

inline uint64_t trace_push(uint64_t h, uint32_t code) { return h * 3u + code + 1u; }

// fiksi/src/solve/lm.rs:21-197
template <class Problem>  // solve/mod.rs:29-49 (trait Problem): Subsystem, ClusteredSystem
inline void levenberg_marquardt(const Problem& problem, double* variables, LmReport& rep,
                                bool keep_artifacts = false) {
    size_t nrows = problem.num_residuals(), ncols = problem.num_variables();
    std::vector<double> xs(variables, variables + ncols);
    std::vector<double> residuals(nrows, 0.), residuals_scratch(nrows, 0.), b_aug(nrows + ncols, 0.);
    solvi::TripletMat jac(nrows, ncols);
    problem.calculate_residuals_and_sparse_jacobian(xs.data(), residuals.data(), jac);
    if (keep_artifacts) rep.initial_residuals = residuals;
    for (double& r : residuals) r = -r;
    for (size_t idx = 0; idx < ncols; idx++) jac.push_triplet(nrows + idx, idx, 0.);
    solvi::SparseColMat csc = solvi::SparseColMat::from_triplet_mat(jac);
    solvi::SymbolicQr sym = solvi::SymbolicQr::build(csc.structure, solvi::QrOrdering::Colamd);
    solvi::Qr qr(sym);
    if (keep_artifacts) rep.jacobian_structure = csc.structure;

    double ssr = sum_squares(residuals.data(), nrows);
    double lambda = 0.5;
    rep.exit_reason = EXIT_MAX_OUTER;
    bool done = false;
    for (int outer = 0; outer < 100 && !done; outer++) {
        if (ssr < 1e-8) {
            rep.exit_reason = EXIT_CONVERGED_RESIDUAL;
            break;
        }
        rep.outer_iters++;
        for (;;) {
            if (!std::isfinite(lambda)) {
                rep.exit_reason = EXIT_LAMBDA_OVERFLOW;
                done = true;
                break;
            }
            double sl = std::sqrt(lambda);
            for (size_t idx = 0; idx < ncols; idx++)
                csc.values[csc.structure.column_pointers[idx + 1] - 1] = sl;
            qr.factorize(csc);
            rep.factorizations++;
            for (size_t i = 0; i < nrows; i++) b_aug[i] = residuals[i];
            for (size_t i = nrows; i < nrows + ncols; i++) b_aug[i] = 0.;
            bool solved = qr.solve_mut(b_aug.data());
            if (!solved) {
                lambda *= 8.;
                rep.trace.push_back('U');
                rep.trace_hash = trace_push(rep.trace_hash, 0);
                continue;
            }
            const double* delta = b_aug.data();
            if (sum_squares(delta, ncols) < 1e-12) {
                rep.exit_reason = EXIT_SMALL_STEP;
                done = true;
                break;
            }
            for (size_t idx = 0; idx < ncols; idx++) xs[idx] = variables[idx] + delta[idx];
            problem.calculate_residuals(xs.data(), residuals_scratch.data());
            double ssr_s = sum_squares(residuals_scratch.data(), nrows);
            if (std::getenv("ORC_PROGRESS"))  // long golden-generation runs only
                std::fprintf(stderr, "[oracle lm] factorization %u lambda %a ssr %a -> %a\n", rep.factorizations, lambda, ssr, ssr_s);
            if (ssr_s < ssr) {
                lambda *= 0.125;
                if (lambda < 1e-50) lambda = 1e-50;
                for (size_t idx = 0; idx < ncols; idx++) variables[idx] = xs[idx];
                rep.accepted++;
                rep.trace.push_back('A');
                rep.trace_hash = trace_push(rep.trace_hash, 1);
                if ((ssr - ssr_s) / ssr <= 1e-6) {
                    ssr = ssr_s;  // report the value at the returned variables
                    rep.exit_reason = EXIT_STALLED;
                    done = true;
                    break;
                }
                ssr = ssr_s;
                jac.clear();
                problem.calculate_residuals_and_sparse_jacobian(xs.data(), residuals.data(), jac);
                for (double& r : residuals) r = -r;
                for (size_t idx = 0; idx < ncols; idx++) jac.push_triplet(nrows + idx, idx, 0.);
                csc = solvi::SparseColMat::from_triplet_mat(jac);
                break;
            } else {
                lambda *= 2.;
                rep.trace.push_back('R');
                rep.trace_hash = trace_push(rep.trace_hash, 2);
            }
        }
    }
    rep.ssr = ssr;
    rep.lambda = lambda;
    if (keep_artifacts) rep.symbolic = std::move(sym);
}

// ---------------------------------------------------------------------------------------------
// Model layer: graph, system, assemble::solve
// ---------------------------------------------------------------------------------------------

// fiksi/src/graph.rs:120-259 (incremental connected components, stale indices preserved)
struct ConnectedComponent {
    std::set<uint32_t> elements, constraints;
};
struct Graph {
    size_t n_elements = 0, n_constraints = 0;
    std::vector<uint32_t> element_cc;  // 0 == None, otherwise 1-based component index
    std::vector<ConnectedComponent> components;
    // graph.rs:98-117: Element { dof, incident_constraints }, Constraint { valency, incident_elements }
    std::vector<int16_t> dof, valency;
    std::vector<std::vector<uint32_t>> incident_constraints, incident_elements;

    uint32_t add_element(int16_t element_dof = 0) {
        element_cc.push_back(0);
        dof.push_back(element_dof);
        incident_constraints.emplace_back();
        return (uint32_t)n_elements++;
    }
    // graph.rs:178-225
    void merge_connected_components(uint32_t constraint, const uint32_t* elements, int n) {
        uint32_t target = 0;
        size_t size_largest = 0;
        for (int k = 0; k < n; k++) {
            uint32_t ci = element_cc[elements[k]];
            if (ci != 0) {
                const ConnectedComponent& c = components[ci - 1];
                if (c.elements.size() > size_largest) {
                    target = ci;
                    size_largest = c.elements.size();
                }
            }
        }
        if (target == 0) {
            components.emplace_back();
            target = (uint32_t)components.size();
        }
        ConnectedComponent tc = std::move(components[target - 1]);
        components[target - 1] = ConnectedComponent();
        for (int k = 0; k < n; k++) {
            uint32_t ci = element_cc[elements[k]];
            if (ci != 0) {
                ConnectedComponent c = std::move(components[ci - 1]);
                components[ci - 1] = ConnectedComponent();
                tc.elements.insert(c.elements.begin(), c.elements.end());
                tc.constraints.insert(c.constraints.begin(), c.constraints.end());
            } else {
                tc.elements.insert(elements[k]);
            }
            element_cc[elements[k]] = target;
        }
        tc.constraints.insert(constraint);
        components[target - 1] = std::move(tc);
    }
    // graph.rs:235-254
    uint32_t add_constraint(const uint32_t* elements, int n, int16_t constraint_valency = 1) {
        uint32_t id = (uint32_t)n_constraints++;
        merge_connected_components(id, elements, n);
        for (int k = 0; k < n; k++) incident_constraints[elements[k]].push_back(id);  // :227-233
        valency.push_back(constraint_valency);
        incident_elements.emplace_back(elements, elements + n);
        return id;
    }
};

enum ElementTag : uint8_t { ELength = 0, EPoint = 1, ELine = 2, ECircle = 3 };
struct EncodedElement {
    uint8_t tag;
    uint32_t a, b;  // Length{idx=a}; Point{idx=a}; Line{p1=a,p2=b}; Circle{center=a,radius=b}
};
struct EncodedConstraint {
    uint8_t tag;  // ConstraintTag order, constraints/mod.rs:893-905 (0 = PointPointCoincidence)
    uint32_t expressions_idx;
};
// constraints/mod.rs:907-924,948-990: only PointPointCoincidence has valency 2.
inline uint8_t valency_of(uint8_t constraint_tag) { return constraint_tag == 0 ? 2 : 1; }


// ---- Decomposer::RecursiveAssembly: the recombination plan (analyze/graph/recursive_assembly.rs) ------
// The reference keeps its vertex / edge / subgraph / frontier sets in `hashbrown` HashSets and iterates
// them (recursive_assembly.rs:219,228,267,399,441,586,606); that order depends on the hasher's
// per-process seed and is unspecified, so the plan of the reference is not a function of its input.  The
// restatement iterates every such set in ASCENDING id order (std::set): one of the orders the reference
// can take.  PARITY UNPINNED beyond that choice: the reference's only test of this decomposer is the
// residual threshold of tests/triangles.rs:8-37 (tests/test_oracle_recursive.py checks it here).
struct RecombinationStep {  // recursive_assembly.rs:75-114
    std::vector<uint32_t> constraints, elements, free_elements;
    std::map<uint32_t, std::vector<uint32_t>> on_frontiers;       // element -> cluster keys
    std::map<uint32_t, std::vector<uint32_t>> owned_elements;     // cluster key -> elements
    std::map<uint32_t, std::vector<uint32_t>> frontier_elements;  // cluster key -> elements
};

// recursive_assembly.rs:499-645 (dense_bfs::<D>): the first subgraph of the breadth-first enumeration with
// `dof > -(D + 1)` that is not blocked.  (With non-negative element dofs the test holds for every subgraph the
// search reaches, so this returns the first unblocked connected subgraph of at least two vertices: the
// reference's behaviour, restated as it is.)
template <int D>
inline bool dense_bfs(const Graph& graph, const std::vector<std::set<uint32_t>>& blocked_subgraphs,
                      const std::set<uint32_t>& available_edges, const std::set<uint32_t>& vertices, std::set<uint32_t>& out) {
    const int k = -(D + 1);
    auto all_in = [&](uint32_t edge, const std::set<uint32_t>& sub) {
        for (uint32_t u : graph.incident_elements[edge])
            if (!sub.count(u)) return false;
        return true;
    };
    auto additional_valency = [&](const std::set<uint32_t>& next_subgraph, uint32_t new_vertex) {  // :510-532
        int add = 0;
        for (uint32_t edge : graph.incident_constraints[new_vertex])
            if (available_edges.count(edge) && all_in(edge, next_subgraph)) add += graph.valency[edge];
        return add;
    };
    auto extend_adjacent_vertices = [&](std::set<uint32_t>& adjacent, uint32_t from_vertex, const std::set<uint32_t>& sub) {  // :536-557
        for (uint32_t edge : graph.incident_constraints[from_vertex]) {
            if (!available_edges.count(edge)) continue;
            for (uint32_t e : graph.incident_elements[edge])
                if (vertices.count(e) && !sub.count(e)) adjacent.insert(e);
        }
    };
    struct SubgraphState {  // :559-574
        std::set<uint32_t> subgraph;
        int dof;
        std::set<uint32_t> adjacent_vertices;
    };
    std::deque<SubgraphState> queue;
    for (uint32_t vertex : vertices) {  // :578-599
        SubgraphState st;
        st.subgraph.insert(vertex);
        extend_adjacent_vertices(st.adjacent_vertices, vertex, st.subgraph);
        st.dof = graph.dof[vertex];
        queue.push_back(std::move(st));
    }
    while (!queue.empty()) {  // :601-641
        SubgraphState cur = std::move(queue.front());
        queue.pop_front();
        for (uint32_t vertex : cur.adjacent_vertices) {
            std::set<uint32_t> next_subgraph = cur.subgraph;
            next_subgraph.insert(vertex);
            const int valency = additional_valency(next_subgraph, vertex);
            const int next_dof = cur.dof + graph.dof[vertex] - valency;
            bool blocked = false;
            for (const std::set<uint32_t>& b : blocked_subgraphs) blocked = blocked || b == next_subgraph;
            if (!blocked && next_dof > k) {
                out = std::move(next_subgraph);
                return true;
            }
            SubgraphState nx;
            nx.adjacent_vertices = cur.adjacent_vertices;
            nx.adjacent_vertices.erase(vertex);
            extend_adjacent_vertices(nx.adjacent_vertices, vertex, next_subgraph);
            nx.subgraph = std::move(next_subgraph);
            nx.dof = next_dof;
            queue.push_back(std::move(nx));
        }
    }
    return false;
}

// recursive_assembly.rs:165-483 (decompose::<D>), the Modified Frontier Algorithm as the reference runs it.
template <int D>
inline std::vector<RecombinationStep> decompose(Graph graph, const std::set<uint32_t>& vertices_in, const std::set<uint32_t>& edges_in) {
    const uint32_t num_real_constraints = (uint32_t)graph.n_constraints, num_real_elements = (uint32_t)graph.n_elements;
    std::set<uint32_t> vertices = vertices_in, available_edges = edges_in, constraints_handled, vertices_handled;
    std::map<uint32_t, std::vector<uint32_t>> on_frontiers, owned_elements, frontier_elements;
    std::map<uint32_t, uint32_t> owning_cluster;
    std::vector<std::set<uint32_t>> blocked_clusters;
    std::vector<RecombinationStep> plan;
    std::vector<uint32_t> step_constraints, step_fixes_elements;
    for (uint32_t step = 0;; step++) {
        const uint32_t cluster_key = step;
        std::set<uint32_t> subgraph;
        if (!dense_bfs<D>(graph, blocked_clusters, available_edges, vertices, subgraph)) {  // :209-251
            RecombinationStep rs;
            for (uint32_t edge : available_edges)
                if (edge < num_real_constraints && !constraints_handled.count(edge)) rs.constraints.push_back(edge);
            for (uint32_t vertex : vertices)
                if (vertex < num_real_elements && !vertices_handled.count(vertex)) rs.free_elements.push_back(vertex);
            if (!rs.constraints.empty()) {
                for (uint32_t vertex : vertices)
                    if (vertex < num_real_elements) rs.elements.push_back(vertex);
                rs.on_frontiers = on_frontiers;
                rs.owned_elements = owned_elements;
                rs.frontier_elements = frontier_elements;
                plan.push_back(std::move(rs));
            }
            break;
        }
        std::vector<uint32_t> core, real_elements;  // :263-310
        std::set<uint32_t> frontier;
        for (uint32_t vertex : subgraph) {
            if (vertex < num_real_elements) real_elements.push_back(vertex);
            if (vertex < num_real_elements && !vertices_handled.count(vertex)) {
                step_fixes_elements.push_back(vertex);
                vertices_handled.insert(vertex);
                owning_cluster[vertex] = cluster_key;
            }
            bool frontier_vertex = false;
            for (uint32_t edge_id : graph.incident_constraints[vertex]) {
                if (!available_edges.count(edge_id)) continue;
                bool inside = true;
                for (uint32_t e : graph.incident_elements[edge_id]) inside = inside && subgraph.count(e);
                if (inside) {
                    if (edge_id < num_real_constraints && !constraints_handled.count(edge_id)) {
                        step_constraints.push_back(edge_id);
                        constraints_handled.insert(edge_id);
                    }
                } else {
                    frontier_vertex = true;
                }
            }
            if (!frontier_vertex) core.push_back(vertex);
            else frontier.insert(vertex);
        }
        if (!step_constraints.empty()) {  // :312-322
            RecombinationStep rs;
            rs.constraints = std::move(step_constraints);
            step_constraints.clear();
            rs.elements = real_elements;
            rs.free_elements = step_fixes_elements;
            rs.on_frontiers = on_frontiers;
            rs.owned_elements = owned_elements;
            rs.frontier_elements = frontier_elements;
            plan.push_back(std::move(rs));
        }
        if (!core.empty() || !step_fixes_elements.empty()) {  // :324-337
            owned_elements[cluster_key] = std::move(step_fixes_elements);
            step_fixes_elements.clear();
        }
        auto in_core = [&](uint32_t e) { return std::find(core.begin(), core.end(), e) != core.end(); };
        for (uint32_t vertex : core) {  // :340-387
            if (vertex < num_real_elements) {
                for (uint32_t edge_id : graph.incident_constraints[vertex]) {
                    bool inside = true;
                    for (uint32_t e : graph.incident_elements[edge_id]) inside = inside && in_core(e);
                    if (inside) available_edges.erase(edge_id);
                }
            }
            const uint32_t old_cluster_key = owning_cluster.at(vertex);  // (`insert(..).unwrap()`: the vertex has an owner)
            owning_cluster[vertex] = cluster_key;
            if (old_cluster_key != cluster_key) {
                std::vector<uint32_t> old_owned = std::move(owned_elements.at(old_cluster_key));
                owned_elements.erase(old_cluster_key);
                for (uint32_t v : old_owned) owning_cluster[v] = cluster_key;
                std::vector<uint32_t>& mine = owned_elements.at(cluster_key);
                mine.insert(mine.end(), old_owned.begin(), old_owned.end());
                std::vector<uint32_t> old_frontier = std::move(frontier_elements.at(old_cluster_key));
                frontier_elements.erase(old_cluster_key);
                for (uint32_t element : old_frontier) {
                    auto it = on_frontiers.find(element);
                    if (it == on_frontiers.end()) continue;
                    std::vector<uint32_t>& of = it->second;
                    const size_t idx = (size_t)(std::find(of.begin(), of.end(), old_cluster_key) - of.begin());
                    if (idx == of.size()) throw std::out_of_range("position(..).unwrap() on None, recursive_assembly.rs:377-381");
                    of[idx] = of.back();  // swap_remove
                    of.pop_back();
                }
            }
            on_frontiers.erase(vertex);
        }
        for (uint32_t vertex : frontier) {  // :388-397
            on_frontiers[vertex].push_back(cluster_key);
            if (vertex < num_real_elements) frontier_elements[cluster_key].push_back(vertex);
        }
        if (subgraph.size() - frontier.size() <= 1) {  // :409-419
            blocked_clusters.push_back(subgraph);
            continue;
        }
        for (uint32_t vertex : core) vertices.erase(vertex);  // :423-428
        const uint32_t core_vertex = graph.add_element(0);
        owning_cluster[core_vertex] = cluster_key;
        vertices.insert(core_vertex);
        int total_frontier_vertex_dof = 0, total_incoming_edge_valency = 0;
        for (uint32_t vertex : frontier) {  // :432-472
            total_frontier_vertex_dof += graph.dof[vertex];
            int binary_edge_cluster_valency = 0;
            const std::vector<uint32_t> incident = graph.incident_constraints[vertex];  // (add_constraint below appends to it)
            for (uint32_t edge_id : incident) {
                if (!available_edges.count(edge_id)) continue;
                bool inside = true;
                for (uint32_t e : graph.incident_elements[edge_id]) inside = inside && subgraph.count(e);
                if (!inside) continue;
                // graph.rs:58-93 merge_elements: every element that is not on the frontier becomes `core_vertex` (once)
                std::vector<uint32_t> merged;
                bool did = false;
                for (uint32_t e : graph.incident_elements[edge_id]) {
                    if (!frontier.count(e)) {
                        if (!did) merged.push_back(core_vertex);
                        did = true;
                    } else {
                        merged.push_back(e);
                    }
                }
                if (merged.size() == 2) {
                    binary_edge_cluster_valency += graph.valency[edge_id];
                    available_edges.erase(edge_id);
                } else {
                    graph.incident_elements[edge_id] = merged;
                }
            }
            if (binary_edge_cluster_valency > 0) {
                const uint32_t inc[2] = {vertex, core_vertex};
                const uint32_t cluster_edge = graph.add_constraint(inc, 2, (int16_t)binary_edge_cluster_valency);
                available_edges.insert(cluster_edge);
                total_incoming_edge_valency += binary_edge_cluster_valency;
            }
        }
        if (total_incoming_edge_valency > 0) graph.dof[core_vertex] = (int16_t)(total_frontier_vertex_dof - total_incoming_edge_valency - D);  // :474-479
        else vertices.erase(core_vertex);
    }
    return plan;
}

struct SolvingOptions {
    bool perturb = true;  // lib.rs:232-236
};

// ---- L-BFGS (fiksi/src/solve/lbfgs.rs) -------------------------------------------------------------
// Exit reasons of the restatement (the reference returns `()`):
enum LbfgsExit : uint32_t {
    LBFGS_EXIT_INITIAL = 0,     // initial sum of squares < 1e-4              lbfgs.rs:53-56
    LBFGS_EXIT_CONVERGED = 1,   // |change of the sum of squares| < 1e-10     lbfgs.rs:177-179
    LBFGS_EXIT_RESIDUAL = 2,    // sum of squares < 1e-6                      lbfgs.rs:180-182
    LBFGS_EXIT_MAX_ITER = 3,    // 100 iterations                             lbfgs.rs:77
    LBFGS_EXIT_GUARD = 4,       // the unbounded bisection of `update` (U3, lbfgs.rs:338-351) exceeded 200 steps:
                                // the reference would keep spinning (NaN / non-descent direction)
};

struct LbfgsReport {
    uint32_t exit_reason = 0;
    uint32_t iterations = 0;    // line searches performed
    uint32_t evaluations = 0;   // calls of calculate_phi (residual + Jacobian + gradient evaluations)
    double ssr = 0.0;           // sum of squared residuals at the returned variables
    double step = 0.0;          // last accepted step size
    uint64_t trace_hash = 0;    // h = 31 h + (evaluations of the line search) per iteration
};

namespace lbfgs_detail {
struct Param { double p, phi, dphi; };  // lbfgs.rs:249-256

// lbfgs.rs:201-212: gradient[i] = sum over rows c (ascending) of J[c][i] * r[c], starting from 0.0
inline void compute_gradient(const double* jacobian, const double* residuals, double* gradient, size_t nvars, size_t nexpr) {
    for (size_t i = 0; i < nvars; i++) {
        double g = 0.0;
        for (size_t c = 0; c < nexpr; c++) g += jacobian[c * nvars + i] * residuals[c];
        gradient[i] = g;
    }
}
inline double dot_product(const double* a, const double* b, size_t n) {  // lbfgs.rs:214-216 (sequential sum from 0.0)
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
inline double ssq_seq(const double* v, size_t n) {  // utils.rs:12-20
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += v[i] * v[i];
    return s;
}

struct LineSearch {  // lbfgs.rs:218-506 (mod hager_zhang)
    static constexpr double DELTA = 1e-4, SIGMA = 0.9, EPSILON = 1e-6, THETA = 0.5, GAMMA = 0.66;
    static constexpr int MAX_ITERATIONS = 100, U3_GUARD = 200;
    const Subsystem& problem;
    const double* variables;
    double* variables_scratch;
    double* jacobian;
    double* residuals;
    double* gradient;
    const double* direction;
    size_t n, m;
    double phi0 = 0, dphi0 = 0;
    uint32_t evaluations = 0;
    bool guard_hit = false;

    Param calculate_phi(double p) {  // :270-286
        for (size_t idx = 0; idx < n; idx++) variables_scratch[idx] = variables[idx] + p * direction[idx];
        problem.calculate_residuals_and_jacobian(variables_scratch, residuals, jacobian);
        compute_gradient(jacobian, residuals, gradient, n, m);
        evaluations++;
        Param out{p, ssq_seq(residuals, m), dot_product(gradient, direction, n)};
#ifdef FK_LBFGS_DEBUG
        fprintf(stderr, "cpu p=%a phi=%a dphi=%a\n", p, out.phi, out.dphi);
#endif
        return out;
    }
    static double secant(Param a, Param b) { return (a.p * b.dphi - b.p * a.dphi) / (b.dphi - a.dphi); }  // :291-293
    bool satisfies_wolfe(Param c) const {  // :307-322
        if ((c.phi <= phi0 + c.p * (DELTA * dphi0)) && (c.dphi >= SIGMA * dphi0)) return true;
        if (c.phi <= phi0 + EPSILON && (2. * DELTA - 1.) * dphi0 >= c.dphi && c.dphi >= SIGMA * dphi0) return true;
        return false;
    }
    void update(Param a, Param b, Param c, Param& oa, Param& ob) {  // :325-353
        if (c.p < a.p || c.p > b.p) { oa = a; ob = b; return; }       // U0
        if (c.dphi >= 0.) { oa = a; ob = c; return; }                  // U1
        if (c.phi <= phi0 + EPSILON) { oa = c; ob = b; return; }       // U2
        Param aa = a, bb = c;                                          // U3
        for (int it = 0;; it++) {
            if (it >= U3_GUARD) { guard_hit = true; oa = aa; ob = bb; return; }
            Param d = calculate_phi((1. - THETA) * aa.p + THETA * bb.p);
            if (d.dphi >= 0.) { oa = aa; ob = d; return; }
            else if (d.phi <= phi0 + EPSILON) aa = d;
            else bb = d;
        }
    }
    // :360-398.  Returns true (and c_out) when a point satisfying the Wolfe conditions was found.
    bool secant2(Param a, Param b, Param& c_out, Param& oa, Param& ob) {
        Param c = calculate_phi(secant(a, b));
        if (satisfies_wolfe(c)) { c_out = c; return true; }
        Param a_, b_;
        update(a, b, c, a_, b_);
        if (guard_hit) { oa = a_; ob = b_; return false; }
        if (c.p == b_.p) {
            Param c_ = calculate_phi(secant(b, b_));
            if (satisfies_wolfe(c_)) { c_out = c_; return true; }
            update(a_, b_, c_, oa, ob);
            return false;
        } else if (c.p == a_.p) {
            Param c_ = calculate_phi(secant(a, a_));
            if (satisfies_wolfe(c_)) { c_out = c_; return true; }
            update(a_, b_, c_, oa, ob);
            return false;
        }
        oa = a_; ob = b_;
        return false;
    }
    Param run() {  // :447-458, bracket :401-410, search :414-443
        Param c = calculate_phi(1.);
        if (satisfies_wolfe(c)) return c;
        Param a{0., phi0, dphi0};
        Param b = calculate_phi(5.);
        for (int it = 0; it < MAX_ITERATIONS; it++) {
            Param found, a_, b_;
            if (secant2(a, b, found, a_, b_)) return found;
            if (guard_hit) break;
            if (b_.p - a_.p > GAMMA * (b.p - a.p)) {
                c = calculate_phi(0.5 * (a.p + b.p));
                if (satisfies_wolfe(c)) return c;
                Param na, nb;
                update(a, b, c, na, nb);
                a = na; b = nb;
                if (guard_hit) break;
            } else {
                a = a_; b = b_;
            }
        }
        return calculate_phi(c.p);  // :440-442: leave the buffers at c
    }
};
}  // namespace lbfgs_detail

// fiksi/src/solve/lbfgs.rs:20-195
inline void lbfgs(const Subsystem& problem, double* variables, LbfgsReport& rep) {
    using namespace lbfgs_detail;
    const int MAX_HISTORY = 5, MAX_ITERATIONS = 100;
    const double CONVERGENCE_THRESHOLD = 1e-10, RESIDUAL_THRESHOLD = 1e-6;
    const size_t n = problem.num_variables(), m = problem.num_residuals();
    std::vector<double> residuals(m, 0.), jacobian(m * n, 0.), gradient(n, 0.);
    problem.calculate_residuals_and_jacobian(variables, residuals.data(), jacobian.data());
    rep = LbfgsReport();
    rep.evaluations = 1;
    double prev_ssr = ssq_seq(residuals.data(), m);
    rep.ssr = prev_ssr;
    if (prev_ssr < 1e-4) { rep.exit_reason = LBFGS_EXIT_INITIAL; return; }
    compute_gradient(jacobian.data(), residuals.data(), gradient.data(), n, m);
    std::vector<double> s_history(n * MAX_HISTORY, 0.), y_history(n * MAX_HISTORY, 0.), rho_history(MAX_HISTORY, 0.), alpha(MAX_HISTORY, 0.);
    std::vector<double> direction(n, 0.), variables_scratch(n, 0.);
    rep.exit_reason = LBFGS_EXIT_MAX_ITER;
    for (int k = 0; k < MAX_ITERATIONS; k++) {
        const int history_len = std::min(k, MAX_HISTORY);
        for (size_t j = 0; j < n; j++) direction[j] = gradient[j];
        for (int i = history_len - 1; i >= 0; i--) {  // :83-99
            const size_t h = (size_t)((k + i) % MAX_HISTORY);
            const double* s_i = &s_history[h * n];
            const double* y_i = &y_history[h * n];
            double dp = 0.;
            for (size_t j = 0; j < n; j++) dp += s_i[j] * direction[j];
            alpha[i] = rho_history[h] * dp;
            for (size_t j = 0; j < n; j++) direction[j] -= alpha[i] * y_i[j];
        }
        if (k > 0) {  // :101-121
            const size_t hp = (size_t)((k - 1) % MAX_HISTORY);
            double s_dot_y = 0., y_dot_y = 0.;
            for (size_t j = 0; j < n; j++) {
                s_dot_y += s_history[hp * n + j] * y_history[hp * n + j];
                y_dot_y += y_history[hp * n + j] * y_history[hp * n + j];
            }
            if (y_dot_y > 0.) {
                const double scale = s_dot_y / y_dot_y;
                for (size_t j = 0; j < n; j++) direction[j] *= scale;
            }
        }
        for (int i = 0; i < history_len; i++) {  // :123-139
            const size_t h = (size_t)((k + i) % MAX_HISTORY);
            double dp = 0.;
            for (size_t j = 0; j < n; j++) dp += y_history[h * n + j] * direction[j];
            const double beta = rho_history[h] * dp;
            for (size_t j = 0; j < n; j++) direction[j] += s_history[h * n + j] * (alpha[i] - beta);
        }
        for (size_t j = 0; j < n; j++) direction[j] *= -1.;  // :141-143
        const size_t h = (size_t)(k % MAX_HISTORY);
        for (size_t j = 0; j < n; j++) y_history[h * n + j] = gradient[j];  // :149-150
        for (size_t j = 0; j < n; j++) variables_scratch[j] = variables[j];
        LineSearch ls{problem, variables, variables_scratch.data(), jacobian.data(), residuals.data(), gradient.data(), direction.data(), n, m};
        ls.phi0 = prev_ssr;                                              // :489
        ls.dphi0 = dot_product(gradient.data(), direction.data(), n);    // :490
        const Param res = ls.run();
        rep.evaluations += ls.evaluations;
        rep.iterations++;
        rep.trace_hash = rep.trace_hash * 31u + ls.evaluations;
        for (size_t j = 0; j < n; j++) variables[j] = variables_scratch[j];  // :169
        rep.step = res.p;
        rep.ssr = res.phi;
        if (ls.guard_hit) { rep.exit_reason = LBFGS_EXIT_GUARD; return; }
        double s_dot_y = 0.;  // :171-180
        for (size_t i = 0; i < n; i++) {
            s_history[h * n + i] = res.p * direction[i];
            y_history[h * n + i] = gradient[i] - y_history[h * n + i];
            s_dot_y += s_history[h * n + i] * y_history[h * n + i];
        }
        rho_history[h] = 1.0 / s_dot_y;
        if (std::fabs(prev_ssr - res.phi) < CONVERGENCE_THRESHOLD) { rep.exit_reason = LBFGS_EXIT_CONVERGED; return; }
        if (res.phi < RESIDUAL_THRESHOLD) { rep.exit_reason = LBFGS_EXIT_RESIDUAL; return; }
        prev_ssr = res.phi;
    }
}

// ---- System::analyze: over-constraint detection (fiksi/src/analyze/numerical/mod.rs) ---------------
// analyze/numerical/mod.rs:33-117.  Row-by-row Gauss-Jordan elimination with column swaps tracked in
// `column_indices`; returns for every row whether it increased the rank.  Rows beyond
// min(nrows, ncols) are never examined and stay `false` (as in the reference loop bound, :64).
static const double ANALYZE_EPSILON = 1e-8;  // :8
inline std::vector<uint8_t> incremental_gauss_jordan_elimination(std::vector<double>& matrix, size_t nrows, size_t ncols,
                                                                 std::vector<size_t>& column_indices) {
    const size_t constraints = nrows, variables = ncols;
    std::vector<uint8_t> constraint_increases_rank(constraints, 0);
    size_t current_col = 0;
    for (size_t row = 0; row < std::min(constraints, variables); row++) {
        size_t rank = 0;
        for (size_t row_idx = 0; row_idx < row; row_idx++) {  // :66-77
            const size_t column_idx = column_indices[rank];
            const double factor = matrix[row * variables + column_idx];
            for (size_t col = 0; col < variables; col++)
                matrix[row * variables + col] -= factor * matrix[row_idx * variables + col];
            if (constraint_increases_rank[row_idx]) rank += 1;
        }
        bool pivot_found = false;  // :81-90: first entry above EPSILON in the remaining column order
        for (size_t idx = current_col; idx < variables; idx++) {
            const size_t real_idx = column_indices[idx];
            if (std::fabs(matrix[row * variables + real_idx]) > ANALYZE_EPSILON) {
                std::swap(column_indices[current_col], column_indices[idx]);
                pivot_found = true;
                break;
            }
        }
        if (!pivot_found) continue;  // :93-95
        const double factor = matrix[row * variables + column_indices[current_col]];
        for (size_t col = 0; col < variables; col++) matrix[row * variables + col] *= 1. / factor;  // :97-100
        const size_t column_idx = column_indices[current_col];
        for (size_t row_idx = 0; row_idx < row; row_idx++) {  // :104-110
            const double f2 = matrix[row_idx * variables + column_idx];
            for (size_t col = 0; col < variables; col++)
                matrix[row_idx * variables + col] -= f2 * matrix[row * variables + col];
        }
        current_col += 1;
        constraint_increases_rank[row] = 1;
    }
    return constraint_increases_rank;
}

// analyze/numerical/mod.rs:123-147: dense Jacobian over ALL variables (IdentityVariableMap: every
// variable is free, fixed ones included, :124), gradient entries ASSIGNED per slot
// (expressions.rs:1003-1007: a variable that fills two slots keeps the later slot's value), then
// the elimination.  Returns the per-expression "independent" flags.
inline std::vector<uint8_t> analyze_expressions(const std::vector<double>& variables, const std::vector<Expression>& expressions) {
    const size_t m = expressions.size(), n = variables.size();
    std::vector<double> jacobian(m * n, 0.0);
    for (size_t row = 0; row < m; row++) {
        uint32_t vi[8];
        double vals[8] = {0, 0, 0, 0, 0, 0, 0, 0}, grad[8];
        const int a = variable_indices(expressions[row], vi);
        for (int k = 0; k < a; k++) vals[k] = variables[vi[k]];
        compute_residual_and_gradient(expressions[row], vals, grad);
        for (int k = 0; k < a; k++) jacobian[row * n + vi[k]] = grad[k];
    }
    std::vector<size_t> column_pivots(n);
    for (size_t k = 0; k < n; k++) column_pivots[k] = k;
    return incremental_gauss_jordan_elimination(jacobian, m, n, column_pivots);
}

// ---- Decomposer::SinglePass: equation graph, maximum matching, strongly connected expressions ----------
// fiksi/src/analyze/graph/equations.rs.  `variables[v]` lists the expressions of variable v in
// insertion order, `expressions[e]` the variables of expression e (variable_indices order).
struct ExpressionGraph {  // equations.rs:146-182
    std::vector<std::vector<uint32_t>> variables, expressions;
    void insert_variables(int n) { for (int k = 0; k < n; k++) variables.emplace_back(); }
    void insert_expression(const uint32_t* vars, int n) {
        uint32_t id = (uint32_t)expressions.size();
        expressions.emplace_back(vars, vars + n);
        for (uint32_t v : expressions.back()) variables[v].push_back(id);
    }
};

// Minimal insertion-ordered map (the reference uses indexmap::IndexMap, equations.rs:96-101): inserting
// an existing key replaces the value and keeps the position.
struct OrderedMapU32 {
    std::vector<std::pair<uint32_t, uint32_t>> items;
    std::map<uint32_t, size_t> pos;
    bool contains(uint32_t k) const { return pos.count(k) != 0; }
    const uint32_t* get(uint32_t k) const {
        auto it = pos.find(k);
        return it == pos.end() ? nullptr : &items[it->second].second;
    }
    void insert(uint32_t k, uint32_t v) {
        auto it = pos.find(k);
        if (it == pos.end()) { pos[k] = items.size(); items.push_back({k, v}); }
        else items[it->second].second = v;
    }
};

struct Matching { OrderedMapU32 a_to_b, b_to_a; };  // equations.rs:96-101

struct StronglyConnectedExpressions {  // equations.rs:154-157
    std::vector<uint32_t> free_variables, expressions;
};

struct SinglePassPlanner {
    const ExpressionGraph& graph;
    const std::set<uint32_t>& free_variables;  // vertices of set A (ascending, BTreeSet)
    SinglePassPlanner(const ExpressionGraph& g, const std::set<uint32_t>& f) : graph(g), free_variables(f) {}
    Matching matching;
    OrderedMapU32 distance;  // keyed by free variable, insertion order == ascending
    uint32_t dummy_a_distance = UINT32_MAX;

    static uint32_t sat_add1(uint32_t v) { return v == UINT32_MAX ? v : v + 1; }

    bool bfs() {  // equations.rs:325-369
        std::deque<uint32_t> queue;
        for (auto& kv : distance.items) {
            if (matching.a_to_b.contains(kv.first)) kv.second = UINT32_MAX;
            else { kv.second = 0; queue.push_back(kv.first); }
        }
        dummy_a_distance = UINT32_MAX;
        while (!queue.empty()) {
            uint32_t a = queue.front();
            queue.pop_front();
            uint32_t a_distance = *distance.get(a);
            if (a_distance >= dummy_a_distance) continue;
            uint32_t new_dist = sat_add1(a_distance);
            for (uint32_t b : graph.variables[a]) {
                const uint32_t* matched_a = matching.b_to_a.get(b);
                if (!matched_a) {
                    if (dummy_a_distance == UINT32_MAX) dummy_a_distance = new_dist;
                } else if (*distance.get(*matched_a) == UINT32_MAX) {
                    distance.insert(*matched_a, new_dist);
                    queue.push_back(*matched_a);
                }
            }
        }
        return dummy_a_distance != UINT32_MAX;
    }
    bool dfs(uint32_t a) {  // equations.rs:371-402
        uint32_t a_distance_plus_one = sat_add1(*distance.get(a));
        for (uint32_t b : graph.variables[a]) {
            const uint32_t* matched_a = matching.b_to_a.get(b);
            if (!matched_a) {
                if (dummy_a_distance == a_distance_plus_one) {
                    matching.a_to_b.insert(a, b);
                    matching.b_to_a.insert(b, a);
                    return true;
                }
            } else {
                uint32_t ma = *matched_a;
                if (*distance.get(ma) == a_distance_plus_one && dfs(ma)) {
                    matching.a_to_b.insert(a, b);
                    matching.b_to_a.insert(b, a);
                    return true;
                }
            }
        }
        distance.insert(a, UINT32_MAX);
        return false;
    }
    void find_maximum_matching() {  // equations.rs:301-322
        for (uint32_t a : free_variables) distance.insert(a, UINT32_MAX);
        while (bfs())
            for (uint32_t a : free_variables)
                if (!matching.a_to_b.contains(a)) dfs(a);
    }
    // MatchedBipartiteGraph::neighbors, equations.rs:438-448
    std::vector<uint32_t> neighbors(uint32_t vertex) const {
        std::vector<uint32_t> out;
        uint32_t matched_a = *matching.b_to_a.get(vertex);
        for (uint32_t a : graph.expressions[vertex]) {
            if (!free_variables.count(a)) continue;  // MaskedExpressionGraph::neighbors_of_b, :272-286
            if (!(a == matched_a || !matching.a_to_b.contains(a))) continue;
            for (uint32_t b : graph.variables[a])
                if (b != vertex && matching.b_to_a.contains(b)) out.push_back(b);
        }
        return out;
    }
    // tarjan::tarjan_pearce + visit, equations.rs:470-567
    uint32_t index = 1, c = 0;
    std::map<uint32_t, uint32_t> root_index;
    std::vector<uint32_t> stack;
    std::vector<std::vector<uint32_t>> sccs;
    void visit(uint32_t vertex) {
        bool root = true;
        uint32_t vertex_index = index;
        root_index[vertex] = vertex_index;
        index += 1;
        for (uint32_t neighbor : neighbors(vertex)) {
            if (!root_index.count(neighbor)) visit(neighbor);
            uint32_t neighbor_index = root_index[neighbor];
            if (neighbor_index < vertex_index) {
                vertex_index = neighbor_index;
                root_index[vertex] = vertex_index;
                root = false;
            }
        }
        if (root) {
            std::vector<uint32_t> scc{vertex};
            index -= 1;
            while (!stack.empty()) {
                uint32_t top = stack.back();
                if (vertex_index > root_index[top]) break;
                stack.pop_back();
                scc.push_back(top);
                root_index[top] = c;
                index -= 1;
            }
            vertex_index = c;
            root_index[vertex] = vertex_index;
            c = c - 1;  // wrapping, :560
            sccs.push_back(std::move(scc));
        } else {
            stack.push_back(vertex);
        }
    }
    // find_strongly_connected_expressions, equations.rs:186-221.  The reference collects an SCC's free
    // variables in a HashSet and hands them out in hash order (unspecified); this restatement sorts them
    // ascending (the order only permutes the columns of the sub-problem).
    std::vector<StronglyConnectedExpressions> plan() {
        find_maximum_matching();
        c = (uint32_t)matching.b_to_a.items.size() - 1u;  // wrapping_sub(1), :476-478
        for (auto& kv : matching.b_to_a.items)
            if (!root_index.count(kv.first)) visit(kv.first);
        std::vector<StronglyConnectedExpressions> out;
        for (size_t k = sccs.size(); k-- > 0;) {
            StronglyConnectedExpressions sc;
            sc.expressions = sccs[k];
            std::set<uint32_t> fv;
            for (uint32_t e : sc.expressions) {
                uint32_t matched_var = *matching.b_to_a.get(e);
                for (uint32_t var : graph.expressions[e])
                    if (var == matched_var || (!matching.a_to_b.contains(var) && free_variables.count(var))) fv.insert(var);
            }
            sc.free_variables.assign(fv.begin(), fv.end());
            out.push_back(std::move(sc));
        }
        return out;
    }
};

struct System {
    Graph graph;
    std::vector<EncodedElement> elements;
    std::vector<double> variables, variables_transformed;
    std::vector<uint32_t> variable_to_primitive;
    std::set<uint32_t> fixed_variables;
    std::vector<EncodedConstraint> constraints;
    std::vector<Expression> expressions, expressions_transformed;
    std::vector<LmReport> last_reports;  // one per solved component, in component order

    // lib.rs:363-407
    uint32_t add_element(const double* vars, int n, EncodedElement (*mk)(uint32_t, uint32_t, uint32_t),
                         uint32_t a, uint32_t b) {
        uint32_t id = (uint32_t)elements.size();
        uint32_t variables_idx = (uint32_t)variables.size();
        for (int k = 0; k < n; k++) {
            variables.push_back(vars[k]);
            variable_to_primitive.push_back(id);
        }
        graph.add_element((int16_t)n);  // lib.rs:372-403: dof = number of variables the element adds
        elements.push_back(mk(variables_idx, a, b));
        return id;
    }
    // elements/mod.rs:280-454
    uint32_t add_length(double length) {
        return add_element(&length, 1, [](uint32_t vi, uint32_t, uint32_t) { return EncodedElement{ELength, vi, 0}; }, 0, 0);
    }
    uint32_t add_point(double x, double y) {
        double v[2] = {x, y};
        return add_element(v, 2, [](uint32_t vi, uint32_t, uint32_t) { return EncodedElement{EPoint, vi, 0}; }, 0, 0);
    }
    uint32_t add_line(uint32_t p1, uint32_t p2) {
        return add_element(nullptr, 0, [](uint32_t, uint32_t a, uint32_t b) { return EncodedElement{ELine, a, b}; },
                           elements[p1].a, elements[p2].a);
    }
    uint32_t add_circle(uint32_t center, uint32_t radius) {
        return add_element(nullptr, 0, [](uint32_t, uint32_t a, uint32_t b) { return EncodedElement{ECircle, a, b}; },
                           elements[center].a, elements[radius].a);
    }
    // elements/mod.rs variable_indices per element type
    int element_variables(uint32_t id, uint32_t out[4]) const {
        const EncodedElement& e = elements[id];
        switch (e.tag) {
            case ELength: out[0] = e.a; return 1;
            case EPoint: out[0] = e.a; out[1] = e.a + 1; return 2;
            case ELine: out[0] = e.a; out[1] = e.a + 1; out[2] = e.b; out[3] = e.b + 1; return 4;
            case ECircle: out[0] = e.a; out[1] = e.a + 1; out[2] = e.b; return 3;
        }
        return 0;
    }
    // elements/mod.rs:60-86
    void fix(uint32_t id) {
        uint32_t v[4];
        int n = element_variables(id, v);
        for (int k = 0; k < n; k++) fixed_variables.insert(v[k]);
    }
    void unfix(uint32_t id) {
        uint32_t v[4];
        int n = element_variables(id, v);
        for (int k = 0; k < n; k++) fixed_variables.erase(v[k]);
    }

    // lib.rs:412-445
    uint32_t push_constraint(uint8_t tag, const Expression* exprs, int n) {
        uint32_t id = (uint32_t)constraints.size();
        constraints.push_back({tag, (uint32_t)expressions.size()});
        for (int k = 0; k < n; k++) expressions.push_back(exprs[k]);
        return id;
    }
    uint32_t prim(uint32_t var) const { return variable_to_primitive[var]; }

    // constraints/mod.rs:317-891.  Arguments are element ids of the kinds the reference takes.
    uint32_t point_point_coincidence(uint32_t p1, uint32_t p2) {
        uint32_t i1 = elements[p1].a, i2 = elements[p2].a;
        uint32_t inc[2] = {p1, p2};
        graph.add_constraint(inc, 2, 2);  // constraints/mod.rs:331-334: valency 2
        Expression e[2];
        e[0].kind = VariableVariableEquality; e[0].idx[0] = i1; e[0].idx[1] = i2;
        e[1].kind = VariableVariableEquality; e[1].idx[0] = i1 + 1; e[1].idx[1] = i2 + 1;
        return push_constraint(0, e, 2);
    }
    uint32_t point_point_distance(uint32_t p1, uint32_t p2, double distance) {
        uint32_t inc[2] = {p1, p2};
        graph.add_constraint(inc, 2);
        Expression e;
        e.kind = PointPointDistance; e.idx[0] = elements[p1].a; e.idx[1] = elements[p2].a; e.param = distance;
        return push_constraint(1, &e, 1);
    }
    uint32_t point_point_point_angle(uint32_t p1, uint32_t p2, uint32_t p3, double angle) {
        uint32_t inc[3] = {p1, p2, p3};
        graph.add_constraint(inc, 3);
        Expression e;
        e.kind = PointPointPointAngle;
        e.idx[0] = elements[p1].a; e.idx[1] = elements[p2].a; e.idx[2] = elements[p3].a; e.param = angle;
        return push_constraint(2, &e, 1);
    }
    uint32_t point_line_incidence(uint32_t point, uint32_t line) {
        uint32_t l1 = elements[line].a, l2 = elements[line].b;
        uint32_t inc[3] = {point, prim(l1), prim(l2)};
        graph.add_constraint(inc, 3);
        Expression e;
        e.kind = PointLineIncidence; e.idx[0] = elements[point].a; e.idx[1] = l1; e.idx[2] = l2;
        return push_constraint(3, &e, 1);
    }
    uint32_t point_line_distance(uint32_t point, uint32_t line, double distance) {
        uint32_t l1 = elements[line].a, l2 = elements[line].b;
        uint32_t inc[3] = {point, prim(l1), prim(l2)};
        graph.add_constraint(inc, 3);
        Expression e;
        e.kind = PointLineDistance; e.idx[0] = elements[point].a; e.idx[1] = l1; e.idx[2] = l2; e.param = distance;
        return push_constraint(4, &e, 1);
    }
    uint32_t point_circle_incidence(uint32_t point, uint32_t circle) {
        uint32_t c = elements[circle].a, r = elements[circle].b;
        uint32_t inc[3] = {point, prim(c), prim(r)};
        graph.add_constraint(inc, 3);
        Expression e;
        e.kind = PointCircleIncidence; e.idx[0] = elements[point].a; e.idx[1] = c; e.idx[2] = r;
        return push_constraint(5, &e, 1);
    }
    uint32_t four_point(uint8_t ctag, uint8_t kind, uint32_t a, uint32_t b, uint32_t c, uint32_t d, double param) {
        uint32_t inc[4] = {prim(a), prim(b), prim(c), prim(d)};
        graph.add_constraint(inc, 4);
        Expression e;
        e.kind = kind; e.idx[0] = a; e.idx[1] = b; e.idx[2] = c; e.idx[3] = d; e.param = param;
        return push_constraint(ctag, &e, 1);
    }
    uint32_t segment_segment_length_equality(uint32_t s1p1, uint32_t s1p2, uint32_t s2p1, uint32_t s2p2) {
        return four_point(6, SegmentSegmentLengthEquality, elements[s1p1].a, elements[s1p2].a,
                          elements[s2p1].a, elements[s2p2].a, 0.);
    }
    uint32_t line_line_angle(uint32_t l1, uint32_t l2, double angle) {
        return four_point(7, LineLineAngle, elements[l1].a, elements[l1].b, elements[l2].a, elements[l2].b, angle);
    }
    uint32_t line_line_parallelism(uint32_t l1, uint32_t l2) {
        return four_point(8, LineLineParallelism, elements[l1].a, elements[l1].b, elements[l2].a, elements[l2].b, 0.);
    }
    uint32_t line_line_perpendicularity(uint32_t l1, uint32_t l2) {
        return four_point(9, LineLinePerpendicularity, elements[l1].a, elements[l1].b, elements[l2].a, elements[l2].b, 0.);
    }
    uint32_t line_circle_tangency(uint32_t line, uint32_t circle) {
        return four_point(10, LineCircleTangency, elements[line].a, elements[line].b, elements[circle].a,
                          elements[circle].b, 0.);
    }

    // constraints/mod.rs:88-110 (IdentityVariableMap: every variable reads system.variables)
    double calculate_residual(uint32_t constraint) const {
        const EncodedConstraint& c = constraints[constraint];
        int val = valency_of(c.tag);
        auto one = [&](const Expression& e) {
            uint32_t vi[8];
            double vals[8] = {0, 0, 0, 0, 0, 0, 0, 0}, grad[8];
            int a = variable_indices(e, vi);
            for (int k = 0; k < a; k++) vals[k] = variables[vi[k]];
            return compute_residual_and_gradient(e, vals, grad);
        };
        if (val > 1) {
            double s = 0.0;
            for (int k = 0; k < val; k++) {
                double r = one(expressions[c.expressions_idx + k]);
                s += r * r;
            }
            return std::sqrt(s);
        }
        return one(expressions[c.expressions_idx]);
    }

    // assemble/mod.rs:32-44
    double calculate_system_scale() const {
        double s = 0.0;
        size_t n = 0;
        for (double v : variables) { s += v * v; n++; }
        for (const Expression& e : expressions)
            if (e.kind == PointPointDistance || e.kind == PointLineDistance) { s += e.param * e.param; n++; }
        return std::sqrt(s / (double)n);
    }

    // One component's flattened problem as handed to `Subsystem::new` (assemble/mod.rs:91-146).
    struct ComponentProblem {
        std::vector<uint32_t> free_variables, rows;
    };

    // assemble/mod.rs:46-146: scale, then per component (Vec order, empty ones skipped) the free
    // set, the perturbation and the row list.  `visit(cp)` is called at the point where the
    // reference constructs the `Subsystem`, i.e. `variables_transformed` holds the perturbation of
    // this and all earlier components only (matters for stale elements).
    template <class Visit>
    double for_each_component(const SolvingOptions& opts, Visit visit) {
        Rng rng(42);
        double system_scale = calculate_system_scale();
        double recip = 1. / system_scale;
        variables_transformed.assign(variables.size(), 0.);
        for (size_t i = 0; i < variables.size(); i++) variables_transformed[i] = variables[i] * recip;
        expressions_transformed.clear();
        for (const Expression& e : expressions) expressions_transformed.push_back(transform(e, recip));
        for (size_t ci = 0; ci < graph.components.size(); ci++) {
            const ConnectedComponent& cc = graph.components[ci];
            if (cc.elements.empty()) continue;
            std::set<uint32_t> free_set;
            for (uint32_t el : cc.elements) {
                uint32_t v[4];
                int n = element_variables(el, v);
                for (int k = 0; k < n; k++)
                    if (!fixed_variables.count(v[k])) free_set.insert(v[k]);
            }
            if (opts.perturb) {
                for (uint32_t fv : free_set) {
                    double& variable = variables_transformed[fv];
                    double r1 = rng.next_f64();
                    double r2 = rng.next_f64();
                    variable += variable * (1. / 8196.) * r1 + (1. / 65568.) * r2;
                }
            }
            ComponentProblem cp;
            cp.free_variables.assign(free_set.begin(), free_set.end());
            for (uint32_t c : cc.constraints)
                for (int off = 0; off < valency_of(constraints[c].tag); off++)
                    cp.rows.push_back(constraints[c].expressions_idx + (uint32_t)off);
            visit(cp, system_scale);
        }
        return system_scale;
    }

    // assemble/mod.rs:127-167 with Decomposer::None + Optimizer::LevenbergMarquardt.
    void solve(const SolvingOptions& opts, bool keep_artifacts = false) {
        last_reports.clear();
        for_each_component(opts, [&](const ComponentProblem& cp, double system_scale) {
            std::vector<double> free_values;
            for (uint32_t fv : cp.free_variables) free_values.push_back(variables_transformed[fv]);
            Subsystem sub(variables_transformed.data(), variables_transformed.size(),
                          expressions_transformed.data(), cp.free_variables, cp.rows);
            LmReport rep;
            levenberg_marquardt(sub, free_values.data(), rep, keep_artifacts);
            for (size_t k = 0; k < cp.free_variables.size(); k++)
                variables[cp.free_variables[k]] = system_scale * free_values[k];
            last_reports.push_back(std::move(rep));
        });
    }

    // The equation graph of lib.rs:262,395,434: variables in creation order, one vertex per expression.
    ExpressionGraph equation_graph() const {
        ExpressionGraph g;
        g.insert_variables((int)variables.size());
        for (const Expression& e : expressions) {
            uint32_t vi[8];
            int a = variable_indices(e, vi);
            g.insert_expression(vi, a);
        }
        return g;
    }

    // assemble/mod.rs:169-210 (Decomposer::SinglePass, Optimizer::LevenbergMarquardt): every connected
    // component is split into strongly connected sets of expressions which are solved in sequence, each
    // with its own free variables; later sets see the variables solved by earlier ones as fixed values.
    // `plan_out` (optional) receives the sequence of (free variables, expressions) of all components.
    void solve_single_pass(const SolvingOptions& opts, std::vector<StronglyConnectedExpressions>* plan_out = nullptr,
                           bool dry_run = false) {
        last_reports.clear();
        const ExpressionGraph eg = equation_graph();
        for_each_component(opts, [&](const ComponentProblem& cp, double system_scale) {
            std::set<uint32_t> free_set(cp.free_variables.begin(), cp.free_variables.end());
            SinglePassPlanner planner(eg, free_set);
            for (const StronglyConnectedExpressions& scc : planner.plan()) {
                if (plan_out) plan_out->push_back(scc);
                if (dry_run) continue;
                std::vector<double> free_values;
                for (uint32_t fv : scc.free_variables) free_values.push_back(variables_transformed[fv]);
                Subsystem sub(variables_transformed.data(), variables_transformed.size(), expressions_transformed.data(),
                              scc.free_variables, scc.expressions);
                LmReport rep;
                levenberg_marquardt(sub, free_values.data(), rep, false);
                for (size_t k = 0; k < scc.free_variables.size(); k++) {
                    variables_transformed[scc.free_variables[k]] = free_values[k];
                    variables[scc.free_variables[k]] = system_scale * free_values[k];
                }
                last_reports.push_back(std::move(rep));
            }
        });
    }

    // lib.rs:454-458 + analyze/numerical/mod.rs:149-160: constraints owning a dependent expression, in
    // expression order (a two-expression constraint can be listed twice, as in the reference).
    std::vector<uint32_t> analyze() const {
        const std::vector<uint8_t> independent = analyze_expressions(variables, expressions);
        std::vector<uint32_t> expression_to_constraint(expressions.size(), 0);
        for (uint32_t c = 0; c < constraints.size(); c++)
            for (int off = 0; off < valency_of(constraints[c].tag); off++)
                expression_to_constraint[constraints[c].expressions_idx + off] = c;
        std::vector<uint32_t> dependent;
        for (size_t e = 0; e < independent.size(); e++)
            if (!independent[e]) dependent.push_back(expression_to_constraint[e]);
        return dependent;
    }
};


// ---- Decomposer::RecursiveAssembly: the per-step problem and the driver (assemble/mod.rs:212-725) ------
// constraints/expressions.rs:1094-1159
struct Pose2D {
    double rotation, tx, ty;
    void transform_point(double u, double v, double& x, double& y) const {  // :1122-1136
        const double s = std::sin(rotation), c = std::cos(rotation);  // f64::sin_cos
        const double uc = u * c, us = u * s, vc = v * c, vs = v * s;
        x = tx + uc - vs;
        y = ty + us + vc;
    }
    void gradient_chain_rule_point(double u, double v, double g0, double g1, double out[3]) const {  // :1139-1158
        const double s = std::sin(rotation), c = std::cos(rotation);
        const double uc = u * c, us = u * s, vc = v * c, vs = v * s;
        out[0] = (-us - vc) * g0 + (uc - vs) * g1;
        out[1] = g0;
        out[2] = g1;
    }
};

// assemble/mod.rs:282-590: cluster poses (3 variables each, initially 0) followed by the variables of the step's elements;
// rows: two coincidence rows per (cluster, frontier point), then the step's expressions.
struct ClusteredSystem {
    const System* system = nullptr;
    uint32_t num_variables_ = 0, num_pose_expressions = 0;
    std::vector<uint32_t> step_plus_frontier_elements, expressions;
    std::vector<std::pair<uint32_t, std::vector<uint32_t>>> clusters;  // IndexMap<ClusterKey, Vec<ElementId>>: insertion order
    std::map<uint32_t, uint32_t> variable_mapping_pose_and_element;

    static const std::vector<uint32_t>* lookup(const std::map<uint32_t, std::vector<uint32_t>>& m, uint32_t key) {
        auto it = m.find(key);
        return it == m.end() ? nullptr : &it->second;
    }
    // :322-478
    void build(const System& sys, const RecombinationStep& step, std::vector<double>& pose_and_element_variables) {
        system = &sys;
        num_variables_ = 0; num_pose_expressions = 0;
        step_plus_frontier_elements.clear(); expressions.clear(); clusters.clear(); variable_mapping_pose_and_element.clear();
        pose_and_element_variables.clear();
        for (uint32_t c : step.constraints)
            for (int off = 0; off < valency_of(sys.constraints[c].tag); off++) expressions.push_back(sys.constraints[c].expressions_idx + (uint32_t)off);
        step_plus_frontier_elements = step.elements;
        auto contains = [](const std::vector<uint32_t>& v, uint32_t x) { return std::find(v.begin(), v.end(), x) != v.end(); };
        std::vector<uint32_t> reachable_clusters;  // :355-398
        for (uint32_t el : step.elements) {
            if (sys.elements[el].tag != EPoint) continue;
            if (const std::vector<uint32_t>* cl = lookup(step.on_frontiers, el))
                for (uint32_t c : *cl)
                    if (!contains(reachable_clusters, c)) reachable_clusters.push_back(c);
        }
        for (size_t i = 0; i < reachable_clusters.size(); i++) {
            const uint32_t cluster = reachable_clusters[i];
            for (uint32_t el : step.frontier_elements.at(cluster)) {
                if (sys.elements[el].tag != EPoint) continue;
                const std::vector<uint32_t>* cl = lookup(step.on_frontiers, el);
                if (cl)
                    for (uint32_t c : *cl)
                        if (!contains(reachable_clusters, c)) reachable_clusters.push_back(c);
                const size_t num_frontiers = cl ? cl->size() : 0;
                if (!contains(step_plus_frontier_elements, el) && num_frontiers > 1) step_plus_frontier_elements.push_back(el);
            }
        }
        for (uint32_t el : step_plus_frontier_elements) {  // :401-429
            const std::vector<uint32_t>* cl = lookup(step.on_frontiers, el);
            if (!cl || sys.elements[el].tag != EPoint) continue;
            for (uint32_t cluster : *cl) {
                num_pose_expressions += 2;
                auto it = std::find_if(clusters.begin(), clusters.end(), [&](const std::pair<uint32_t, std::vector<uint32_t>>& c) { return c.first == cluster; });
                if (it == clusters.end()) {
                    clusters.emplace_back(cluster, std::vector<uint32_t>());
                    it = clusters.end() - 1;
                }
                it->second.push_back(el);
            }
        }
        pose_and_element_variables.assign(clusters.size() * 3, 0.);  // :432
        for (uint32_t el : step_plus_frontier_elements) {  // :437-475
            const EncodedElement& e = sys.elements[el];
            if (e.tag == ELength) {
                variable_mapping_pose_and_element[e.a] = (uint32_t)pose_and_element_variables.size();
                pose_and_element_variables.push_back(sys.variables_transformed[e.a]);
            } else if (e.tag == EPoint) {
                variable_mapping_pose_and_element[e.a] = (uint32_t)pose_and_element_variables.size();
                variable_mapping_pose_and_element[e.a + 1] = (uint32_t)pose_and_element_variables.size() + 1;
                pose_and_element_variables.push_back(sys.variables_transformed[e.a]);
                pose_and_element_variables.push_back(sys.variables_transformed[e.a + 1]);
            }
        }
        num_variables_ = (uint32_t)pose_and_element_variables.size();
    }
    uint32_t num_variables() const { return num_variables_; }                                          // :620-623
    uint32_t num_residuals() const { return (uint32_t)expressions.size() + num_pose_expressions; }     // :625-628

    // :630-685
    void calculate_residuals(const double* x, double* residuals) const {
        for (uint32_t r = 0; r < num_residuals(); r++) residuals[r] = 0.;
        uint32_t vi[8];
        double vals[8] = {0, 0, 0, 0, 0, 0, 0, 0}, grad[8];
        const size_t offset = num_pose_expressions;
        for (size_t k = 0; k < expressions.size(); k++) {
            const Expression& e = system->expressions_transformed[expressions[k]];
            const int a = variable_indices(e, vi);
            for (int q = 0; q < a; q++) vals[q] = x[variable_mapping_pose_and_element.at(vi[q])];
            residuals[offset + k] = compute_residual_and_gradient(e, vals, grad);
        }
        size_t r = 0;
        for (size_t ci = 0; ci < clusters.size(); ci++) {
            const Pose2D pose{x[3 * ci], x[3 * ci + 1], x[3 * ci + 2]};
            for (uint32_t point : clusters[ci].second) {
                const uint32_t idx = system->elements[point].a;
                double tx_, ty_;
                pose.transform_point(system->variables_transformed[idx], system->variables_transformed[idx + 1], tx_, ty_);
                const uint32_t updated_idx = variable_mapping_pose_and_element.at(idx);
                residuals[r] = tx_ - x[updated_idx];
                residuals[r + 1] = ty_ - x[updated_idx + 1];
                r += 2;
            }
        }
    }
    // :486-588 through PushTriplet for TripletMat (:598-602)
    void calculate_residuals_and_sparse_jacobian(const double* x, double* residuals, solvi::TripletMat& jac) const {
        uint32_t vi[8];
        double vals[8] = {0, 0, 0, 0, 0, 0, 0, 0}, grad[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const size_t offset = num_pose_expressions;
        for (size_t k = 0; k < expressions.size(); k++) {
            const Expression& e = system->expressions_transformed[expressions[k]];
            const int a = variable_indices(e, vi);
            for (int q = 0; q < a; q++) vals[q] = x[variable_mapping_pose_and_element.at(vi[q])];
            residuals[offset + k] = compute_residual_and_gradient(e, vals, grad);
            for (int q = 0; q < a; q++) jac.push_triplet(offset + k, variable_mapping_pose_and_element.at(vi[q]), grad[q]);
        }
        size_t r = 0;
        for (size_t ci = 0; ci < clusters.size(); ci++) {
            const size_t pose_start_idx = 3 * ci;
            const Pose2D pose{x[pose_start_idx], x[pose_start_idx + 1], x[pose_start_idx + 2]};
            for (uint32_t point : clusters[ci].second) {
                const uint32_t idx = system->elements[point].a;
                const double u = system->variables_transformed[idx], v = system->variables_transformed[idx + 1];
                double tx_, ty_;
                pose.transform_point(u, v, tx_, ty_);
                const uint32_t updated_idx = variable_mapping_pose_and_element.at(idx);
                residuals[r] = tx_ - x[updated_idx];
                residuals[r + 1] = ty_ - x[updated_idx + 1];
                double gx[3], gy[3];
                pose.gradient_chain_rule_point(u, v, 1., 0., gx);
                pose.gradient_chain_rule_point(u, v, 0., 1., gy);
                for (int q = 0; q < 3; q++) jac.push_triplet(r, pose_start_idx + q, gx[q]);
                for (int q = 0; q < 3; q++) jac.push_triplet(r + 1, pose_start_idx + q, gy[q]);
                jac.push_triplet(r, updated_idx, -1.);
                jac.push_triplet(r + 1, updated_idx + 1, -1.);
                r += 2;
            }
        }
    }
};

// assemble/mod.rs:212-277.  `plan_out` (optional) receives the steps of every component in order.
inline void solve_recursive_assembly(System& sys, const SolvingOptions& opts, std::vector<RecombinationStep>* plan_out = nullptr,
                                     bool dry_run = false) {
    sys.last_reports.clear();
    // (for_each_component hands over the component's rows and free variables; the decomposition starts again from its
    // element and constraint sets, so the loop of assemble/mod.rs:81-125 is walked here with the same generator)
    size_t comp_at = 0;
    sys.for_each_component(opts, [&](const System::ComponentProblem&, double system_scale) {
        while (sys.graph.components[comp_at].elements.empty()) comp_at++;
        const ConnectedComponent& cc = sys.graph.components[comp_at++];
        const std::vector<RecombinationStep> steps = decompose<3>(sys.graph, cc.elements, cc.constraints);
        ClusteredSystem clustered;
        std::vector<double> x;
        for (const RecombinationStep& step : steps) {
            if (plan_out) plan_out->push_back(step);
            if (dry_run) continue;
            clustered.build(sys, step, x);
            LmReport rep;
            levenberg_marquardt(clustered, x.data(), rep, false);
            for (const auto& kv : clustered.variable_mapping_pose_and_element) {  // :226-233
                sys.variables_transformed[kv.first] = x[kv.second];
                sys.variables[kv.first] = system_scale * x[kv.second];
            }
            for (size_t ci = 0; ci < clustered.clusters.size(); ci++) {  // :235-273
                const Pose2D pose{x[3 * ci], x[3 * ci + 1], x[3 * ci + 2]};
                const std::vector<uint32_t>* owned = ClusteredSystem::lookup(step.owned_elements, clustered.clusters[ci].first);
                if (!owned) continue;
                for (uint32_t el : *owned) {
                    if (std::find(clustered.step_plus_frontier_elements.begin(), clustered.step_plus_frontier_elements.end(), el) !=
                        clustered.step_plus_frontier_elements.end())
                        continue;
                    if (sys.elements[el].tag != EPoint) continue;
                    const uint32_t idx = sys.elements[el].a;
                    double nx, ny;
                    pose.transform_point(sys.variables_transformed[idx], sys.variables_transformed[idx + 1], nx, ny);
                    sys.variables_transformed[idx] = nx;
                    sys.variables_transformed[idx + 1] = ny;
                    sys.variables[idx] = system_scale * nx;
                    sys.variables[idx + 1] = system_scale * ny;
                }
            }
            sys.last_reports.push_back(std::move(rep));
        }
    });
}

}  // namespace fiksi
}  // namespace orc
