"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/liboracle.so``, the CPU restatement of the reference's numeric solve
path (colamd_rs + solvi + fiksi, see the headers in this directory for file:line citations).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; nothing under ``fiksi_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only, a few seconds)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class FkProblem(C.Structure):
    """Layout of ``fk_problem`` in include/fiksi_b200.h."""

    _fields_ = [
        ("n_vars", C.c_uint32), ("vars", C.POINTER(C.c_double)),
        ("n_expr", C.c_uint32), ("kind", C.POINTER(C.c_uint8)),
        ("idx", C.POINTER(C.c_uint32)), ("param", C.POINTER(C.c_double)),
        ("n_free", C.c_uint32), ("free_vars", C.POINTER(C.c_uint32)),
        ("n_rows", C.c_uint32), ("rows", C.POINTER(C.c_uint32)),
    ]


class FkReport(C.Structure):
    """Layout of ``fk_report`` in include/fiksi_b200.h."""

    _fields_ = [
        ("exit_reason", C.c_uint32), ("outer_iters", C.c_uint32),
        ("factorizations", C.c_uint32), ("accepted", C.c_uint32),
        ("ssr", C.c_double), ("lambda_", C.c_double), ("trace_hash", C.c_uint64),
    ]


REPORT_DTYPE = np.dtype([
    ("exit_reason", "<u4"), ("outer_iters", "<u4"), ("factorizations", "<u4"), ("accepted", "<u4"),
    ("ssr", "<f8"), ("lambda", "<f8"), ("trace_hash", "<u8"),
])
assert REPORT_DTYPE.itemsize == C.sizeof(FkReport) == 40

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_expr_eval.restype = C.c_double
        L.orc_calculate_residual.restype = C.c_double
        L.orc_system_scale.restype = C.c_double
        L.orc_prepared_problem.restype = C.c_double
        L.orc_lm_solve_batch_uniform.restype = C.c_double
        L.orc_lbfgs_solve_batch_uniform.restype = C.c_double
        L.orc_system_new.restype = C.c_void_p
        L.orc_prepare.restype = C.c_void_p
        L.orc_sym_build.restype = C.c_void_p
        L.orc_sym_len.restype = C.c_uint64
        L.orc_node_depth_levels.restype = C.c_uint64
        for name in ("orc_add_point", "orc_add_length", "orc_add_line", "orc_add_circle",
                     "orc_point_point_coincidence", "orc_point_point_distance",
                     "orc_point_point_point_angle", "orc_point_line_incidence",
                     "orc_point_line_distance", "orc_point_circle_incidence",
                     "orc_segment_segment_length_equality", "orc_line_line_angle",
                     "orc_line_line_parallelism", "orc_line_line_perpendicularity",
                     "orc_line_circle_tangency", "orc_num_variables", "orc_num_expressions",
                     "orc_num_constraints", "orc_num_reports", "orc_num_components",
                     "orc_component_sizes", "orc_prepared_count", "orc_element_variable", "orc_system_analyze",
                     "orc_single_pass_problem", "orc_recursive_assembly_plan"):
            getattr(L, name).restype = C.c_uint32
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def make_problem(vars_, kind, idx, param, free_vars, rows):
    """Build an FkProblem over numpy arrays; returns (problem, keepalive)."""
    vars_ = np.ascontiguousarray(vars_, dtype=np.float64)
    kind = np.ascontiguousarray(kind, dtype=np.uint8)
    idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 4)
    param = np.ascontiguousarray(param, dtype=np.float64)
    free_vars = np.ascontiguousarray(free_vars, dtype=np.uint32)
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    p = FkProblem(len(vars_), _p(vars_, C.c_double), len(kind), _p(kind, C.c_uint8),
                  _p(idx, C.c_uint32), _p(param, C.c_double), len(free_vars),
                  _p(free_vars, C.c_uint32), len(rows), _p(rows, C.c_uint32))
    return p, (vars_, kind, idx, param, free_vars, rows)


# ---- colamd ---------------------------------------------------------------------------------
def colamd_recommended(nnz, n_row, n_col):
    out = C.c_uint64(0)
    ok = lib().orc_colamd_recommended(nnz, n_row, n_col, C.byref(out))
    return out.value if ok else None


def colamd(n_row, n_col, row_indices, col_ptr, a_len=None, dense_row=10.0, dense_col=10.0,
           aggressive=True):
    """Returns (ok, perm_with_trailing_entry, stats)."""
    nnz = len(row_indices)
    if a_len is None:
        a_len = colamd_recommended(nnz, n_row, n_col)
    a = np.zeros(max(a_len, 1), dtype=np.int32)
    a[:min(nnz, a_len)] = np.asarray(row_indices, dtype=np.int32)[:min(nnz, a_len)]
    p = np.array(col_ptr, dtype=np.int32)
    stats = np.zeros(20, dtype=np.int32)
    ok = lib().orc_colamd(n_row, n_col, C.c_uint64(a_len), _p(a, C.c_int), _p(p, C.c_int),
                          C.c_double(dense_row), C.c_double(dense_col), int(aggressive),
                          _p(stats, C.c_int))
    return bool(ok), p, stats


def symamd(n, row_indices, col_ptr):
    a = np.asarray(row_indices, dtype=np.int32)
    if len(a) == 0:
        a = np.zeros(1, dtype=np.int32)
    p = np.asarray(col_ptr, dtype=np.int32)
    perm = np.zeros(n + 1, dtype=np.int32)
    stats = np.zeros(20, dtype=np.int32)
    ok = lib().orc_symamd(n, _p(a, C.c_int), _p(p, C.c_int), _p(perm, C.c_int), C.c_double(10.0),
                          C.c_double(10.0), 1, _p(stats, C.c_int))
    return bool(ok), perm, stats


# ---- solvi ----------------------------------------------------------------------------------
def from_triplets(m, n, rows, cols, vals):
    rows = np.asarray(rows, dtype=np.uint64)
    cols = np.asarray(cols, dtype=np.uint64)
    vals = np.asarray(vals, dtype=np.float64)
    nnz = len(vals)
    n_out = max(n, int(cols.max()) + 1 if nnz else 0)
    shape = np.zeros(3, dtype=np.uint64)
    colptr = np.zeros(n_out + 1, dtype=np.uint64)
    rowidx = np.zeros(max(nnz, 1), dtype=np.uint64)
    out = np.zeros(max(nnz, 1), dtype=np.float64)
    lib().orc_from_triplets(C.c_uint64(m), C.c_uint64(n), C.c_uint64(nnz), _p(rows, C.c_uint64),
                            _p(cols, C.c_uint64), _p(vals, C.c_double), _p(shape, C.c_uint64),
                            _p(colptr, C.c_uint64), _p(rowidx, C.c_uint64), _p(out, C.c_double))
    k = int(shape[2])
    return (int(shape[0]), int(shape[1])), colptr[:int(shape[1]) + 1].copy(), rowidx[:k].copy(), out[:k].copy()


def upper_solve(colptr, rowidx, vals, b):
    colptr = np.asarray(colptr, dtype=np.uint64)
    rowidx = np.asarray(rowidx, dtype=np.uint64)
    vals = np.asarray(vals, dtype=np.float64)
    b = np.array(b, dtype=np.float64)
    ok = lib().orc_upper_solve(C.c_uint64(len(colptr) - 1), _p(colptr, C.c_uint64),
                               _p(rowidx, C.c_uint64), _p(vals, C.c_double), _p(b, C.c_double))
    return bool(ok), b


NONE = np.uint64(2**64 - 1)


def post_order(parents):
    p = np.array([NONE if x is None or x < 0 else x for x in parents], dtype=np.uint64)
    out = np.zeros(len(p), dtype=np.uint64)
    lib().orc_post_order(C.c_uint64(len(p)), _p(p, C.c_uint64), _p(out, C.c_uint64))
    return out.astype(np.int64)


def node_depth_levels(parents):
    p = np.array([NONE if x is None or x < 0 else x for x in parents], dtype=np.uint64)
    out = np.zeros(len(p), dtype=np.uint64)
    mx = lib().orc_node_depth_levels(C.c_uint64(len(p)), _p(p, C.c_uint64), _p(out, C.c_uint64))
    return out.astype(np.int64), int(mx)


def permute(perm, data, how="gather"):
    perm = np.asarray(perm, dtype=np.uint64)
    d = np.array(data, dtype=np.float64)
    fn = lib().orc_permute_by_gather if how == "gather" else lib().orc_permute_by_swaps
    fn(C.c_uint64(len(perm)), _p(perm, C.c_uint64), _p(d, C.c_double))
    return d


class Symbolic:
    """SymbolicQr::build + Qr (solvi/src/decomposition/sparse/qr.rs)."""

    NAMES = {"parents": 0, "postorder": 1, "row_counts": 2, "col_counts": 3, "r_colptr": 4,
             "r_rowidx": 5, "h_colptr": 6, "h_rowidx": 7, "row_permutation": 8,
             "col_permutation": 9, "levels": 10, "first_columns": 11}

    def __init__(self, m, n, colptr, rowidx, ordering="natural"):
        self.m, self.n = m, n
        self._colptr = np.asarray(colptr, dtype=np.uint64)
        self._rowidx = np.asarray(rowidx, dtype=np.uint64)
        self.h = lib().orc_sym_build(C.c_uint64(m), C.c_uint64(n), _p(self._colptr, C.c_uint64),
                                     _p(self._rowidx, C.c_uint64), 1 if ordering == "colamd" else 0)
        if not self.h:
            raise RuntimeError("symbolic build failed")

    def get(self, name):
        which = self.NAMES[name]
        n = lib().orc_sym_len(C.c_void_p(self.h), which)
        out = np.zeros(max(int(n), 1), dtype=np.uint64)
        lib().orc_sym_get(C.c_void_p(self.h), which, _p(out, C.c_uint64))
        out = out[:int(n)].astype(np.int64)
        out[out == -1] = -1  # usize::MAX wraps to -1
        return out

    def factorize(self, values):
        v = np.asarray(values, dtype=np.float64)
        lib().orc_qr_factorize(C.c_void_p(self.h), _p(v, C.c_double))

    def r_values(self):
        out = np.zeros(len(self.get("r_rowidx")), dtype=np.float64)
        lib().orc_qr_r_values(C.c_void_p(self.h), _p(out, C.c_double))
        return out

    def solve(self, b):
        b = np.array(b, dtype=np.float64)
        ok = lib().orc_qr_solve(C.c_void_p(self.h), _p(b, C.c_double))
        return bool(ok), b

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sym_free(C.c_void_p(self.h))
            self.h = None


# ---- fiksi ----------------------------------------------------------------------------------
def rng_u32(seed, n):
    out = np.zeros(n, dtype=np.uint32)
    lib().orc_rng_u32(seed, n, _p(out, C.c_uint32))
    return out


def rng_f64(seed, n):
    out = np.zeros(n, dtype=np.float64)
    lib().orc_rng_f64(seed, n, _p(out, C.c_double))
    return out


def expr_eval(kind, param, vars8):
    v = np.zeros(8, dtype=np.float64)
    v[:len(vars8)] = vars8
    g = np.zeros(8, dtype=np.float64)
    r = lib().orc_expr_eval(int(kind), C.c_double(param), _p(v, C.c_double), _p(g, C.c_double))
    return r, g


def expr_slots(kind):
    return lib().orc_expr_slots(int(kind))


def lbfgs_solve(problem, free_values):
    """lbfgs(problem, variables) (fiksi/src/solve/lbfgs.rs:20): returns (x, report dict)."""
    x = np.array(free_values, dtype=np.float64)
    rep = FkReport()
    lib().orc_lbfgs_solve(C.byref(problem), _p(x, C.c_double), C.byref(rep))
    return x, report_dict(rep)


def lbfgs_solve_batch_uniform(topo_problem, vars_, param, threads=1):
    vars_ = np.ascontiguousarray(vars_, dtype=np.float64)
    param = np.ascontiguousarray(param, dtype=np.float64)
    n = vars_.shape[0]
    out = np.zeros((n, topo_problem.n_free), dtype=np.float64)
    reports = np.zeros(n, dtype=REPORT_DTYPE)
    secs = lib().orc_lbfgs_solve_batch_uniform(C.byref(topo_problem), n, _p(vars_, C.c_double), _p(param, C.c_double),
                                               _p(out, C.c_double), reports.ctypes.data_as(C.POINTER(FkReport)), threads)
    return out, reports, secs


def single_pass_problem(problem, vars_, cap=4096):
    """assemble/mod.rs:169-210 on one flattened component: returns (vars after the pass, reports)."""
    v = np.array(vars_, dtype=np.float64)
    reports = np.zeros(cap, dtype=REPORT_DTYPE)
    n = lib().orc_single_pass_problem(C.byref(problem), _p(v, C.c_double), reports.ctypes.data_as(C.POINTER(FkReport)), C.c_uint32(cap))
    return v, reports[:n]


def analyze(vars_, kind, idx, param):
    """find_overconstraints' per-expression "independent" flags (analyze/numerical/mod.rs:123-147)."""
    vars_ = np.ascontiguousarray(vars_, dtype=np.float64)
    kind = np.ascontiguousarray(kind, dtype=np.uint8)
    idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 4)
    param = np.ascontiguousarray(param, dtype=np.float64)
    out = np.zeros(max(len(kind), 1), dtype=np.uint8)
    lib().orc_analyze(C.c_uint32(len(vars_)), _p(vars_, C.c_double), C.c_uint32(len(kind)), _p(kind, C.c_uint8),
                      _p(idx, C.c_uint32), _p(param, C.c_double), _p(out, C.c_uint8))
    return out[:len(kind)].astype(bool)


def gauss_jordan(matrix, column_indices=None):
    """incremental_gauss_jordan_elimination (analyze/numerical/mod.rs:33-117): returns (reduced matrix,
    column order, increases-rank flags)."""
    m = np.array(matrix, dtype=np.float64, order="C")
    nr, nc = m.shape
    ci = np.arange(nc, dtype=np.uint64) if column_indices is None else np.array(column_indices, dtype=np.uint64)
    out = np.zeros(max(nr, 1), dtype=np.uint8)
    lib().orc_gauss_jordan(_p(m, C.c_double), C.c_uint64(nr), C.c_uint64(nc), _p(ci, C.c_uint64), _p(out, C.c_uint8))
    return m, ci, out[:nr].astype(bool)


class qr_fast:
    """Context manager: run the oracle's sparse QR without the reference's per-column O(m) scratch fill
    (solvi qr.rs:287).  Results are bit-identical (tests/test_oracle_fast_mode.py); the slow mode stays
    the default because it is the reference's cost, which ``cpu_baseline`` times."""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        self.prev = lib().orc_get_qr_fast()
        lib().orc_set_qr_fast(1 if self.on else 0)
        return self

    def __exit__(self, *exc):
        lib().orc_set_qr_fast(self.prev)
        return False


def lm_solve(problem, free_values):
    """levenberg_marquardt on a flattened problem.  Returns (x, report dict, trace string)."""
    x = np.array(free_values, dtype=np.float64)
    rep = FkReport()
    trace = C.create_string_buffer(4096)
    lib().orc_lm_solve(C.byref(problem), _p(x, C.c_double), C.byref(rep), trace, 4096)
    return x, report_dict(rep), trace.value.decode()


def report_dict(rep):
    return {"exit_reason": rep.exit_reason, "outer_iters": rep.outer_iters,
            "factorizations": rep.factorizations, "accepted": rep.accepted, "ssr": rep.ssr,
            "lambda": rep.lambda_, "trace_hash": rep.trace_hash}


def symbolic(problem):
    n = problem.n_free
    sizes = np.zeros(2, dtype=np.uint32)
    lib().orc_symbolic(C.byref(problem), None, None, None, None, None, None, _p(sizes, C.c_uint32))
    colptr = np.zeros(n + 1, dtype=np.uint32)
    rowidx = np.zeros(max(int(sizes[0]), 1), dtype=np.uint32)
    perm = np.zeros(max(n, 1), dtype=np.int32)
    parent = np.zeros(max(n, 1), dtype=np.int32)
    rcolptr = np.zeros(n + 1, dtype=np.uint32)
    rrowidx = np.zeros(max(int(sizes[1]), 1), dtype=np.uint32)
    lib().orc_symbolic(C.byref(problem), _p(colptr, C.c_uint32), _p(rowidx, C.c_uint32),
                       _p(perm, C.c_int32), _p(parent, C.c_int32), _p(rcolptr, C.c_uint32),
                       _p(rrowidx, C.c_uint32), _p(sizes, C.c_uint32))
    return {"aug_colptr": colptr, "aug_rowidx": rowidx[:int(sizes[0])], "perm": perm[:n],
            "etree_parent": parent[:n], "r_colptr": rcolptr, "r_rowidx": rrowidx[:int(sizes[1])]}


def evaluate(problem, free_values, jac_nnz=None):
    x = np.ascontiguousarray(free_values, dtype=np.float64)
    r = np.zeros(max(problem.n_rows, 1), dtype=np.float64)
    if jac_nnz is None:
        lib().orc_residuals(C.byref(problem), _p(x, C.c_double), _p(r, C.c_double))
        return r[:problem.n_rows]
    j = np.zeros(max(jac_nnz, 1), dtype=np.float64)
    lib().orc_eval(C.byref(problem), _p(x, C.c_double), _p(r, C.c_double), _p(j, C.c_double))
    return r[:problem.n_rows], j[:jac_nnz]


def lm_solve_batch_uniform(topo_problem, vars_, param, threads=1):
    vars_ = np.ascontiguousarray(vars_, dtype=np.float64)
    param = np.ascontiguousarray(param, dtype=np.float64)
    n = vars_.shape[0]
    free_out = np.zeros((n, topo_problem.n_free), dtype=np.float64)
    reports = np.zeros(n, dtype=REPORT_DTYPE)
    secs = lib().orc_lm_solve_batch_uniform(C.byref(topo_problem), n, _p(vars_, C.c_double),
                                            _p(param, C.c_double), _p(free_out, C.c_double),
                                            reports.ctypes.data_as(C.POINTER(FkReport)), threads)
    return free_out, reports, secs


class System:
    """Mirror of fiksi::System (fiksi/src/lib.rs:252-467) over the oracle."""

    def __init__(self):
        self.h = lib().orc_system_new()

    def _c(self, name, *args):
        return getattr(lib(), name)(C.c_void_p(self.h), *args)

    def add_point(self, x, y): return self._c("orc_add_point", C.c_double(x), C.c_double(y))
    def add_length(self, l): return self._c("orc_add_length", C.c_double(l))
    def add_line(self, p1, p2): return self._c("orc_add_line", p1, p2)
    def add_circle(self, c, r): return self._c("orc_add_circle", c, r)
    def fix(self, e): self._c("orc_fix", e)
    def unfix(self, e): self._c("orc_unfix", e)
    def point_point_coincidence(self, a, b): return self._c("orc_point_point_coincidence", a, b)
    def point_point_distance(self, a, b, d): return self._c("orc_point_point_distance", a, b, C.c_double(d))
    def point_point_point_angle(self, a, b, c, ang): return self._c("orc_point_point_point_angle", a, b, c, C.c_double(ang))
    def point_line_incidence(self, p, l): return self._c("orc_point_line_incidence", p, l)
    def point_line_distance(self, p, l, d): return self._c("orc_point_line_distance", p, l, C.c_double(d))
    def point_circle_incidence(self, p, c): return self._c("orc_point_circle_incidence", p, c)
    def segment_segment_length_equality(self, a, b, c, d): return self._c("orc_segment_segment_length_equality", a, b, c, d)
    def line_line_angle(self, a, b, ang): return self._c("orc_line_line_angle", a, b, C.c_double(ang))
    def line_line_parallelism(self, a, b): return self._c("orc_line_line_parallelism", a, b)
    def line_line_perpendicularity(self, a, b): return self._c("orc_line_line_perpendicularity", a, b)
    def line_circle_tangency(self, l, c): return self._c("orc_line_circle_tangency", l, c)

    @property
    def variables(self):
        n = self._c("orc_num_variables")
        out = np.zeros(max(n, 1), dtype=np.float64)
        self._c("orc_get_variables", _p(out, C.c_double))
        return out[:n]

    def set_variable(self, i, v): self._c("orc_set_variable", i, C.c_double(v))
    def element_variable(self, e): return self._c("orc_element_variable", e)
    def set_parameter(self, c, v): self._c("orc_set_parameter", c, C.c_double(v))
    def calculate_residual(self, c): return self._c("orc_calculate_residual", c)
    def num_constraints(self): return self._c("orc_num_constraints")
    def system_scale(self): return self._c("orc_system_scale")

    def solve_single_pass(self, perturb=True):
        """System::solve with Decomposer::SinglePass (assemble/mod.rs:169-210), LM optimizer."""
        self._c("orc_solve_single_pass", int(perturb))

    def single_pass_plan(self):
        """[(free variables, expressions), ...]: the sequence of sub-problems SinglePass solves."""
        sizes = np.zeros(3, dtype=np.uint32)
        self._c("orc_single_pass_plan", _p(sizes, C.c_uint32), None, None, None, None)
        n, nf, ne = (int(x) for x in sizes)
        fp = np.zeros(n + 1, np.uint32); fv = np.zeros(max(nf, 1), np.uint32)
        ep = np.zeros(n + 1, np.uint32); ex = np.zeros(max(ne, 1), np.uint32)
        self._c("orc_single_pass_plan", _p(sizes, C.c_uint32), _p(fp, C.c_uint32), _p(fv, C.c_uint32), _p(ep, C.c_uint32), _p(ex, C.c_uint32))
        return [(fv[fp[k]:fp[k + 1]].tolist(), ex[ep[k]:ep[k + 1]].tolist()) for k in range(n)]

    def solve_recursive_assembly(self, perturb=True):
        """System::solve with Decomposer::RecursiveAssembly (assemble/mod.rs:212-277), LM optimizer; hash-set iteration
        orders replaced by ascending ids (see fiksi_ref.hpp)."""
        if self._c("orc_solve_recursive_assembly", int(perturb)):
            raise RuntimeError("the reference panics on this system (unwrap on a missing cluster entry)")

    def recursive_assembly_plan(self):
        """(number of steps, serialised plan words) of the recombination plan (format: oracle_capi.cpp)."""
        steps = C.c_uint32(0)
        n = self._c("orc_recursive_assembly_plan", None, C.c_uint32(0), C.byref(steps))
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self._c("orc_recursive_assembly_plan", _p(out, C.c_uint32), C.c_uint32(n), C.byref(steps))
        if steps.value == 0xFFFFFFFF:
            raise RuntimeError("the reference panics on this system (unwrap on a missing cluster entry)")
        return int(steps.value), out[:n].tolist()

    def analyze(self):
        """System::analyze (lib.rs:454-458): ids of the constraints flagged as over-constraining."""
        out = np.zeros(max(self.num_constraints() * 2, 1), dtype=np.uint32)
        n = self._c("orc_system_analyze", _p(out, C.c_uint32), C.c_uint32(len(out)))
        return out[:n].tolist()

    def point(self, e):
        i = self.element_variable(e)
        v = self.variables
        return float(v[i]), float(v[i + 1])

    def solve(self, perturb=True):
        self._c("orc_solve", int(perturb), 0)

    def reports(self):
        out = []
        for i in range(self._c("orc_num_reports")):
            rep = FkReport()
            trace = C.create_string_buffer(4096)
            self._c("orc_get_report", i, C.byref(rep), trace, 4096)
            d = report_dict(rep)
            d["trace"] = trace.value.decode()
            out.append(d)
        return out

    def components(self):
        out = []
        for ci in range(self._c("orc_num_components")):
            nc = C.c_uint32(0)
            ne = self._c("orc_component_sizes", ci, C.byref(nc))
            el = np.zeros(max(ne, 1), dtype=np.uint32)
            co = np.zeros(max(nc.value, 1), dtype=np.uint32)
            self._c("orc_component_get", ci, _p(el, C.c_uint32), _p(co, C.c_uint32))
            out.append((el[:ne].tolist(), co[:nc.value].tolist()))
        return out

    def prepare(self, perturb=True):
        """Flattened problems as handed to Subsystem::new: list of (FkProblem, scale, keepalive)."""
        ps = self._c("orc_prepare", int(perturb))
        n = lib().orc_prepared_count(C.c_void_p(ps))
        out = []
        for i in range(n):
            p = FkProblem()
            scale = lib().orc_prepared_problem(C.c_void_p(ps), i, C.byref(p))
            # copy into numpy so the problems outlive the prepared set
            arrs = dict(
                vars_=np.ctypeslib.as_array(p.vars, (p.n_vars,)).copy() if p.n_vars else np.zeros(0),
                kind=np.ctypeslib.as_array(p.kind, (p.n_expr,)).copy() if p.n_expr else np.zeros(0, np.uint8),
                idx=np.ctypeslib.as_array(p.idx, (p.n_expr * 4,)).copy() if p.n_expr else np.zeros(0, np.uint32),
                param=np.ctypeslib.as_array(p.param, (p.n_expr,)).copy() if p.n_expr else np.zeros(0),
                free_vars=np.ctypeslib.as_array(p.free_vars, (p.n_free,)).copy() if p.n_free else np.zeros(0, np.uint32),
                rows=np.ctypeslib.as_array(p.rows, (p.n_rows,)).copy() if p.n_rows else np.zeros(0, np.uint32),
            )
            q, keep = make_problem(**arrs)
            out.append((q, scale, keep))
        lib().orc_prepared_free(C.c_void_p(ps))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_system_free(C.c_void_p(self.h))
            self.h = None
