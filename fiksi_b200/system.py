"""Python mirror of fiksi::System over the C ABI's fk_system_* functions (same operations, names
and argument order as the reference's constructors; see include/fiksi_b200.h).  Same duck-typed
interface as ``oracle.System`` so that tests/scenarios.py builds both."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import FkReport, REPORT_DTYPE, check, lib, ptr

_BAD = 0xFFFFFFFF


class System:
    def __init__(self):
        self._h = C.c_void_p()
        check(lib().fk_system_create(C.byref(self._h)))
        self._reports = np.zeros(0, dtype=REPORT_DTYPE)

    def _id(self, v):
        if v == _BAD:
            raise ValueError("invalid element / constraint arguments")
        return v

    # elements (fiksi/src/elements/mod.rs:280-454)
    def add_point(self, x, y): return self._id(lib().fk_system_add_point(self._h, C.c_double(x), C.c_double(y)))
    def add_length(self, l): return self._id(lib().fk_system_add_length(self._h, C.c_double(l)))
    def add_line(self, p1, p2): return self._id(lib().fk_system_add_line(self._h, p1, p2))
    def add_circle(self, c, r): return self._id(lib().fk_system_add_circle(self._h, c, r))
    def fix(self, e): check(lib().fk_system_fix(self._h, e, 1))
    def unfix(self, e): check(lib().fk_system_fix(self._h, e, 0))

    def _c(self, tag, elements, param=0.0):
        arr = (C.c_uint32 * len(elements))(*elements)
        return self._id(lib().fk_system_add_constraint(self._h, tag, arr, len(elements), C.c_double(param)))

    # constraints (fiksi/src/constraints/mod.rs:317-891)
    def point_point_coincidence(self, a, b): return self._c(0, [a, b])
    def point_point_distance(self, a, b, d): return self._c(1, [a, b], d)
    def point_point_point_angle(self, a, b, c, ang): return self._c(2, [a, b, c], ang)
    def point_line_incidence(self, p, l): return self._c(3, [p, l])
    def point_line_distance(self, p, l, d): return self._c(4, [p, l], d)
    def point_circle_incidence(self, p, c): return self._c(5, [p, c])
    def segment_segment_length_equality(self, a, b, c, d): return self._c(6, [a, b, c, d])
    def line_line_angle(self, a, b, ang): return self._c(7, [a, b], ang)
    def line_line_parallelism(self, a, b): return self._c(8, [a, b])
    def line_line_perpendicularity(self, a, b): return self._c(9, [a, b])
    def line_circle_tangency(self, l, c): return self._c(10, [l, c])

    @property
    def variables(self):
        n = lib().fk_system_num_variables(self._h)
        out = np.zeros(max(n, 1))
        check(lib().fk_system_get_variables(self._h, ptr(out, C.c_double)))
        return out[:n]

    def set_variable(self, i, v): check(lib().fk_system_set_variable(self._h, i, C.c_double(v)))
    def set_parameter(self, c, v): check(lib().fk_system_set_parameter(self._h, c, C.c_double(v)))
    def element_variable(self, e): return lib().fk_system_element_variable(self._h, e)
    def num_constraints(self): return lib().fk_system_num_constraints(self._h)

    def point(self, e):
        i = self.element_variable(e)
        v = self.variables
        return float(v[i]), float(v[i + 1])

    def solve(self, perturb=True):
        """System::solve(SolvingOptions::DEFAULT) (perturb=True) on the GPU."""
        cap = max(1, lib().fk_system_num_components(self._h))
        reps = np.zeros(cap, dtype=REPORT_DTYPE)
        n = C.c_uint32(0)
        check(lib().fk_system_solve(self._h, int(perturb), reps.ctypes.data_as(C.POINTER(FkReport)), cap, C.byref(n)))
        self._reports = reps[:n.value]

    def solve_single_pass(self, perturb=True):
        """System::solve with Decomposer::SinglePass (assemble/mod.rs:169-210) on the GPU."""
        cap = max(1, len(self.single_pass_plan()))
        reps = np.zeros(cap, dtype=REPORT_DTYPE)
        n = C.c_uint32(0)
        check(lib().fk_system_solve_opts(self._h, 1, int(perturb), reps.ctypes.data_as(C.POINTER(FkReport)), cap, C.byref(n)))
        self._reports = reps[:n.value]

    def solve_recursive_assembly(self, perturb=True):
        """System::solve with Decomposer::RecursiveAssembly (assemble/mod.rs:212-277) on the GPU."""
        cap = max(1, self.recursive_assembly_plan()[0])
        reps = np.zeros(cap, dtype=REPORT_DTYPE)
        n = C.c_uint32(0)
        check(lib().fk_system_solve_opts(self._h, 2, int(perturb), reps.ctypes.data_as(C.POINTER(FkReport)), cap, C.byref(n)))
        self._reports = reps[:n.value]

    def recursive_assembly_plan(self):
        """(number of steps, serialised plan words) of fk_system_recursive_assembly_plan (host only)."""
        nw, ns = C.c_uint32(0), C.c_uint32(0)
        check(lib().fk_system_recursive_assembly_plan(self._h, None, 0, C.byref(nw), C.byref(ns)))
        out = np.zeros(max(nw.value, 1), dtype=np.uint32)
        check(lib().fk_system_recursive_assembly_plan(self._h, ptr(out, C.c_uint32), nw.value, C.byref(nw), C.byref(ns)))
        return int(ns.value), out[:nw.value].tolist()

    def single_pass_plan(self):
        """[(free variables, expressions), ...] of fk_system_single_pass_plan (host only)."""
        sizes = np.zeros(3, dtype=np.uint32)
        check(lib().fk_system_single_pass_plan(self._h, ptr(sizes, C.c_uint32), None, None, None, None))
        n, nf, ne = (int(x) for x in sizes)
        fp = np.zeros(n + 1, np.uint32); fv = np.zeros(max(nf, 1), np.uint32)
        ep = np.zeros(n + 1, np.uint32); ex = np.zeros(max(ne, 1), np.uint32)
        check(lib().fk_system_single_pass_plan(self._h, ptr(sizes, C.c_uint32), ptr(fp, C.c_uint32), ptr(fv, C.c_uint32),
                                               ptr(ep, C.c_uint32), ptr(ex, C.c_uint32)))
        return [(fv[fp[k]:fp[k + 1]].tolist(), ex[ep[k]:ep[k + 1]].tolist()) for k in range(n)]

    def reports(self):
        return [{k: r[k].item() for k in REPORT_DTYPE.names} for r in self._reports]

    def residuals(self):
        n = self.num_constraints()
        out = np.zeros(max(n, 1))
        check(lib().fk_system_residuals(self._h, ptr(out, C.c_double)))
        return out[:n]

    def analyze(self):
        """System::analyze (lib.rs:454-458): ids of the constraints flagged as over-constraining."""
        cap = max(1, 2 * self.num_constraints())
        out = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint32(0)
        check(lib().fk_system_analyze(self._h, ptr(out, C.c_uint32), cap, C.byref(n)))
        return out[:n.value].tolist()

    def calculate_residual(self, c):
        return float(self.residuals()[c])

    def components(self):
        out = []
        for ci in range(lib().fk_system_num_components(self._h)):
            ne, nc = C.c_uint32(0), C.c_uint32(0)
            check(lib().fk_system_component(self._h, ci, C.byref(ne), None, C.byref(nc), None))
            el = np.zeros(max(ne.value, 1), np.uint32)
            co = np.zeros(max(nc.value, 1), np.uint32)
            check(lib().fk_system_component(self._h, ci, None, ptr(el, C.c_uint32), None, ptr(co, C.c_uint32)))
            out.append((el[:ne.value].tolist(), co[:nc.value].tolist()))
        return out

    def close(self):
        if self._h:
            lib().fk_system_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
