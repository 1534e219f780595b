"""fk_system_* (the host-side mirror of fiksi::System above the C ABI) against the oracle's System.
Graph bookkeeping is host logic and is checked without a GPU; solves run on the GPU."""
import math

import numpy as np
import pytest

import scenarios as sc
import fiksi_b200 as fk


@pytest.mark.parametrize("name", sorted(sc.ALL))
def test_components_match_reference_graph(oracle, name):
    # fiksi/src/graph.rs:178-254 incl. the stale-index behaviour
    a = sc.ALL[name](oracle.System)["s"].components()
    b = sc.ALL[name](fk.System)["s"].components()
    assert a == b


def test_variable_layout_and_invalid_arguments():
    s = fk.System()
    p0, p1 = s.add_point(1., 2.), s.add_point(3., 4.)
    r = s.add_length(5.)
    line = s.add_line(p0, p1)
    circle = s.add_circle(p1, r)
    assert s.variables.tolist() == [1., 2., 3., 4., 5.]  # Length 1 var, Point 2, Line/Circle 0 (lib.rs:363-407)
    assert s.element_variable(p1) == 2 and s.element_variable(r) == 4
    with pytest.raises(ValueError):
        s.add_line(p0, r)            # a line needs two points
    with pytest.raises(ValueError):
        s.point_line_incidence(p0, circle)
    with pytest.raises(ValueError):
        s.add_circle(r, p0)
    c = s.point_line_distance(p0, line, 1.0)
    assert c == 0 and s.num_constraints() == 1


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(sc.ALL))
def test_system_solve_matches_oracle(oracle, name):
    bo = sc.ALL[name](oracle.System)
    bg = sc.ALL[name](fk.System)
    so, sg = bo["s"], bg["s"]
    so.solve()
    sg.solve()
    ro, rg = so.reports(), sg.reports()
    assert len(ro) == len(rg)
    for a, b in zip(ro, rg):
        assert a["exit_reason"] == b["exit_reason"] and a["trace_hash"] == b["trace_hash"], (name, a, b)
        assert a["lambda"] == b["lambda"]
    vo, vg = so.variables, sg.variables
    assert np.max(np.abs(vo - vg)) <= 1e-9 * np.max(np.abs(vo))
    res_o = np.array([so.calculate_residual(c) for c in bo["constraints"]])
    res_g = sg.residuals()[bg["constraints"]]
    scale = max(1.0, float(np.max(np.abs(vo)))) ** 2
    assert np.max(np.abs(res_o - res_g)) <= 1e-7 * scale


@pytest.mark.gpu
def test_fixed_elements_stay_bit_identical():
    # tests/fixed.rs:36-40,66-75
    b = sc.single_triangle(fk.System, fixed=1)
    b["s"].solve()
    assert b["s"].point(b["points"][1]) == (1., 0.5)
    assert sc.rms(b["s"].residuals()) < 1e-4
    b = sc.fixed_point_and_circle_center_incidence(fk.System)
    s = b["s"]
    s.solve()
    assert s.point(b["points"][0]) == (0., 0.) and s.point(b["points"][1]) == (4., 3.)
    assert abs(s.variables[s.element_variable(b["radius"])] - 5.) < 1e-4
    s.unfix(b["points"][0])
    s.solve()
    assert s.reports()[0]["exit_reason"] == 0 and abs(s.residuals()[0]) < 1e-3


@pytest.mark.gpu
def test_reference_thresholds_through_the_facade():
    # the reference's own assertions (RESIDUAL_THRESHOLD = 1e-4, fiksi/src/tests/mod.rs:13)
    for name, make in sc.REFERENCE_SOLVED.items():
        b = make(fk.System)
        b["s"].solve()
        assert sc.rms(b["s"].residuals()[b["constraints"]]) < 1e-4, name
    b = sc.overconstrained_triangle_line_incidence(fk.System)
    b["s"].solve()
    res = b["s"].residuals()
    assert sc.rms(res[:3]) >= 1e-4 and res[3] < 1e-4  # tests/basic.rs:77-86


@pytest.mark.gpu
def test_update_value_and_parameter_then_resolve():
    # fiksi/benches/fiksi_bench.rs:34-38 resets element values and solves again
    b = sc.hinged_triangles_bench(fk.System, 4)
    s = b["s"]
    start = s.variables.copy()
    s.solve()
    first = s.variables.copy()
    for i, v in enumerate(start):
        s.set_variable(i, float(v))
    s.solve()
    assert np.array_equal(s.variables, first)  # deterministic: Rng::from_seed(42) per solve
    s.set_parameter(b["constraints"][2], 2.5)
    s.solve()
    assert abs(s.residuals()[b["constraints"][2]]) < 1e-4
    p1, p2 = s.point(b["points"][1]), s.point(b["points"][2])
    assert abs(math.hypot(p1[0] - p2[0], p1[1] - p2[1]) - 2.5) < 1e-3
