"""Oracle restatement of System::analyze (fiksi/src/analyze/numerical/mod.rs) pinned by the
reference's own expectation (fiksi/src/tests/basic.rs:90-112: the constraint added last is the one
reported) and by the algebra of the elimination."""
import numpy as np


def _overconstrained(System):
    s = System()
    p = [s.add_point(0.123, 0.1), s.add_point(1.2, 0.0), s.add_point(-0.5, 1.1), s.add_point(1.599, 1.2)]
    s.point_point_distance(p[0], p[1], 1.0)
    s.point_point_distance(p[0], p[2], 1.5)
    s.point_point_distance(p[1], p[3], 1.7)
    s.point_point_distance(p[2], p[3], 1.2)
    s.point_point_distance(p[1], p[2], 2.0)
    last = s.point_point_distance(p[0], p[3], 5.0)
    return s, last


def test_reference_overconstrained_scenario(oracle):
    s, last = _overconstrained(oracle.System)
    assert s.analyze() == [last]          # basic.rs:107-111


def test_well_constrained_and_duplicate(oracle):
    s = oracle.System()
    a, b, c = s.add_point(0, 0), s.add_point(1, 0.5), s.add_point(0.3, 1.1)
    s.point_point_distance(a, b, 1.0); s.point_point_distance(a, c, 1.0); s.point_point_distance(b, c, 1.0)
    assert s.analyze() == []
    dup = s.point_point_distance(a, b, 1.0)   # same gradient row twice -> dependent
    assert s.analyze() == [dup]
    # collinear start (tests/singular.rs): the three distance rows are dependent AT this configuration
    s2 = oracle.System()
    a, b, c = s2.add_point(0, 0), s2.add_point(1, 0.5), s2.add_point(2, 1)
    s2.point_point_distance(a, b, 1.0); s2.point_point_distance(a, c, 1.0)
    third = s2.point_point_distance(b, c, 1.0)
    assert s2.analyze() == [third]


def test_gauss_jordan_properties(oracle):
    rng = np.random.default_rng(7)
    for m, n in [(3, 5), (6, 6), (9, 4), (7, 12)]:
        A = rng.standard_normal((m, n))
        if m >= 3:
            A[2] = 2.0 * A[0] - A[1]          # a dependent row
        R, ci, inc = oracle.gauss_jordan(A)
        k = min(m, n)
        rank = int(np.linalg.matrix_rank(A[:k]))
        assert int(inc.sum()) == rank and not inc[k:].any()   # rows beyond min(m, n) are never examined (:64)
        if m >= 3 and k > 2:
            assert not inc[2]
        piv = ci[:rank].astype(int)            # reduced row echelon form on the pivot columns
        rows = np.flatnonzero(inc)
        assert np.allclose(R[np.ix_(rows, piv)], np.eye(rank), atol=1e-9)
        assert sorted(ci.tolist()) == list(range(n))


def test_gradient_assignment_quirk(oracle):
    # a degenerate line (both end points the same element): the dense path ASSIGNS per slot
    # (expressions.rs:1003-1007), so the later slot wins and the row is not the summed sparse row
    s = oracle.System()
    p, q = s.add_point(0.3, 0.4), s.add_point(1.0, 2.0)
    l = s.add_line(q, q)
    s.point_line_incidence(p, l)
    vars_ = s.variables
    ind = oracle.analyze(vars_, [3], [[0, 2, 2, 0]], [0.0])
    assert ind.shape == (1,)
