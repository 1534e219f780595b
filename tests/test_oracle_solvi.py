"""Oracle solvi restatement against the reference's own known-answer tests:
solvi/src/decomposition/sparse/cholesky.rs:602-796, qr.rs:376-652, sparse_col_mat.rs:835-869,
utils.rs doctests, permutation.rs:96-125, triplet_mat.rs."""
import json
import math
import os

import numpy as np
import pytest

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


def _csc(cols):
    ptr = np.cumsum([0] + [len(c) for c in cols])
    rows = [r for c in cols for r in c]
    return ptr, rows


def test_davis_fig1_known_matrix(oracle):
    k = KATS["davis_fig1"]
    ptr, rows = _csc(k["columns"])
    s = oracle.Symbolic(k["nrows"], k["ncols"], ptr, rows, "natural")
    assert s.get("parents").tolist() == k["parents"]
    assert s.get("row_counts").tolist() == k["row_counts"]
    assert s.get("col_counts").tolist() == k["col_counts"]
    assert s.get("r_rowidx").tolist() == k["r_row_indices"]
    post = s.get("postorder").tolist()
    assert sorted(post) == list(range(12))
    pos = {c: i for i, c in enumerate(post)}
    for c, p in enumerate(k["parents"]):
        if p >= 0:
            assert pos[c] < pos[p]


def test_dense_known_matrix(oracle):
    s = oracle.Symbolic(3, 3, [0, 3, 6, 9], [0, 1, 2, 0, 1, 2, 0, 1, 2])
    assert s.get("row_counts").tolist() == [3, 2, 1]
    assert s.get("col_counts").tolist() == [1, 2, 3]
    assert s.get("r_rowidx").tolist() == [0, 0, 1, 0, 1, 2]


def test_sparse_known_matrix(oracle):
    s = oracle.Symbolic(10, 6, [0, 2, 4, 7, 10, 13, 16], [0, 4, 0, 5, 0, 1, 6, 0, 1, 7, 1, 2, 8, 1, 3, 9])
    assert s.get("r_rowidx").tolist() == [0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 2, 3, 4, 2, 3, 4, 5]


def test_qr_underdetermined_damped(oracle):
    # qr.rs:376-415
    shape, ptr, rows, vals = oracle.from_triplets(5, 3, [0, 0, 1, 2, 3, 4], [0, 2, 2, 0, 1, 2], [2., 5., 5., 1., 1., 1.])
    assert shape == (5, 3)
    s = oracle.Symbolic(5, 3, ptr, rows)
    assert s.get("r_colptr").tolist() == [0, 1, 2, 4]
    assert s.get("r_rowidx").tolist() == [0, 1, 0, 2]
    s.factorize(vals)
    exp = [math.sqrt(5.), 1., 2. * math.sqrt(5.), math.sqrt(31.)]
    assert np.allclose(np.abs(s.r_values()), exp, atol=1e-8, rtol=0)


def test_qr_underdetermined_rank_deficient(oracle):
    # qr.rs:417-465: TripletMat::new(5, 3) keeps the 5x3 shape
    shape, ptr, rows, vals = oracle.from_triplets(5, 3, [0, 0, 0, 1, 1, 1], [0, 1, 2, 0, 1, 2], [3., 5., 3., 1., 3., 2.])
    s = oracle.Symbolic(5, 3, ptr, rows)
    assert s.get("r_colptr").tolist() == [0, 1, 3, 6]
    assert s.get("r_rowidx").tolist() == [0, 0, 1, 0, 1, 2]
    s.factorize(vals)
    q = math.sqrt(10.)
    exp = [q, 9. / 5. * q, 2. / 5. * q, 11. / 10. * q, 3. / 10. * q, 0.]
    assert np.allclose(np.abs(s.r_values()), exp, atol=1e-8, rtol=0)


@pytest.mark.parametrize("ordering", ["natural", "colamd"])
def test_qr_big_underdetermined_damped(oracle, ordering):
    # qr.rs:467-652: a real fiksi augmented Jacobian (9 rows + 12 damping rows at sqrt(0.5))
    k = KATS["big_underdetermined_damped"]
    s = oracle.Symbolic(k["nrows"], k["ncols"], k["column_pointers"], k["row_indices"], ordering)
    if ordering == "natural":
        assert s.get("r_colptr").tolist() == k["r_column_pointers"]
        assert s.get("r_rowidx").tolist() == k["r_row_indices"]
    s.factorize(k["values"])
    if ordering == "natural":
        assert np.allclose(np.abs(s.r_values()), np.abs(k["expected_abs_r_values"]), atol=1e-8, rtol=0)
    b = np.zeros(21)
    b[:9] = k["b"]
    ok, x = s.solve(b)
    assert ok
    assert np.allclose(x[:12], k["x_expected"], atol=1e-8, rtol=0)
    # and it is the damped least-squares solution
    A = np.zeros((21, 12))
    for c in range(12):
        for p in range(k["column_pointers"][c], k["column_pointers"][c + 1]):
            A[k["row_indices"][p], c] = k["values"][p]
    ref = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.allclose(x[:12], ref, atol=1e-12)


def test_upper_triangular_solve(oracle):
    # sparse_col_mat.rs:835-869 / doctest :760-786
    shape, ptr, rows, vals = oracle.from_triplets(3, 3, [0, 0, 1, 1, 2], [0, 1, 1, 2, 2], [1., -2., 4., 1., 2.])
    ok, x = oracle.upper_solve(ptr, rows, vals, [2., 1., 4.])
    assert ok and np.allclose(x, [1.5, -0.25, 2.0])
    # zero diagonal -> reported unsolvable
    shape, ptr, rows, vals = oracle.from_triplets(2, 2, [0, 0], [0, 1], [1., 1.])
    ok, _ = oracle.upper_solve(ptr, rows, vals, [1., 1.])
    assert not ok


def test_from_triplets_sums_duplicates_and_sorts(oracle):
    shape, ptr, rows, vals = oracle.from_triplets(0, 0, [2, 0, 2, 1], [1, 1, 1, 0], [1., 2., 3., 4.])
    assert shape == (3, 2)  # triplet_mat.rs:99-105: shape grows on push
    assert ptr.tolist() == [0, 1, 3] and rows.tolist() == [1, 0, 2] and vals.tolist() == [4., 2., 4.]
    # empty leading/trailing columns keep valid pointers
    shape, ptr, rows, vals = oracle.from_triplets(2, 4, [1], [2], [7.])
    assert ptr.tolist() == [0, 0, 0, 1, 1]


def test_post_order_and_levels(oracle):
    # utils.rs:43-46,147-151
    parents = [1, 5, 5, 4, 5, 6, -1]
    post = oracle.post_order(parents).tolist()
    assert post == [3, 4, 2, 0, 1, 5, 6]  # children visited highest-numbered first
    levels, mx = oracle.node_depth_levels(parents)
    assert levels.tolist() == [3, 2, 2, 3, 2, 1, 0] and mx == 3


def test_permutation_gather_semantics(oracle):
    # permutation.rs doc example: p[i] = a[permutation[i]]
    perm = [2, 3, 4, 1, 0]
    data = [10., 11., 12., 13., 14.]
    exp = [data[p] for p in perm]
    assert oracle.permute(perm, data, "gather").tolist() == exp
    assert oracle.permute(perm, data, "swaps").tolist() == exp
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 33):
        p = rng.permutation(n)
        d = rng.normal(size=n)
        assert np.array_equal(oracle.permute(p, d, "gather"), d[p])
        assert np.array_equal(oracle.permute(p, d, "swaps"), d[p])


def test_qr_matches_dense_lstsq_on_random_augmented(oracle):
    rng = np.random.default_rng(11)
    for trial in range(10):
        m, n = int(rng.integers(3, 25)), int(rng.integers(2, 15))
        dense = rng.normal(size=(m, n)) * (rng.random((m, n)) < 0.3)
        A = np.vstack([dense, 0.7 * np.eye(n)])
        r, c = np.nonzero(A)
        shape, ptr, rows, vals = oracle.from_triplets(m + n, n, r, c, A[r, c])
        for ordering in ("natural", "colamd"):
            s = oracle.Symbolic(m + n, n, ptr, rows, ordering)
            s.factorize(vals)
            b = rng.normal(size=m + n)
            ok, x = s.solve(b)
            assert ok
            ref = np.linalg.lstsq(A, b, rcond=None)[0]
            assert np.allclose(x[:n], ref, atol=1e-10)
            # R pattern from the symbolic analysis covers chol(A^T A) of the permuted matrix
            perm = s.get("col_permutation")
            L = np.linalg.cholesky((A[:, perm]).T @ A[:, perm])
            rp, ri = s.get("r_colptr"), s.get("r_rowidx")
            pat = np.zeros((n, n), bool)
            for j in range(n):
                pat[ri[rp[j]:rp[j + 1]], j] = True
            assert not np.any((np.abs(L.T) > 1e-13) & ~pat)
