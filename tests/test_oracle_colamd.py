"""Oracle COLAMD restatement against the reference's known-answer tests
(colamd_rs/src/lib.rs:253-321, colamd_rs/src/status.rs:167-237)."""
import numpy as np


def test_colamd_known_value(oracle):
    # colamd_rs/src/lib.rs:253-267 (A_LEN = 100, matches the C original's output)
    ok, p, stats = oracle.colamd(5, 4, [0, 1, 4, 2, 4, 0, 1, 2, 3, 1, 3], [0, 3, 5, 9, 11], a_len=100)
    assert ok
    assert p.tolist() == [1, 0, 2, 3, -1]


def test_colamd_known_value_recommended_len(oracle):
    # doc example colamd_rs/src/lib.rs:88-102 uses colamd_recommended
    ok, p, _ = oracle.colamd(5, 4, [0, 1, 4, 2, 4, 0, 1, 2, 3, 1, 3], [0, 3, 5, 9, 11])
    assert ok and p.tolist() == [1, 0, 2, 3, -1]


def test_colamd_no_aggressive_absorption(oracle):
    # colamd_rs/src/lib.rs:269-281
    ok, p, _ = oracle.colamd(4, 3, [0, 1, 2, 1, 0, 1, 3], [0, 3, 4, 7], aggressive=False)
    assert ok and p.tolist() == [1, 2, 0, -1]


def test_symamd_known_value(oracle):
    # colamd_rs/src/lib.rs:283-321 (two encodings of the same matrix)
    ok, perm, _ = oracle.symamd(5, [1, 2, 3, 4], [0, 1, 3, 3, 4, 4])
    assert ok and perm.tolist() == [0, 2, 1, 3, 4, -1]
    ok, perm, _ = oracle.symamd(5, [0, 1, 2, 3, 1, 4, 0, 3, 4], [0, 2, 4, 5, 6, 9])
    assert ok and perm.tolist() == [0, 2, 1, 3, 4, -1]
    # doctest colamd_rs/src/lib.rs:165-176
    ok, perm, _ = oracle.symamd(5, [0, 1, 0, 2, 3, 1, 1, 4, 3], [0, 2, 5, 6, 8, 9])
    assert ok and perm.tolist() == [0, 2, 1, 3, 4, -1]


def test_recommended_and_required_size(oracle):
    # status.rs:167-177: required length 57 for nnz 5, 4x3
    a_len = oracle.colamd_recommended(5, 4, 3)
    assert a_len == 2 * 5 + 6 * 4 + 4 * 5 + 3 + 1
    ok, _, stats = oracle.colamd(4, 3, [0, 1, 1, 10, 2], [0, 2, 3, 5], a_len=0)
    assert not ok and stats[3] == -7 and stats[4] == 57 and stats[5] == 0
    assert oracle.colamd_recommended(-1, 1, 1) is None


def test_error_codes(oracle):
    # status.rs:179-200
    ok, _, stats = oracle.colamd(4, 3, [0, 1, 1, 10, 2], [0, 2, 3, 5])
    assert not ok and stats[3] == -9 and (stats[4], stats[5], stats[6]) == (2, 10, 4)
    ok, _, stats = oracle.colamd(4, 3, [0, 1, 1, 0, 2], [2, 2, 3, 5])
    assert not ok and stats[3] == -6 and stats[4] == 2
    ok, _, stats = oracle.colamd(4, 3, [0, 1, 1, 0, 2], [0, 2, 0, 5])
    assert not ok and stats[3] == -8 and (stats[4], stats[5]) == (1, -2)


def test_jumbled_statistics(oracle):
    # status.rs:203-237
    ok, _, stats = oracle.colamd(4, 3, [0, 1, 1, 0, 1, 2, 3], [0, 2, 3, 7])
    assert ok and stats[3] == 0
    ok, _, stats = oracle.colamd(4, 3, [0, 1, 1, 2, 1, 0, 3], [0, 2, 3, 7])
    assert ok and stats[3] == 1 and (stats[4], stats[5], stats[6]) == (2, 0, 2)
    ok, _, stats = oracle.colamd(4, 3, [0, 1, 1, 0, 1, 1, 2], [0, 2, 3, 7])
    assert ok and stats[3] == 1 and (stats[4], stats[5], stats[6]) == (2, 1, 1)


def test_permutation_is_valid_on_random_patterns(oracle):
    rng = np.random.default_rng(7)
    for _ in range(30):
        m, n = int(rng.integers(1, 40)), int(rng.integers(1, 30))
        cols = [sorted(set(rng.integers(0, m, size=int(rng.integers(0, min(m, 6) + 1))).tolist())) for _ in range(n)]
        ptr = np.cumsum([0] + [len(c) for c in cols])
        rows = [r for c in cols for r in c]
        ok, p, stats = oracle.colamd(m, n, rows, ptr)
        assert ok and stats[3] == 0
        assert sorted(p[:n].tolist()) == list(range(n))
