"""Transcribes the known-answer vectors that the reference's own unit tests hold for the solve
path into tests/golden/reference_kats.json.  Run in the build container (needs /root/reference):

    python tests/golden/extract_reference_kats.py

Sources: solvi/src/decomposition/sparse/qr.rs:467-652 (big_underdetermined_damped),
cholesky.rs:602-736 (Davis 2011 Fig. 1 matrix), fiksi/src/rand.rs:49-63 (LCG sequence).
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")


def numbers(text):
    return [float(x) if ("." in x or "e" in x) else int(x) for x in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", text)]


def block(src, start_pat, open_ch="[", close_ch="]"):
    i = src.index(start_pat)
    i = src.index(open_ch, i + len(start_pat) - 1 if start_pat.endswith(open_ch) else i)
    depth, j = 0, i
    while True:
        if src[j] == open_ch:
            depth += 1
        elif src[j] == close_ch:
            depth -= 1
            if depth == 0:
                return src[i + 1:j], j
        j += 1


def main():
    out = {}
    qr = open(os.path.join(REF, "solvi/src/decomposition/sparse/qr.rs")).read()
    t = qr[qr.index("fn big_underdetermined_damped"):]
    rows, e = block(t, "row_indices: vec![")
    cols, e2 = block(t, "column_pointers: vec![")
    vals, e3 = block(t[e2:], "values: vec![")
    t2 = t[e2 + e3:]
    b, _ = block(t2, "let b = [")
    rs = t2[t2.index("&SparseColMatStructure"):]
    r_rows, q = block(rs, "row_indices: vec![")
    r_cols, _ = block(rs[q:], "column_pointers: vec![")
    exp_r, q2 = block(t2, "let expected_r_values: &[f64] = &[")
    x_exp, _ = block(t2[q2:], "let x_expected = [")
    out["big_underdetermined_damped"] = {
        "source": "solvi/src/decomposition/sparse/qr.rs:467-652",
        "nrows": 21, "ncols": 12,
        "row_indices": numbers(rows), "column_pointers": numbers(cols), "values": numbers(vals),
        "b": numbers(b), "r_row_indices": numbers(r_rows), "r_column_pointers": numbers(r_cols),
        "expected_abs_r_values": numbers(exp_r), "x_expected": numbers(x_exp),
    }
    assert len(out["big_underdetermined_damped"]["values"]) == 54
    assert len(out["big_underdetermined_damped"]["expected_abs_r_values"]) == 70
    assert len(out["big_underdetermined_damped"]["x_expected"]) == 12

    ch = open(os.path.join(REF, "solvi/src/decomposition/sparse/cholesky.rs")).read()
    t = ch[ch.index("fn known_matrix"):ch.index("fn dense_known_matrix")]
    cols_txt, e = block(t, "let row_indices: [&'static [usize]; 12] = [")
    col_rows = [numbers(c) for c in re.findall(r"&\[([^\]]*)\]", cols_txt)]
    parents_txt = re.search(r"&\[(1, 5, 5, 4[^\]]*)\]", t).group(1)
    rc = re.search(r"l_counts\.row_counts, &\[([^\]]*)\]", t).group(1)
    cc = re.search(r"l_counts\.col_counts, &\[([^\]]*)\]", t).group(1)
    rr, _ = block(t[t.index("&l_structure.row_indices"):], "&[")
    out["davis_fig1"] = {
        "source": "solvi/src/decomposition/sparse/cholesky.rs:602-736",
        "nrows": 23, "ncols": 12, "columns": col_rows,
        "parents": [(-1 if "MAX" in p else int(p)) for p in parents_txt.split(",") if p.strip()],
        "row_counts": numbers(rc), "col_counts": numbers(cc), "r_row_indices": numbers(rr),
    }
    assert sum(len(c) for c in col_rows) == 78

    rnd = open(os.path.join(REF, "fiksi/src/rand.rs")).read()
    seq, _ = block(rnd[rnd.index("fn known_sequence"):], "let sequence = [")
    out["lcg_sequence"] = {"source": "fiksi/src/rand.rs:49-63",
                           "values": [int(x, 16) for x in re.findall(r"0x[0-9A-Fa-f]+", seq)]}
    json.dump(out, open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
