"""The CPU oracle (oracle/sw_oracle.c) pinned to everything the reference offers for this path:
golden heads of test/reference/*.txt, the survey probe's 20 x 111 vectors, the self-scores, the
Wikipedia pair, and the scores / aligned strings printed by the compiled reference cpu.cpp."""
import json
import os

import numpy as np

from conftest import GOLDEN
from oracle_lib import pack_db


def _gold(name):
    return np.array([int(x) for x in open(os.path.join(GOLDEN, name + ".head111.txt")).read().split()], dtype=np.int32)


def test_golden_heads(oracle, subset, queries):
    m = oracle.matrix("blosum50")
    for name, total in (("P01008", 31802), ("P02232", 18440)):
        got = oracle.scan(oracle.encode(queries[name]), subset["codes"], subset["offsets"], m)
        assert np.array_equal(got, _gold(name))
        assert int(got.sum()) == total
    assert list(_gold("P01008")[:5]) == [364, 368, 550, 223, 509]
    assert list(_gold("P02232")[:5]) == [192, 206, 208, 165, 227]


def test_survey_vectors_all_queries(oracle, subset, queries, survey_exp):
    m = oracle.matrix("blosum50")
    assert len(survey_exp) == 20
    for name, want in survey_exp.items():
        got = oracle.scan(oracle.encode(queries[name]), subset["codes"], subset["offsets"], m)
        assert np.array_equal(got, np.array(want, dtype=np.int32)), name


def test_self_scores_and_wikipedia_pair(oracle, queries):
    m = oracle.matrix("blosum50")
    for name, want in (("P01008", 3037), ("P02232", 910)):
        q = oracle.encode(queries[name])
        assert oracle.score(q, q, m) == want
    assert oracle.score(oracle.encode("GGTTGACTA"), oracle.encode("TGTTACGG"), m) == 34
    mi = oracle.matrix("ident3")
    assert oracle.score(oracle.encode("GGTTGACTA", "ident3"), oracle.encode("TGTTACGG", "ident3"), mi) == 13


def test_against_compiled_cpu_cpp(oracle, subset, queries):
    ref = json.load(open(os.path.join(GOLDEN, "cpu_ref_ident3.json")))
    mi = oracle.matrix("ident3")
    for p in ref["pairs"]:
        s, a, b, _ = oracle.align(p["a"], p["b"], "ident3")
        assert (s, a, b) == (p["score"], p["aligned_a"], p["aligned_b"])
        assert oracle.score(oracle.encode(p["a"], "ident3"), oracle.encode(p["b"], "ident3"), mi) == p["score"]
    codes, offs = pack_db([oracle.encode(s, "ident3") for s in subset["seqs"]])
    for name, want in ref["scans"].items():
        got = oracle.scan(oracle.encode(queries[name], "ident3"), codes, offs, mi)
        assert np.array_equal(got, np.array(want, dtype=np.int32)), name


def test_matrix_properties(oracle):
    m = oracle.matrix("blosum50").astype(int)
    assert (m == m.T).all() and m.min() == -5 and m.max() == 15
    assert (m[24:, :] == 0).all() and (m[:, 24:] == 0).all()  # '*' and spare codes score 0 (SWSolver.cu:80)
    assert m[17, 17] == 15 and m[4, 4] == 13 and m[0, 0] == 5
    mi = oracle.matrix("ident3").astype(int)
    assert (np.diag(mi)[:31] == 3).all() and mi[31].max() == 0 and mi[0, 1] == -3


def test_padding_is_score_neutral(oracle, queries):
    """'/' padding of FASTAParsers.h:94-96 / SWSolver.cu:268-269 encodes to '*' and cannot change a score"""
    m = oracle.matrix("blosum50")
    q = queries["P02232"]
    d = queries["P05013"]
    base = oracle.score(oracle.encode(q), oracle.encode(d), m)
    assert oracle.score(oracle.encode(q + "////"), oracle.encode(d + "///////"), m) == base
    assert list(oracle.encode("AUO/z*\r")) == [0, 24, 24, 24, 24, 24, 24]


def _gotoh_py(q, d, m, go, ge):
    """independent restatement (pure Python, full matrices) of the affine recurrences in oracle/sw_oracle.c"""
    neg = -10 ** 9
    H = [[0] * (len(d) + 1) for _ in range(len(q) + 1)]
    E = [[neg] * (len(d) + 1) for _ in range(len(q) + 1)]
    F = [[neg] * (len(d) + 1) for _ in range(len(q) + 1)]
    best = 0
    for i in range(1, len(q) + 1):
        for j in range(1, len(d) + 1):
            E[i][j] = max(E[i][j - 1] - ge, H[i][j - 1] - go)
            F[i][j] = max(F[i - 1][j] - ge, H[i - 1][j] - go)
            H[i][j] = max(0, H[i - 1][j - 1] + int(m[q[i - 1], d[j - 1]]), E[i][j], F[i][j])
            best = max(best, H[i][j])
    return best


def test_affine_restatement_against_independent_gotoh(oracle):
    """the reference has no affine mode (SWSolver.cu:8 is a comment), so the C restatement is pinned against a second,
    independent implementation and against two hand-checked cases"""
    from oracle_lib import pack_db
    m = oracle.matrix("blosum50")
    rng = np.random.default_rng(11)
    seqs = [rng.integers(0, 5, int(n)).astype(np.uint8) for n in (0, 1, 7, 30, 64, 90)]
    codes, offs = pack_db(seqs)
    for ql in (1, 9, 40):
        q = rng.integers(0, 5, ql).astype(np.uint8)
        for go, ge in ((10, 2), (4, 1), (3, 0), (2, 2)):
            got = oracle.scan_affine(q, codes, offs, m, go, ge)
            want = [_gotoh_py(q, s, m, go, ge) for s in seqs]
            assert got.tolist() == want, (ql, go, ge)
    # hand-checked: AAAA vs AAGGAA with A:A = 5, gap of two: open 3 + extend 1 -> 4 * 5 - 4 = 16 (ungapped best: 10);
    # with open 12 the gap no longer pays: 10
    a = np.zeros(4, np.uint8)
    d = np.array([0, 0, 7, 7, 0, 0], np.uint8)
    c, o = pack_db([d])
    assert oracle.scan_affine(a, c, o, m, 3, 1)[0] == 16
    assert oracle.scan_affine(a, c, o, m, 12, 1)[0] == 10


def test_published_textbook_vector_blosum50_gap8(oracle):
    """An externally published known answer for the linear recurrence under BLOSUM50: Durbin, Eddy, Krogh & Mitchison,
    'Biological sequence analysis' (1998), fig. 2.6 -- HEAGAWGHEE against PAWHEAE with gap penalty d = 8 has the best
    local alignment AWGHE / AW-HE with score 28. Same matrix as the reference (SWSolver.cu:54-81), another gap than its
    g = 2, so it pins the gap handling of the restatement independently of the reference's goldens. The affine
    restatement with open == extend == 8 must give the same 28."""
    m = oracle.matrix("blosum50")
    s, a, b, end = oracle.align("HEAGAWGHEE", "PAWHEAE", "blosum50", gap=8)
    assert (s, a, b) == (28, "AWGHE", "AW-HE") and end == (9, 5)
    q, d = oracle.encode("HEAGAWGHEE"), oracle.encode("PAWHEAE")
    assert oracle.score(q, d, m, gap=8) == 28
    codes, offs = pack_db([d])
    assert oracle.scan_affine(q, codes, offs, m, 8, 8)[0] == 28


def test_affine_is_sandwiched_by_the_pinned_linear_oracle(oracle, subset, queries):
    """No affine vectors exist in the reference ("parity unpinned"), but two facts tie the Gotoh restatement to the pinned
    linear one: open == extend IS the linear recurrence, and a gap of length L costs between L * extend and L * open, so
    linear(open) <= affine(open, extend) <= linear(extend) for every pair."""
    m = oracle.matrix("blosum50")
    codes, offs = pack_db([oracle.encode(s) for s in subset["seqs"][:40]])
    for name in ("P02232", "P01008"):
        q = oracle.encode(queries[name])
        lin2 = oracle.scan(q, codes, offs, m)  # the reference's gap 2, pinned by its goldens
        assert np.array_equal(oracle.scan_affine(q, codes, offs, m, 2, 2), lin2)
        lin10 = oracle.scan_affine(q, codes, offs, m, 10, 10)
        aff = oracle.scan_affine(q, codes, offs, m, 10, 2)
        assert (lin10 <= aff).all() and (aff <= lin2).all() and (lin10 < lin2).any()


def _score_of_alignment(a, b, m, enc, go, ge):
    """score of two aligned strings under an affine model (a gap of length L costs go + (L-1) ge)"""
    total, gap_a, gap_b = 0, False, False
    for x, y in zip(a, b):
        if x == "-":
            total -= ge if gap_a else go
            gap_a, gap_b = True, False
        elif y == "-":
            total -= ge if gap_b else go
            gap_a, gap_b = False, True
        else:
            total += int(m[enc(x)[0], enc(y)[0]])
            gap_a = gap_b = False
    return total


def test_affine_traceback_restatement(oracle, subset, queries):
    """swo_align_affine (cpu.cpp:39-103 extended to Gotoh's three states; parity unpinned like the affine scan): its
    open == extend case must be the linear traceback that the compiled cpu.cpp pins, character for character; in
    general the score must equal the score-only recurrence, the printed alignment must re-score to it, and ungapped
    both strings must be substrings ending at the reported cell."""
    ref = json.load(open(os.path.join(GOLDEN, "cpu_ref_ident3.json")))
    for p in ref["pairs"]:
        s, a, b, end = oracle.align_affine(p["a"], p["b"], 2, 2, "ident3")
        assert (s, a, b) == (p["score"], p["aligned_a"], p["aligned_b"])
    m = oracle.matrix("blosum50")
    enc = lambda ch: oracle.encode(ch)
    for name in ("P02232", "P01008"):
        for sid in (0, 13, 16, 56):
            subj = subset["seqs"][sid]
            lin = oracle.align(queries[name], subj, "blosum50")
            assert oracle.align_affine(queries[name], subj, 2, 2) == lin
            for go, ge in ((10, 2), (5, 1), (12, 0)):
                s, a, b, (ei, ej) = oracle.align_affine(queries[name], subj, go, ge)
                c, o = pack_db([oracle.encode(subj)])
                assert s == oracle.scan_affine(oracle.encode(queries[name]), c, o, m, go, ge)[0]
                assert _score_of_alignment(a, b, m, enc, go, ge) == s
                ua, ub = a.replace("-", ""), b.replace("-", "")
                assert queries[name][ei - len(ua):ei] == ua and subj[ej - len(ub):ej] == ub
    # hand-checked: AAAA x AAGGAA with open 3 / extend 1 joins the two halves through a two-residue gap (16)
    assert oracle.align_affine("AAAA", "AAGGAA", 3, 1)[:3] == (16, "AA--AA", "AAGGAA")
    assert oracle.align_affine("HEAGAWGHEE", "PAWHEAE", 8, 8)[:3] == (28, "AWGHE", "AW-HE")
