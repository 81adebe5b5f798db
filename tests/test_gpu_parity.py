"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors and the CPU oracle.
Bit-exact integer scores are the bar everywhere."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, random_db
from oracle_lib import pack_db

pytestmark = pytest.mark.gpu


def _gold(name):
    return np.array([int(x) for x in open(os.path.join(GOLDEN, name + ".head111.txt")).read().split()], dtype=np.int32)


def test_config1_golden_heads(engine, swb, subset, queries):
    """reference goldens (test/swissprot_tests.cpp:68-72), lines 1-111, both pinned queries"""
    engine.set_scoring_preset(swb.SWB_SCORING_BLOSUM50_REF)
    engine.db_load(subset["codes"], subset["offsets"])
    for name in ("P01008", "P02232"):
        got = engine.search(swb.encode(queries[name]))
        assert np.array_equal(got, _gold(name)), name


def test_config1_all_queries_vs_oracle_and_survey(engine, swb, oracle, subset, queries, survey_exp):
    engine.set_scoring_preset(swb.SWB_SCORING_BLOSUM50_REF)
    engine.db_load(subset["codes"], subset["offsets"])
    m = oracle.matrix("blosum50")
    names = sorted(queries)
    batch = engine.search_batch([swb.encode(queries[n]) for n in names])
    for i, name in enumerate(names):
        want = oracle.scan(oracle.encode(queries[name]), subset["codes"], subset["offsets"], m)
        assert np.array_equal(batch[i], want), name
        assert np.array_equal(batch[i], np.array(survey_exp[name], dtype=np.int32)), name
        single = engine.search(swb.encode(queries[name]))
        assert np.array_equal(single, want), name


def test_self_scores(engine, swb, queries):
    """golden maxima: P01008 x itself = 3037, P02232 x itself = 910"""
    engine.set_scoring_preset(swb.SWB_SCORING_BLOSUM50_REF)
    for name, want in (("P01008", 3037), ("P02232", 910)):
        q = swb.encode(queries[name])
        codes, offs = swb.pack_sequences([q])
        engine.db_load(codes, offs)
        assert engine.search(q)[0] == want


def test_ident3_matches_compiled_cpu_cpp(engine, swb, subset, queries):
    """+3/-3 scheme against the scores the compiled reference cpu.cpp printed (tests/golden/cpu_ref_ident3.json)"""
    import json
    ref = json.load(open(os.path.join(GOLDEN, "cpu_ref_ident3.json")))
    engine.set_scoring_preset(swb.SWB_SCORING_IDENT3)
    codes, offs = swb.pack_sequences([swb.encode(s, swb.SWB_SCORING_IDENT3) for s in subset["seqs"]])
    engine.db_load(codes, offs)
    for name, want in ref["scans"].items():
        got = engine.search(swb.encode(queries[name], swb.SWB_SCORING_IDENT3))
        assert np.array_equal(got, np.array(want, dtype=np.int32)), name
    for p in ref["pairs"]:
        c, o = swb.pack_sequences([swb.encode(p["b"], swb.SWB_SCORING_IDENT3)])
        engine.db_load(c, o)
        assert engine.search(swb.encode(p["a"], swb.SWB_SCORING_IDENT3))[0] == p["score"]
    engine.set_scoring_preset(swb.SWB_SCORING_BLOSUM50_REF)


@pytest.mark.parametrize("k", [0, 8, 16, 32])
@pytest.mark.parametrize("group_len", [16, 96, 384, 100000])
def test_random_db_all_kernel_shapes(swb, oracle, k, group_len):
    """every K (0 = per-group choice, several concurrent launch groups) and lane-group mix (group_len 16 forces
    32-lane wavefronts, 100000 forces one lane per pair)"""
    rng = np.random.default_rng(1000 + k + group_len)
    lens = np.concatenate([rng.integers(0, 40, 70), rng.integers(40, 700, 300), rng.integers(700, 3000, 9), [0, 1, 2, 3]])
    rng.shuffle(lens)
    enc = random_db(rng, lens)
    codes, offs = pack_db(enc)
    m = oracle.matrix("blosum50")
    e = swb.Engine(0, group_len=group_len, k=k)
    try:
        e.db_load(codes, offs)
        for ql in (1, 7, 33, 144, 257, 1000):
            q = rng.integers(0, 24, ql).astype(np.uint8)
            want = oracle.scan(q, codes, offs, m)
            # a database this small takes the small-shard path (pipelined passes for the lane-group tiles, spread
            # launches); split = 0 is what a full-size shard runs
            for split in (-1, 0):
                e.set_option("split", split)
                assert np.array_equal(e.search(q), want), (k, group_len, ql, split)
    finally:
        e.close()


def test_long_query_chunked_and_int32_recompute(swb, oracle):
    """query longer than one shared-memory chunk + scores far above 32767 (self hits of W-rich sequences)"""
    rng = np.random.default_rng(77)
    m = oracle.matrix("blosum50")
    long_q = rng.integers(0, 20, 9000).astype(np.uint8)
    wq = np.full(2600, 17, dtype=np.uint8)  # 'W' x 2600: self score 15 * 2600 = 39000
    enc = random_db(rng, rng.integers(20, 900, 150)) + [long_q.copy(), wq.copy(), long_q[:5000].copy(), wq[:2300].copy()]
    codes, offs = pack_db(enc)
    e = swb.Engine(0)
    try:
        e.db_load(codes, offs)
        for q in (long_q, wq):
            want = oracle.scan(q, codes, offs, m)
            for split in (-1, 0, 1):  # auto (on: small shard), the full-size-shard path, forced
                e.set_option("split", split)
                got = e.search(q)
                assert np.array_equal(got, want), split
                assert got.max() > 32767
                assert e.stats()["recomputed_tiles"] >= 1
        e.set_option("split", -1)
        e.set_option("chunk_rows", 1024)
        for k in (8, 16, 32):
            e.set_option("k", k)
            assert np.array_equal(e.search(long_q[:4100]), oracle.scan(long_q[:4100], codes, offs, m)), k
    finally:
        e.close()


def test_edge_cases(swb, oracle):
    m = oracle.matrix("blosum50")
    e = swb.Engine(0)
    try:
        q = swb.encode("MKVLAAGIW")
        # empty database, database of empty sequences, empty query, one residue
        c, o = pack_db([])
        e.db_load(c, o)
        assert e.db_count() == 0 and len(e.search(q)) == 0
        c, o = pack_db([np.zeros(0, np.uint8)] * 5)
        e.db_load(c, o)
        assert np.array_equal(e.search(q), np.zeros(5, np.int32))
        c, o = pack_db([swb.encode("W"), swb.encode("MKVLAAGIW"), np.zeros(0, np.uint8)])
        e.db_load(c, o)
        assert np.array_equal(e.search(np.zeros(0, np.uint8)), np.zeros(3, np.int32))
        assert np.array_equal(e.search(q), oracle.scan(q, c, o, m))
        assert np.array_equal(e.search(swb.encode("W")), np.array([15, 15, 0], np.int32))
        # unknown characters behave like '*' (score 0) and padding is neutral (FASTAParsers.h:94-96)
        a = e.search(swb.encode("MKVLAAGIW////"))
        assert np.array_equal(a, e.search(q))
    finally:
        e.close()


def test_batch_equals_single_and_resident_fetch(engine, swb, oracle, subset, queries):
    engine.set_scoring_preset(swb.SWB_SCORING_BLOSUM50_REF)
    engine.db_load(subset["codes"], subset["offsets"])
    names = ["P02232", "Q9UKN1", "P01008", "P27895", "P05013"]
    qs = [swb.encode(queries[n]) for n in names]
    batch = engine.search_batch(qs)
    assert engine.search_batch(qs, fetch=False) is None
    st = engine.stats()
    assert st["kernel_launches"] >= 3 * len(qs) and st["device_ms"] > 0
    for i in range(len(qs)):
        assert np.array_equal(engine.fetch_scores(i), batch[i])
    for i, q in enumerate(qs):
        assert np.array_equal(engine.search(q), batch[i])
    # the batch runs its longest query first; results stay in the caller's order either way, also with more queries
    # than job slots (slots are reused as jobs finish)
    many = [qs[i % len(qs)][: 40 + 37 * i] for i in range(40)]
    engine.set_option("batch_order", 1)
    as_given = engine.search_batch(many)
    engine.set_option("batch_order", 0)
    longest_first = engine.search_batch(many)
    m = oracle.matrix("blosum50")
    for i, q in enumerate(many):
        assert np.array_equal(as_given[i], longest_first[i]), i
        if i % 7 == 0:
            assert np.array_equal(as_given[i], oracle.scan(q, subset["codes"], subset["offsets"], m)), i


def test_shards_partition_and_agree(swb, oracle):
    """config 3 on one GPU: 1, 2, 4, 8 shards must reproduce the unsharded score vector exactly"""
    rng = np.random.default_rng(3)
    enc = random_db(rng, np.clip(np.round(rng.lognormal(5.0, 0.8, 700)), 2, 4000))
    codes, offs = pack_db(enc)
    q = rng.integers(0, 20, 300).astype(np.uint8)
    want = oracle.scan(q, codes, offs, oracle.matrix("blosum50"))
    e = swb.Engine(0)
    try:
        for nshards in (1, 2, 4, 8):
            merged = np.full(len(enc), -1, np.int32)
            res = []
            for s in range(nshards):
                e.db_load(codes, offs, s, nshards)
                ids = e.db_ids()
                merged[ids] = e.search(q)
                res.append(e.stats()["db_residues"])
            assert np.array_equal(merged, want), nshards
            assert max(res) - min(res) <= 2 * 4000
        ids, top = e.topk(e.search(q), 5)
    finally:
        e.close()


def test_dropin_result_order(swb, oracle, subset, queries, tmp_path):
    """smith_waterman_cuda mirror: (id, score) pairs in descending padded length, file order inside a bucket
    (SWSolver.cu:383-390); first ids 56, 34, 13 per SURVEY appendix A"""
    db = swb.FASTADatabase(os.path.join(GOLDEN, "uniprot_subset.fasta"))
    query = swb.FASTAQuery(os.path.join(GOLDEN, "queries", "P01008.fasta"))
    result = []
    swb.smith_waterman_cuda(query, db, result)
    assert [r[0] for r in result[:3]] == [56, 34, 13]
    gold = _gold("P01008")
    assert len(result) == 111
    for sid, score in result:
        assert score == gold[sid]


def test_reference_cuda_binary_agrees(swb, subset, queries, tmp_path):
    """the reference's own SWSolver.cu, compiled unmodified for sm_100a (oracle/_ref/ref_cuda_scan), on the same
    query/database files: identical id:score lines (queries within its 1024-row limit only, SWSolver.cu:85)"""
    binary = os.path.join(ROOT, "oracle", "_ref", "ref_cuda_scan")
    if not os.path.exists(binary):
        pytest.skip("oracle/_ref/ref_cuda_scan not built")
    dbpath = os.path.join(GOLDEN, "uniprot_subset.fasta")
    db = swb.FASTADatabase(dbpath)
    for name in ("P02232", "P01008", "P27895"):
        qpath = os.path.join(GOLDEN, "queries", name + ".fasta")
        out = subprocess.run([binary, qpath, dbpath], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        ref_lines = [l for l in out.stdout.split("\n") if l and not l.startswith("#")]
        result = []
        swb.smith_waterman_cuda(swb.FASTAQuery(qpath), db, result)
        assert ref_lines == ["%d:%d" % r for r in result], name


def test_cli_and_reference_style_caller(swb, tmp_path):
    """bin/main prints the reference's text (main.cpp:45-47, 58-60, 65-72); a caller written against the reference
    headers links against libswb.so and reproduces the golden file"""
    import re
    pkg = os.path.join(ROOT, "ece1782-smith-waterman-cuda_b200")
    main = os.path.join(pkg, "bin", "main")
    qpath = os.path.join(GOLDEN, "queries", "P01008.fasta")
    dbpath = os.path.join(GOLDEN, "uniprot_subset.fasta")
    r = subprocess.run([main, "--query", qpath, "--db=" + dbpath], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")
    query = swb.FASTAQuery(qpath).get_buffer()
    assert lines[0] == "Input buffer:" + query and lines[1] == ""
    gold = _gold("P01008")
    body = lines[2:2 + 111]
    assert [int(l.split(":")[0]) for l in body[:3]] == [56, 34, 13]
    for l in body:
        sid, sc = l.split(":")
        assert int(sc) == gold[int(sid)]
    tail = lines[2 + 111:]
    assert tail[0] == "=" * 80 and tail[1] == "METRICS:"
    assert tail[2] == "Query length: 464 chars." and tail[3] == "Num subjects: 111"
    assert tail[4] == "Sum of DB length: 26728 chars."
    assert re.fullmatch(r"Time elapsed: [0-9.e+-]+ seconds\.", tail[5])
    assert re.fullmatch(r"Performance: [0-9.e+-]+ GCUPS\.", tail[6])
    # reference-style caller
    exe = str(tmp_path / "caller")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "caller_compat.cpp"), os.path.join(pkg, "lib", "SWSolver.o"),
                    "-L" + os.path.join(pkg, "lib"), "-lswb", "-Wl,-rpath," + os.path.join(pkg, "lib")], check=True)
    r = subprocess.run([exe, qpath, dbpath, os.path.join(GOLDEN, "P01008.head111.txt")], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches 0" in r.stdout


def test_config4_titin_scale_queries_and_targets(swb, oracle):
    """BASELINE configs[3]: queries of 7,000 / 20,000 / 35,213 residues against long targets, a 10 %-mutated copy and
    the query itself (self score ~5.5 x L >> 32767: either the s16 pass flags it and the exact pass re-scores it, or
    the tile goes to the rebased s16 policy at once). 16- and 32-lane wavefront tiles carry the targets."""
    rng = np.random.default_rng(1784)
    m = oracle.matrix("blosum50")
    targets = random_db(rng, np.round(np.exp(rng.uniform(np.log(5000), np.log(35213), 20))), alphabet=20)
    e = swb.Engine(0)
    try:
        for qlen in (7000, 20000, 35213):
            q = rng.integers(0, 20, qlen).astype(np.uint8)
            mutated = q.copy()
            pos = rng.choice(qlen, qlen // 10, replace=False)
            mutated[pos] = rng.integers(0, 20, len(pos))
            enc = targets + [mutated, q.copy(), rng.integers(0, 20, 300).astype(np.uint8)]
            codes, offs = pack_db(enc)
            e.db_load(codes, offs)
            got = e.search(q)
            want = oracle.scan(q, codes, offs, m)
            assert np.array_equal(got, want), qlen
            assert got[len(targets) + 1] > 5 * qlen and 5 * qlen > 32767  # the self hit
            st = e.stats()
            assert st["tiles_by_group"][5] >= 1
            # below direct_len (10,000 rows) the self hit is flagged by the s16 pass and re-scored; above it the long
            # tiles are scored by the rebased policy at once and nothing may be left to flag
            assert st["recomputed_tiles"] >= 1 or qlen >= 10000
    finally:
        e.close()


def test_config5_many_queries_sampled_parity_and_shard_checksum(swb, oracle):
    """BASELINE configs[4], scaled to one GPU: a Swiss-Prot-shaped database (bench.synth_db, 0.25 scale) against 120
    queries drawn from the same length law; parity on a 1/256 stride sample of the database for 12 queries, and the
    size-independent property that the checksum of all scores is the same for 1 shard and for 4 shards."""
    import bench
    codes, offs = bench.synth_db(scale=0.25)
    n = len(offs) - 1
    rng = np.random.default_rng(1785)
    qlens = np.clip(np.round(rng.lognormal(5.58, 0.75, 120)), 30, 5478).astype(int)
    queries = [rng.integers(0, 20, l).astype(np.uint8) for l in qlens]
    m = oracle.matrix("blosum50")
    e = swb.Engine(0)
    try:
        e.db_load(codes, offs)
        full = e.search_batch(queries)
        assert full.shape == (120, n) and (full >= 0).all()
        for qi in range(0, 120, 10):
            want = oracle.scan(queries[qi], codes, offs, m, start=3, stride=256)
            sel = want >= 0
            assert sel.sum() >= n // 256
            assert np.array_equal(full[qi][sel], want[sel]), qi
        checksum = full.astype(np.int64).sum(axis=1)
        acc = np.zeros(120, dtype=np.int64)
        seen = 0
        for s in range(4):
            e.db_load(codes, offs, s, 4)
            part = e.search_batch(queries)
            acc += part.astype(np.int64).sum(axis=1)
            ids = e.db_ids()
            assert np.array_equal(part[:, :50], full[:, ids[:50]])
            seen += len(ids)
        assert seen == n and np.array_equal(acc, checksum)
    finally:
        e.close()


def test_traceback_alignment_matches_cpu_cpp_and_oracle(swb, oracle, subset, queries):
    """swb_align (GPU traceback of a hit) against the aligned strings the compiled reference cpu.cpp printed (+3/-3
    scheme, tests/golden/cpu_ref_ident3.json) and against the oracle's restatement of cpu.cpp:39-103 under BLOSUM50"""
    import json
    ref = json.load(open(os.path.join(GOLDEN, "cpu_ref_ident3.json")))
    e = swb.Engine(0)
    try:
        e.set_scoring_preset(swb.SWB_SCORING_IDENT3)
        for p in ref["pairs"]:
            c, o = swb.pack_sequences([swb.encode(p["b"], swb.SWB_SCORING_IDENT3)])
            e.db_load(c, o)
            score, ei, ej, ops = e.align(swb.encode(p["a"], swb.SWB_SCORING_IDENT3), 0, len(p["b"]))
            assert score == p["score"]
            assert swb.render_alignment(p["a"], p["b"], ei, ej, ops) == (p["aligned_a"], p["aligned_b"])
        e.set_scoring_preset(swb.SWB_SCORING_BLOSUM50_REF)
        e.db_load(subset["codes"], subset["offsets"])
        for name in ("P02232", "P01008"):
            q = swb.encode(queries[name])
            scores = e.search(q)
            ids, top = e.topk(scores, 4)
            for sid, sc in zip(ids, top):
                subj = subset["seqs"][int(sid)]
                score, ei, ej, ops = e.align(q, int(sid), len(subj))
                want = oracle.align(queries[name], subj, "blosum50")
                assert score == sc == want[0]
                assert (ei, ej) == want[3]
                assert swb.render_alignment(queries[name], subj, ei, ej, ops) == (want[1], want[2])
        assert e.align(np.zeros(0, np.uint8), 3, 10)[0] == 0
    finally:
        e.close()


def test_pack_time_reruns_the_pack_kernel_without_side_effects(swb, oracle, subset, queries):
    """swb_pack_time (measurement support of bench.py's roofline.hbm_pack): re-running the pack kernel leaves the
    resident database as it was"""
    e = swb.Engine(0)
    try:
        with pytest.raises(swb.SwbError):
            e.pack_time(3)
        e.db_load(subset["codes"], subset["offsets"])
        q = swb.encode(queries["P01008"])
        before = e.search(q)
        us, nbytes = e.pack_time(3)
        assert us > 0 and nbytes >= 2 * int(subset["offsets"][-1])
        assert np.array_equal(e.search(q), before) and np.array_equal(before, _gold("P01008"))
    finally:
        e.close()


def test_published_textbook_vector_on_gpu(swb):
    """Durbin et al. 1998, fig. 2.6: HEAGAWGHEE x PAWHEAE under BLOSUM50 with gap 8 -> 28, AWGHE / AW-HE -- through the scan,
    the traceback and the affine kernels with open == extend == 8 (an external known answer, see tests/test_oracle.py)"""
    e = swb.Engine(0)
    try:
        m = swb.scoring_matrix(swb.SWB_SCORING_BLOSUM50_REF)[0]
        q, d = swb.encode("HEAGAWGHEE"), swb.encode("PAWHEAE")
        c, o = swb.pack_sequences([d, q, d])
        e.set_scoring(m, 8)
        e.db_load(c, o)
        assert e.search(q).tolist()[0::2] == [28, 28]
        score, ei, ej, ops = e.align(q, 0, len(d))
        assert score == 28 and swb.render_alignment("HEAGAWGHEE", "PAWHEAE", ei, ej, ops) == ("AWGHE", "AW-HE")
        e.set_scoring_affine(m, 8, 8)
        e.db_load(c, o)
        assert e.search(q).tolist()[0::2] == [28, 28]
        e.set_scoring_affine(m, 12, 2)  # the same alignment with a one-residue gap at 12: 5 + 15 - 12 + 10 + 6
        e.db_load(c, o)
        assert e.search(q)[0] == 24
    finally:
        e.close()


def test_affine_traceback_matches_oracle(swb, oracle, subset, queries):
    """swb_align_batch under affine gaps (Gotoh's three states, 4 direction bits per cell) == the oracle's swo_align_affine
    -- scores, end cells and both aligned strings -- for the top hits of two queries under three gap models; open ==
    extend through the affine kernel path of the traceback is the linear walk the compiled cpu.cpp pins"""
    m = swb.scoring_matrix(swb.SWB_SCORING_BLOSUM50_REF)[0]
    e = swb.Engine(0)
    try:
        names = ("P02232", "P01008")
        qs = [swb.encode(queries[nm]) for nm in names]
        for go, ge in ((10, 2), (5, 1), (12, 0)):
            e.set_scoring_affine(m, go, ge)
            e.db_load(subset["codes"], subset["offsets"])
            ids, top = e.search_batch_topk(*swb.pack_sequences(qs), 5)
            hits = [(qi, int(sid)) for qi in range(len(qs)) for sid in ids[qi]] + [(0, 56), (1, 110)]
            got = e.align_batch(qs, hits, [len(subset["seqs"][sid]) for _, sid in hits])
            for (qi, sid), (score, ei, ej, ops) in zip(hits, got):
                subj = subset["seqs"][sid]
                want = oracle.align_affine(queries[names[qi]], subj, go, ge)
                assert score == want[0] and (ei, ej) == want[3], (go, ge, qi, sid)
                assert swb.render_alignment(queries[names[qi]], subj, ei, ej, ops) == (want[1], want[2]), (go, ge, qi, sid)
            assert [g[0] for g in got[:5]] == [int(v) for v in top[0]]
        # a gap that is extended: AAAA x AAGGAA with open 3 / extend 1
        c, o = swb.pack_sequences([swb.encode("AAGGAA")])
        e.set_scoring_affine(m, 3, 1)
        e.db_load(c, o)
        score, ei, ej, ops = e.align(swb.encode("AAAA"), 0, 6)
        assert score == 16 and swb.render_alignment("AAAA", "AAGGAA", ei, ej, ops) == ("AA--AA", "AAGGAA")
    finally:
        e.close()


def test_align_batch_matches_single_calls_and_oracle(swb, oracle, subset, queries):
    """swb_align_batch: the hit lists of several queries in ONE launch (one block per hit, H diagonals in shared memory,
    2-bit directions) == swb_align hit by hit == the oracle's restatement of cpu.cpp:39-103; empty hits, repeated hits
    and a 20,000-row query whose diagonals live in global scratch included"""
    e = swb.Engine(0)
    try:
        e.db_load(subset["codes"], subset["offsets"])
        names = ("P02232", "P01008", "P27895", "Q9UKN1")
        qs = [swb.encode(queries[nm]) for nm in names]
        ids, top = e.search_batch_topk(*swb.pack_sequences(qs), 6)
        hits = [(qi, int(sid)) for qi in range(len(qs)) for sid in ids[qi]]
        hits += [hits[3], (0, 110), (1, 0)]
        lens = [len(subset["seqs"][sid]) for _, sid in hits]
        got = e.align_batch(qs, hits, lens)
        assert len(got) == len(hits)
        for (qi, sid), (score, ei, ej, ops) in zip(hits, got):
            subj = subset["seqs"][sid]
            one = e.align(qs[qi], sid, len(subj))
            assert (score, ei, ej) == one[:3] and np.array_equal(ops, one[3])
            want = oracle.align(queries[names[qi]], subj, "blosum50")
            assert score == want[0] and (ei, ej) == want[3]
            assert swb.render_alignment(queries[names[qi]], subj, ei, ej, ops) == (want[1], want[2])
        for qi in range(len(qs)):
            assert [g[0] for g in got[6 * qi:6 * qi + 6]] == [int(v) for v in top[qi]]
        # empty list, empty query
        assert e.align_batch(qs, [], []) == []
        z = e.align_batch([np.zeros(0, np.uint8), qs[0]], [(0, 5), (1, 5)], [len(subset["seqs"][5])] * 2)
        assert z[0][0] == 0 and len(z[0][3]) == 0 and z[1][0] == e.align(qs[0], 5, len(subset["seqs"][5]))[0]
        # a query beyond the shared-memory limit of the diagonals (3 * (m + 2) ints > 216 KB)
        rng = np.random.default_rng(7)
        subj = subset["seqs"][56]
        letters = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
        long_txt = bytes(letters[rng.integers(0, 20, 20000)]).decode()
        long_txt = long_txt[:9000] + subj[100:700] + long_txt[9600:]
        lq = swb.encode(long_txt)
        (score, ei, ej, ops), = e.align_batch([lq], [(0, 56)], [len(subj)])
        want = oracle.align(long_txt, subj, "blosum50")
        assert score == want[0] and (ei, ej) == want[3]
        assert swb.render_alignment(long_txt, subj, ei, ej, ops) == (want[1], want[2])
        with pytest.raises(swb.SwbError):
            e.align_batch(qs, [(9, 0)], [10])
    finally:
        e.close()


def test_pipelined_passes_option(swb, oracle):
    """option split=1: 32-lane tiles wider than xl_len hand out their passes as pipelined work items (progress
    counters in global memory, atomicMax merge); same scores, including a chunked query and an s16 overflow"""
    rng = np.random.default_rng(99)
    m = oracle.matrix("blosum50")
    w = np.full(2400, 17, dtype=np.uint8)
    enc = random_db(rng, [9000, 8200, 5000, 4100, 3000, 2500, 700, 300, 120, 40], alphabet=20) + [w.copy()]
    codes, offs = pack_db(enc)
    e = swb.Engine(0, split=1, xl_len=2048)
    try:
        e.db_load(codes, offs)
        for q in (rng.integers(0, 20, 144).astype(np.uint8), rng.integers(0, 20, 1000).astype(np.uint8),
                  rng.integers(0, 20, 5478).astype(np.uint8), w):
            assert np.array_equal(e.search(q), oracle.scan(q, codes, offs, m)), len(q)
        e.set_option("chunk_rows", 1024)
        q = rng.integers(0, 20, 3000).astype(np.uint8)
        assert np.array_equal(e.search(q), oracle.scan(q, codes, offs, m))
        batch = e.search_batch([q[:500], q[:1500], q])
        assert np.array_equal(batch[2], oracle.scan(q, codes, offs, m))
    finally:
        e.close()


def _host_topk(scores, ids, k):
    """k best of a score vector: score descending, database id ascending (numpy restatement for the tests)"""
    order = np.lexsort((ids, -scores.astype(np.int64)))[:k]
    return ids[order], scores[order]


def test_device_topk_equals_host_selection(swb, oracle):
    """swb_search_batch_topk: the hit list selected on the device (radix select + sort in one block) must equal the
    host selection over the full vectors -- many equal scores (ties break by ascending database id), k from 1 to
    1024 and beyond the shard size, sharded ids, an empty query"""
    rng = np.random.default_rng(31)
    enc = random_db(rng, np.clip(np.round(rng.lognormal(4.2, 0.9, 5000)), 1, 3000), alphabet=6)  # small alphabet: ties
    codes, offs = pack_db(enc)
    qs = [rng.integers(0, 6, l).astype(np.uint8) for l in (5, 40, 333, 1200)] + [np.zeros(0, np.uint8)]
    qcodes, qoffs = swb.pack_sequences(qs)
    e = swb.Engine(0)
    try:
        for shard, nshards in ((0, 1), (1, 3)):
            e.db_load(codes, offs, shard, nshards)
            ids_all = e.db_ids()
            full = e.search_batch(qs)
            for k in (1, 10, 100, 1024):
                ids, top = e.search_batch_topk(qcodes, qoffs, k)
                for qi in range(len(qs)):
                    assert np.array_equal(e.fetch_scores(qi), full[qi])  # the full vectors stay resident
                    wi, wt = _host_topk(full[qi], ids_all, k)
                    assert np.array_equal(ids[qi], wi) and np.array_equal(top[qi], wt), (shard, k, qi)
                    hi, ht = e.topk(full[qi], k)  # the older host-side call agrees too
                    assert np.array_equal(hi, wi) and np.array_equal(ht, wt)
        # a shard smaller than k: the tail is (0xffffffff, -1)
        e.db_load(codes[:int(offs[7])], offs[:8])
        ids, top = e.search_batch_topk(qcodes, qoffs, 16)
        full = e.search_batch(qs)
        for qi in range(len(qs)):
            wi, wt = _host_topk(full[qi], np.arange(7, dtype=np.uint32), 16)
            assert np.array_equal(ids[qi][:7], wi) and np.array_equal(top[qi][:7], wt)
            assert (ids[qi][7:] == 0xFFFFFFFF).all() and (top[qi][7:] == -1).all()
        with pytest.raises(swb.SwbError):
            e.search_batch_topk(qcodes, qoffs, 2000)
    finally:
        e.close()


def test_scatter_by_database_id_and_stale_results(swb, oracle):
    """swb_search_batch_scatter: three engines (one per shard) fill one nq x n matrix by database id; results of a
    previous database must not be readable after a reload (swb_fetch_scores used to index the new shard with them)"""
    rng = np.random.default_rng(12)
    enc = random_db(rng, rng.integers(1, 900, 400), alphabet=20)
    codes, offs = pack_db(enc)
    qs = [rng.integers(0, 20, l).astype(np.uint8) for l in (64, 500)]
    qcodes, qoffs = swb.pack_sequences(qs)
    m = oracle.matrix("blosum50")
    out = np.full((2, len(enc)), -9, dtype=np.int32)
    engines = [swb.Engine(0) for _ in range(3)]
    try:
        for s, e in enumerate(engines):
            e.db_load(codes, offs, s, 3)
            e.search_batch_scatter(qcodes, qoffs, out)
        for qi, q in enumerate(qs):
            assert np.array_equal(out[qi], oracle.scan(q, codes, offs, m)), qi
        e = engines[0]
        e.search_batch(qs, fetch=False)
        assert e.fetch_scores(1) is not None
        e.db_load(codes, offs, 0, 2)  # a larger shard: the old result matrix no longer matches
        with pytest.raises(swb.SwbError):
            e.fetch_scores(0)
        e.search_batch(qs[:1], fetch=False)
        assert np.array_equal(e.fetch_scores(0), oracle.scan(qs[0], codes, offs, m)[e.db_ids()])
        with pytest.raises(swb.SwbError):
            e.fetch_scores(1)
    finally:
        for e in engines:
            e.close()


def test_exact_pass_policies_far_beyond_the_s16_range(swb, oracle):
    """True scores up to 151,500 (10,100 W-W matches at 15): the rebased s16 policy (V16R) as recompute of flagged
    tiles, as the direct pass of long-against-long tiles (option direct_len), pipelined with 8 and 16 rows per lane and
    not pipelined, and the int32 policy (option exact=1) must all reproduce the oracle"""
    rng = np.random.default_rng(2026)
    m = oracle.matrix("blosum50")
    w = np.full(10100, 17, dtype=np.uint8)
    noisy = w.copy()
    noisy[rng.choice(len(w), 700, replace=False)] = rng.integers(0, 20, 700)
    enc = random_db(rng, rng.integers(20, 2500, 200), alphabet=20) + [
        w.copy(), noisy, w[:9000].copy(), rng.integers(0, 20, 12000).astype(np.uint8), w[:2300].copy(),
        rng.integers(0, 20, 7000).astype(np.uint8)]
    codes, offs = pack_db(enc)
    queries = [w, rng.integers(0, 20, 9000).astype(np.uint8), noisy[:6000]]
    want = [oracle.scan(q, codes, offs, m) for q in queries]
    assert want[0].max() == 151500
    configs = [dict(), dict(exact=1), dict(split=1, split_k=8, direct_len=4000), dict(split=1, split_k=16, direct_len=4000),
               dict(split=1, split_k=16, direct_len=0), dict(split=0), dict(split=0, exact=1), dict(group_len=1536, split=1)]
    for cfg in configs:
        e = swb.Engine(0, **cfg)
        try:
            e.db_load(codes, offs)
            for q, wv in zip(queries, want):
                got = e.search(q)
                assert np.array_equal(got, wv), (cfg, len(q))
            batch = e.search_batch(queries)
            for qi in range(len(queries)):
                assert np.array_equal(batch[qi], want[qi]), (cfg, qi)
        finally:
            e.close()


def test_rebased_s16_padding_lanes_regression(swb, oracle):
    """Two configs[3] targets of ~21,400 residues against the 35,213-row query: true scores just above 32767 with a
    rebase block boundary inside the padding behind the last column -- the case in which padding lanes at the
    clamped floor once wrapped into the running maximum (scores came out ~32000 too high). All pipelined strip
    heights, direct and flagged-then-rescored."""
    import bench
    codes, offs, qs = bench.synth_config4()
    m = oracle.matrix("blosum50")
    ids = (68, 1, 149, 27, 79, 130)
    c2, o2 = pack_db([codes[int(offs[i]):int(offs[i + 1])] for i in ids])
    want = oracle.scan(qs[3], c2, o2, m)
    assert want.min() > 32767
    for cfg in (dict(), dict(split_k=8), dict(split_k=32), dict(direct_len=0), dict(direct_len=0, split_k=32)):
        e = swb.Engine(0, **cfg)
        try:
            e.db_load(c2, o2)
            assert np.array_equal(e.search(qs[3]), want), cfg
        finally:
            e.close()


def _group_devices(n):
    """n engines: real devices when the box has them, else several engines on device 0"""
    import torch
    have = torch.cuda.device_count()
    return list(range(n)) if have >= n else [i % max(have, 1) for i in range(n)]


@pytest.mark.parametrize("ndev,parts", [(2, 0), (4, 2), (4, 4), (8, 0), (8, 2), (3, 1)])
def test_engine_group_matches_single_engine(swb, oracle, ndev, parts):
    """swb_group_*: P database parts x R query groups in one process must return the matrix of a single engine, and
    the merged device-side hit lists the host selection over that matrix"""
    rng = np.random.default_rng(100 + ndev)
    enc = random_db(rng, np.clip(np.round(rng.lognormal(4.5, 0.9, 1500)), 0, 5000), alphabet=8)
    codes, offs = pack_db(enc)
    qs = [rng.integers(0, 8, l).astype(np.uint8) for l in (700, 33, 1500, 260, 90, 1100, 8, 410, 5, 0, 222)]
    qcodes, qoffs = swb.pack_sequences(qs)
    m = oracle.matrix("blosum50")
    g = swb.EngineGroup(_group_devices(ndev), min_part_sequences=300)
    try:
        if parts:
            g.set_option("db_parts", parts)
        g.db_load(codes, offs)
        want_parts = parts if parts else swb.layout_parts(len(enc), ndev, 300)
        assert g.db_parts() == want_parts and g.size() == ndev
        full = g.search_batch_packed(qcodes, qoffs)
        for qi in (0, 2, 4, 9):
            assert np.array_equal(full[qi], oracle.scan(qs[qi], codes, offs, m)), qi
        single = swb.Engine(0)
        try:
            single.db_load(codes, offs)
            assert np.array_equal(single.search_batch(qs), full)
        finally:
            single.close()
        ids, top = g.search_batch_topk(qcodes, qoffs, 25)
        all_ids = np.arange(len(enc), dtype=np.uint32)
        for qi in range(len(qs)):
            wi, wt = _host_topk(full[qi], all_ids, 25)
            assert np.array_equal(ids[qi], wi) and np.array_equal(top[qi], wt), qi
        st = g.stats()
        assert st["db_sequences"] == len(enc) and st["cells"] == sum(len(q) for q in qs) * int(offs[-1])
        # lone queries and a reload
        one = g.search_batch([qs[2]])
        assert np.array_equal(one[0], full[2])
        g.db_load(codes[:int(offs[40])], offs[:41])
        assert np.array_equal(g.search_batch([qs[0]])[0], oracle.scan(qs[0], codes[:int(offs[40])], offs[:41], m))
    finally:
        g.close()


@pytest.mark.parametrize("devices", ["0", "0,0", "0,0,0,0,0,0,0,0"])
def test_dropin_on_several_devices(swb, subset, tmp_path, devices):
    """the C++ drop-in (smith_waterman_cuda, bin/main) shards the database over every device of its engine group:
    with 1, 2 and 8 engines the reference-style caller reproduces the golden file and bin/main prints the same
    id:score lines (SWB_DEVICES names the devices; on a box with fewer GPUs the indices repeat)"""
    import torch
    have = torch.cuda.device_count()
    n = len(devices.split(","))
    env = dict(os.environ)
    env["SWB_DEVICES"] = ",".join(str(i) for i in range(n)) if have >= n else devices
    pkg = os.path.join(ROOT, "ece1782-smith-waterman-cuda_b200")
    exe = str(tmp_path / "caller")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "caller_compat.cpp"), os.path.join(pkg, "lib", "SWSolver.o"),
                    "-L" + os.path.join(pkg, "lib"), "-lswb", "-Wl,-rpath," + os.path.join(pkg, "lib")], check=True)
    qpath = os.path.join(GOLDEN, "queries", "P01008.fasta")
    dbpath = os.path.join(GOLDEN, "uniprot_subset.fasta")
    r = subprocess.run([exe, qpath, dbpath, os.path.join(GOLDEN, "P01008.head111.txt")], capture_output=True, text=True,
                       timeout=300, env=env)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout + r.stderr
    main = os.path.join(pkg, "bin", "main")
    r = subprocess.run([main, "--query", qpath, "--db", dbpath], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    body = [l for l in r.stdout.split("\n")[2:2 + 111]]
    gold = _gold("P01008")
    assert [int(l.split(":")[0]) for l in body[:3]] == [56, 34, 13]
    for l in body:
        sid, sc = l.split(":")
        assert int(sc) == gold[int(sid)]
    if n == 1:  # --gpus is an extension that must not change the output
        r2 = subprocess.run([main, "--query", qpath, "--db", dbpath, "--gpus", "1"], capture_output=True, text=True,
                            timeout=300)
        assert r2.returncode == 0 and r2.stdout.split("\n")[:113] == r.stdout.split("\n")[:113]


def test_cli_missing_files_and_error_codes(swb, tmp_path):
    """reference behaviour Q9: a nonexistent database is one empty record with id -1, a nonexistent query an
    empty buffer -> every score is 0, exit code 0; C ABI argument errors come back as codes with a message"""
    main = os.path.join(ROOT, "ece1782-smith-waterman-cuda_b200", "bin", "main")
    r = subprocess.run([main, "--query", "/nonexistent/q.fasta", "--db", "/nonexistent/db.fasta"], capture_output=True,
                       text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")
    assert lines[0] == "Input buffer:" and lines[2] == "-1:0" and lines[3] == "=" * 80
    assert "Query length: 0 chars." in r.stdout and "Num subjects: 1" in r.stdout and "Sum of DB length: 0 chars." in r.stdout
    e = swb.Engine(0)
    try:
        with pytest.raises(swb.SwbError):
            e.search(np.zeros(5, np.uint8))          # no database yet
        with pytest.raises(swb.SwbError):
            e.set_option("k", 7)
        with pytest.raises(swb.SwbError):
            e.set_scoring(np.full((4, 4), 127, np.int8), 2)   # S + gap does not fit int8
        with pytest.raises(swb.SwbError):
            e.db_load(np.zeros(4, np.uint8), np.array([0, 3, 2], np.uint64))  # decreasing offsets
        e.db_load(np.zeros(4, np.uint8), np.array([0, 3, 4], np.uint64))
        with pytest.raises(swb.SwbError):
            e.align(np.zeros(3, np.uint8), 7, 3)     # id not in the database
        assert e.search(np.zeros(3, np.uint8)).tolist() == [15, 5]   # AAA vs AAA / A, BLOSUM50 A-A = 5
    finally:
        e.close()


def test_cli_encoded_database_matches_text_database(swb, tmp_path):
    """bin/main --db <file>.swbdb (written by bin/swb_mkdb) prints the same ids, scores, order and METRICS counts as
    the same database given as FASTA text"""
    pkg = os.path.join(ROOT, "ece1782-smith-waterman-cuda_b200")
    main, mkdb = os.path.join(pkg, "bin", "main"), os.path.join(pkg, "bin", "swb_mkdb")
    fasta = os.path.join(GOLDEN, "uniprot_subset.fasta")
    enc = str(tmp_path / "subset.swbdb")
    assert subprocess.run([mkdb, fasta, enc], capture_output=True).returncode == 0
    q = os.path.join(GOLDEN, "queries", "P02232.fasta")
    outs = []
    for db in (fasta, enc):
        r = subprocess.run([main, "--query", q, "--db", db], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append([l for l in r.stdout.split("\n") if not l.startswith(("Time elapsed", "Performance"))])
    assert outs[0] == outs[1] and len(outs[0]) > 115
    # header-less text file: one record with id -1 in both paths (FASTAParsers.h:82)
    enc2 = str(tmp_path / "test.swbdb")
    assert subprocess.run([mkdb, os.path.join(GOLDEN, "test.dat"), enc2], capture_output=True).returncode == 0
    outs = []
    for db in (os.path.join(GOLDEN, "test.dat"), enc2):
        r = subprocess.run([main, "--query", q, "--db", db], capture_output=True, text=True, timeout=300)
        outs.append([l for l in r.stdout.split("\n") if not l.startswith(("Time elapsed", "Performance"))])
    assert outs[0] == outs[1] and outs[0][2].startswith("-1:")


def test_cpp_shim_two_databases_in_one_process(swb, subset, tmp_path):
    """the C++ drop-in caches the packed database per FASTADatabase content: alternating between two databases (the
    second = the first with its records reversed) must give each its own scores"""
    pkg = os.path.join(ROOT, "ece1782-smith-waterman-cuda_b200")
    rev = str(tmp_path / "reversed.fasta")
    with open(rev, "w") as f:
        for k in range(len(subset["seqs"]) - 1, -1, -1):
            f.write(">r%d\n%s\n" % (k, subset["seqs"][k]))
    exe = str(tmp_path / "two_dbs")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "two_databases.cpp"), os.path.join(pkg, "lib", "SWSolver.o"),
                    "-L" + os.path.join(pkg, "lib"), "-lswb", "-Wl,-rpath," + os.path.join(pkg, "lib")], check=True)
    r = subprocess.run([exe, os.path.join(GOLDEN, "queries", "P01008.fasta"), os.path.join(GOLDEN, "uniprot_subset.fasta"),
                        rev, os.path.join(GOLDEN, "P01008.head111.txt")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("group_len,k", [(384, 0), (16, 8), (96, 16), (100000, 0)])
def test_affine_gaps_vs_gotoh_oracle(swb, oracle, subset, queries, group_len, k):
    """the affine model the reference left as a comment (SWSolver.cu:8). Kernels V16A / V32A against the
    Gotoh restatement in oracle/sw_oracle.c; gap_open == gap_extend must reproduce the reference's linear goldens."""
    rng = np.random.default_rng(4242 + group_len)
    m = oracle.matrix("blosum50")
    lens = np.concatenate([rng.integers(0, 40, 50), rng.integers(40, 700, 300), rng.integers(700, 3000, 9), [0, 1, 2]])
    rng.shuffle(lens)
    wq = np.full(2400, 17, dtype=np.uint8)  # self score 36000: s16 overflow -> V32A recompute
    enc = random_db(rng, lens) + [wq.copy(), wq[:2250].copy()]
    codes, offs = pack_db(enc)
    e = swb.Engine(0, group_len=group_len, k=k)
    try:
        e.db_load(codes, offs)
        for go, ge in ((10, 2), (12, 1), (3, 0)):
            e.set_scoring_affine(m, go, ge)
            for ql in (1, 9, 33, 257, 1000):
                q = rng.integers(0, 6 if ql < 100 else 24, ql).astype(np.uint8)
                assert np.array_equal(e.search(q), oracle.scan_affine(q, codes, offs, m, go, ge)), (go, ge, ql)
            got = e.search(wq)
            assert np.array_equal(got, oracle.scan_affine(wq, codes, offs, m, go, ge))
            assert got.max() == 36000 and e.stats()["recomputed_tiles"] >= 1
        # batches and a query longer than one profile chunk
        e.set_scoring_affine(m, 10, 2)
        e.set_option("chunk_rows", 1024)
        qs = [rng.integers(0, 24, n).astype(np.uint8) for n in (300, 2100, 50, 1025)]
        batch = e.search_batch(qs)
        for q, got in zip(qs, batch):
            assert np.array_equal(got, oracle.scan_affine(q, codes, offs, m, 10, 2)), len(q)
        # traceback under the affine model: the score of the walk is the scan's score
        assert e.align(qs[0], 0, int(offs[1] - offs[0]))[0] == batch[0][0]
        # go == ge: the linear kernels, the reference's goldens
        e.set_scoring_affine(m, 2, 2)
        e.set_option("chunk_rows", 7168)
        e.db_load(subset["codes"], subset["offsets"])
        for name in ("P01008", "P02232"):
            assert np.array_equal(e.search(swb.encode(queries[name])), _gold(name)), name
    finally:
        e.close()
