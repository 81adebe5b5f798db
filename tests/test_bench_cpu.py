"""CPU-side checks of bench.py's workload generators and of the sampled parity helper (no GPU)."""
import os
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_config2_generator_is_the_documented_database():
    """SURVEY 8(d) config 2, seed 1782: the sizes DESIGN.md and the bench line quote"""
    codes, offsets = bench.synth_db()
    assert len(offsets) - 1 == 570065 and int(offsets[-1]) == 200838771
    lens = np.diff(offsets.astype(np.int64))
    assert lens.max() == 35213 and lens.min() >= 2
    assert codes.dtype == np.uint8 and codes.max() < 32
    # record order is shuffled against length (ids != length order)
    assert not np.all(np.diff(lens[:2000]) <= 0)
    c2, o2 = bench.synth_db()
    assert np.array_equal(codes[:100000], c2[:100000]) and np.array_equal(offsets, o2)


def test_config4_and_config5_generators():
    codes, offsets, qs = bench.synth_config4()
    assert [len(q) for q in qs] == [5000, 10000, 20000, 35213] and len(offsets) - 1 == 264
    lens = np.diff(offsets.astype(np.int64))
    assert lens.min() >= 5000 and lens.max() == 35213
    # the query itself and its 10 %-mutated copy are in the database
    k = 256
    for q in qs:
        mutated = codes[int(offsets[k]):int(offsets[k + 1])]
        same = codes[int(offsets[k + 1]):int(offsets[k + 2])]
        assert np.array_equal(same, q) and len(mutated) == len(q)
        assert 0.05 < float(np.mean(mutated != q)) < 0.1
        k += 2
    q5 = bench.synth_queries(50)
    assert len(q5) == 50 and min(len(q) for q in q5) >= 30 and max(len(q) for q in q5) <= 5478


def test_cpu_sample_scores_line_up_with_full_scans(oracle):
    """bench.py compares the oracle's stride sample with the GPU's full score vectors: the sampled entries must be the
    entries 0, stride, 2*stride, ... of a full scan"""
    codes, offsets = bench.synth_db(scale=0.004)
    names, qtexts = bench.load_queries(None)
    kept = []
    cb, dt = bench.cpu_sample_gcups(codes, offsets, names, qtexts[:2], 0.02, keep=kept)
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "every" in cb["sample"]
    m = oracle.matrix("blosum50")
    for (stride, want), q in zip(kept, qtexts[:2]):
        full = oracle.scan(oracle.encode(q), codes, offsets, m)
        assert stride >= 1 and np.array_equal(full[0::stride], want)


def test_workload_config_names_the_workload():
    import argparse
    codes, offsets = bench.synth_db(scale=0.002)
    for wl in ("config2", "config4", "config5"):
        cfg = bench.workload_config(offsets, [np.zeros(5, np.uint8)], argparse.Namespace(workload=wl, scale=1.0, affine=""))
        assert "workload" in cfg and "model" not in cfg and cfg["db_sequences"] == len(offsets) - 1


def test_reference_arm_uses_every_core_under_torchrun(monkeypatch):
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs must not inherit that (round-1 SCALE lines compared against a
    one-core reference arm)"""
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    assert bench.host_threads() == len(os.sched_getaffinity(0))
    codes, offsets = bench.synth_db(scale=0.002)
    names, qtexts = bench.load_queries(None)
    cb, _ = bench.cpu_sample_gcups(codes, offsets, names, qtexts[:1], 0.01)
    assert cb["cores"] == bench.host_threads() and ("%d threads" % cb["cores"]) in cb["sample"]


class _FakeEngine:
    """stands in for the CUDA engine in the host-side parity helper: scores come from the oracle, one entry is
    optionally corrupted"""

    def __init__(self, oracle, codes, offsets, ids, queries, corrupt=False):
        self.ids = ids
        m = oracle.matrix("blosum50")
        self.rows = [oracle.scan(q, codes, offsets, m)[ids] for q in queries]
        if corrupt:
            self.rows[-1][0] += 1

    def db_ids(self):
        return self.ids


def test_per_rank_sample_parity_helper(oracle):
    import importlib
    swb = importlib.import_module("ece1782-smith-waterman-cuda_b200")
    codes, offsets = bench.synth_db(scale=0.003)
    n = len(offsets) - 1
    _, _, ids = swb.plan_describe(offsets, 1, 2, want_ids=True)
    rng = np.random.default_rng(0)
    qs = {3: rng.integers(0, 20, 50).astype(np.uint8), 7: rng.integers(0, 20, 120).astype(np.uint8)}
    local_of = {3: 0, 7: 1}
    for corrupt in (False, True):
        eng = _FakeEngine(oracle, codes, offsets, ids, list(qs.values()), corrupt)
        ok, desc = bench.sample_parity(swb, eng, codes, offsets, qs, local_of, lambda k: eng.rows[k], 2, 0.0001)
        assert ok == (not corrupt) and "every" in desc
    assert len(ids) < n
