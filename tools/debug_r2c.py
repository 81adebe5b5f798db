#!/usr/bin/env python3
"""GPU debugging helper of round 2: (perf) lane-group tiles alone and a half-size benchmark database, for whichever
library SWB_LIB names; (exact) pipelined long-sequence options against the oracle on a cut of the configs[3] database."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

swb = importlib.import_module(bench.PKG)


def timed(eng, qcodes, qoffs, reps=3):
    ms = []
    for _ in range(reps + 1):
        eng.search_batch_packed(qcodes, qoffs, fetch=False)
        ms.append(eng.stats()["device_ms"])
    return min(ms[1:])


def perf():
    rng = np.random.default_rng(7)
    print("library:", swb.LIB_PATH, flush=True)
    for lo, hi, gl in ((1600, 3000, 768), (1600, 3000, 1536), (800, 1500, 384), (200, 380, 96)):
        seqs = [rng.integers(0, 20, int(l)).astype(np.uint8) for l in rng.integers(lo, hi, 4000)]
        codes, offs = swb.pack_sequences(seqs)
        for ql in (1000, 5478):
            q = [rng.integers(0, 20, ql).astype(np.uint8)]
            qc, qo = swb.pack_sequences(q)
            eng = swb.Engine(0, group_len=gl, split=0)
            try:
                eng.db_load(codes, offs)
                t = timed(eng, qc, qo)
                st = eng.stats()
                print("lane-group tiles only: lengths %d..%d group_len %d query %d: %.2f ms %.0f GCUPS tiles %s k %d" % (
                    lo, hi, gl, ql, t, ql * float(offs[-1]) / t * 1e-6, st["tiles_by_group"], st["last_k"]), flush=True)
            finally:
                eng.close()
    codes, offs = bench.synth_db(scale=0.5)
    qs = bench.load_queries(swb)[1]
    qc, qo = swb.pack_sequences(qs)
    for gl in (768, 100000):
        eng = swb.Engine(0, group_len=gl)
        try:
            eng.db_load(codes, offs)
            t = timed(eng, qc, qo)
            print("config2 x 0.5, group_len %d: %.1f ms %.0f GCUPS tiles %s" % (
                gl, t, sum(len(q) for q in qs) * float(offs[-1]) / t * 1e-6, eng.stats()["tiles_by_group"]), flush=True)
        finally:
            eng.close()


def perf1():
    """one lane-group-only launch (for ncu)"""
    rng = np.random.default_rng(7)
    seqs = [rng.integers(0, 20, int(l)).astype(np.uint8) for l in rng.integers(1600, 3000, 4000)]
    codes, offs = swb.pack_sequences(seqs)
    q = [rng.integers(0, 20, 1000).astype(np.uint8)]
    qc, qo = swb.pack_sequences(q)
    eng = swb.Engine(0, group_len=1536, split=0)
    eng.db_load(codes, offs)
    print("%.2f ms" % timed(eng, qc, qo, 1))
    eng.close()


def exact_full():
    """every target of configs[3] x the 5,000- and 35,213-row queries, as a BATCH of all four queries"""
    from oracle_lib import Oracle
    o = Oracle()
    m = o.matrix("blosum50")
    codes, offs, qs = bench.synth_config4()
    lens = np.diff(offs.astype(np.int64))
    want = {qi: o.scan(qs[qi], codes, offs, m, 2, 0, 1, bench.host_threads()) for qi in (0, 3)}
    print("oracle done", flush=True)
    for spec in ("", "split_k=8", "split_k=32", "direct_len=0,exact=1"):
        opts = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in spec.split(",") if kv)
        eng = swb.Engine(0, **opts)
        try:
            eng.db_load(codes, offs)
            for mode in ("batch", "lone"):
                got = eng.search_batch(qs) if mode == "batch" else None
                for qi in (0, 3):
                    g = got[qi] if got is not None else eng.search(qs[qi])
                    bad = np.nonzero(g != want[qi])[0]
                    print("  {%s} %s query %d: %s" % (spec, mode, len(qs[qi]), "ok" if len(bad) == 0 else
                          "MISMATCH at %d sequences: %s" % (len(bad), [(int(i), int(lens[i]), int(g[i]), int(want[qi][i]))
                                                                         for i in bad[:8]])), flush=True)
        finally:
            eng.close()


def exact():
    from oracle_lib import Oracle
    o = Oracle()
    m = o.matrix("blosum50")
    codes, offs, qs = bench.synth_config4()
    keep = list(range(0, 256, 6)) + list(range(256, 264))
    seqs = [codes[int(offs[i]):int(offs[i + 1])] for i in keep]
    c2, o2 = swb.pack_sequences(seqs)
    lens = np.diff(o2.astype(np.int64))
    for q in (qs[0], qs[3]):
        t0 = time.time()
        want = o.scan(q, c2, o2, m, 2, 0, 1, bench.host_threads())
        print("query %d: oracle %.1f s, max score %d" % (len(q), time.time() - t0, want.max()), flush=True)
        for spec in ("", "split_k=8", "split_k=16", "split_k=32", "direct_len=0,split_k=8", "direct_len=0,split_k=32",
                     "direct_len=0,exact=1", "split=0", "split=0,exact=1"):
            opts = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in spec.split(",") if kv)
            eng = swb.Engine(0, **opts)
            try:
                eng.db_load(c2, o2)
                got = eng.search(q)
                bad = np.nonzero(got != want)[0]
                print("  {%s}: %s" % (spec, "ok" if len(bad) == 0 else "MISMATCH at %d sequences: %s" % (
                    len(bad), [(int(i), int(lens[i]), int(got[i]), int(want[i])) for i in bad[:6]])), flush=True)
            finally:
                eng.close()


if __name__ == "__main__":
    {"perf": perf, "exact": exact, "perf1": perf1, "exact_full": exact_full}[sys.argv[1]]()
