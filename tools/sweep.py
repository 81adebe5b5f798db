#!/usr/bin/env python3
"""One process, many engine configurations: device-timed GCUPS of a query batch per (database scale, option set).
usage: sweep.py WORKLOAD SCALE[,SCALE...] "opt=val,opt=val" ["opt=val,..." ...]     (an empty string = defaults)
WORKLOAD: config2 (20 reference queries), config4 (long sequences; SCALE ignored), short (150 synthetic queries)"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def main():
    swb = importlib.import_module("ece1782-smith-waterman-cuda_b200")
    workload = sys.argv[1]
    scales = [float(x) for x in sys.argv[2].split(",")]
    optsets = sys.argv[3:] or [""]
    reps = int(os.environ.get("SWEEP_REPS", "3"))
    for scale in scales:
        if workload == "config4":
            codes, offsets, qs = bench.synth_config4()
        else:
            codes, offsets = bench.synth_db(scale=scale)
            qs = bench.synth_queries(150) if workload == "short" else bench.load_queries(swb)[1]
        ref = None
        for spec in optsets:
            opts = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in spec.split(",") if kv)
            shard = opts.pop("shard", None)
            nshards = opts.pop("nshards", 1)
            qgroups, qgroup = opts.pop("qgroups", 1), opts.pop("qgroup", 0)
            allq = qs
            if qgroups > 1:  # the queries one rank of a P x R layout gets
                gof = swb.layout_query_groups(swb.pack_sequences(allq)[1], qgroups)
                sel = [q for q, g in zip(allq, gof) if g == qgroup]
            else:
                sel = allq
            qcodes, qoffs = swb.pack_sequences(sel)
            cells = float(sum(len(q) for q in sel)) * float(offsets[-1])
            eng = swb.Engine(0, **opts)
            try:
                eng.db_load(codes, offsets, shard or 0, nshards)
                frac = float(eng.stats()["db_residues"]) / float(offsets[-1])
                ms = []
                for _ in range(reps + 1):
                    eng.search_batch_packed(qcodes, qoffs, fetch=False)
                    ms.append(eng.stats()["device_ms"])
                st = eng.stats()
                best = min(ms[1:])
                chk = int(np.sum(eng.fetch_scores(len(sel) - 1).astype(np.int64)) + np.sum(eng.fetch_scores(0).astype(np.int64)))
                if nshards == 1 and qgroups == 1:
                    ref = chk if ref is None else ref
                print("%s scale %.3f opts {%s}: %.1f GCUPS (best of %d, %.2f ms; mean %.2f ms) tiles %s recomputed %d "
                      "launches %d checksum %s" % (workload, scale, spec, cells * frac / (best * 1e-3) * 1e-9, reps, best,
                                                   float(np.mean(ms[1:])), st["tiles_by_group"], st["recomputed_tiles"],
                                                   st["kernel_launches"],
                                                   "same" if nshards > 1 or qgroups > 1 or chk == ref else "DIFFERENT %d vs %d" % (chk, ref)),
                      flush=True)
            finally:
                eng.close()


if __name__ == "__main__":
    main()
