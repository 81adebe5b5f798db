"""configs[3] lone-query timing helper: python tools/c4_lone.py QLEN [option=value ...] (run plain, then under ncu)"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
swb = importlib.import_module(bench.PKG)
codes, offsets, qs = bench.synth_config4()
ql = int(sys.argv[1])
opts = {k: int(v) for k, v in (a.split("=") for a in sys.argv[2:])}
q = [x for x in qs if len(x) == ql][0]
e = swb.Engine(0, **opts)
e.db_load(codes, offsets)
for _ in range(3):
    t = time.time(); e.search(q); dt = time.time() - t
st = e.stats()
print("qlen %d: %.2f ms wall, device %.2f ms, %.0f GCUPS, launches %d, recomputed %d" % (
    ql, dt * 1e3, st["device_ms"], ql * float(offsets[-1]) / st["device_ms"] * 1e-6, st["kernel_launches"], st["recomputed_tiles"]))
