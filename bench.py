#!/usr/bin/env python3
"""Headline benchmark: GCUPS of the Smith-Waterman database scan (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Swiss-Prot-shaped synthetic database (570,065 sequences, ~2.0e8
residues, numpy seed 1782, log-normal lengths plus a 5k..35,213 tail) against the reference's standard
20-query set (144..5478 residues, tests/golden/queries). One "step" = one scan of the whole database
by all 20 queries. With N > 1 (launched by torchrun, one rank per GPU) the database is residue-sharded
across the ranks (strong scaling: the total work is fixed) and rank 0 merges per-rank top hits.

value   = true cells (sum qlen x sum len) / device time, database already resident in HBM
e2e     = the same through the C ABI with HOST buffers: swb_db_load (plan + H2D + device pack) +
          swb_search_batch with the score matrix copied back, wall clock around the calls
roofline= score kernel against the integer SIMD issue peak measured live by swb_microbench
cpu_baseline / --impl reference = the CPU oracle (a port of the reference recurrence, OpenMP over
          sequences, all host cores) on a bounded sample of the same workload
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PKG = "ece1782-smith-waterman-cuda_b200"
METRIC = "GCUPS (whole box, device-timed) on Swiss-Prot scan at 1/2/4/8 B200"

# Swiss-Prot amino-acid composition in percent (UniProt release statistics)
COMPOSITION = {"L": 9.65, "A": 8.25, "G": 7.07, "V": 6.86, "E": 6.72, "S": 6.64, "I": 5.91, "K": 5.80, "R": 5.53,
               "D": 5.46, "T": 5.35, "P": 4.74, "N": 4.06, "Q": 3.93, "F": 3.86, "Y": 2.92, "M": 2.41, "H": 2.27,
               "C": 1.38, "W": 1.10}
ORDER = "ARNDCQEGHILKMFPSTWYVBJZX"


def synth_db(n=570000, seed=1782, scale=1.0):
    """config 2 generator: returns (codes uint8, offsets uint64[n+1]); ids are shuffled w.r.t. length."""
    rng = np.random.default_rng(seed)
    n = int(n * scale)
    lens = np.clip(np.round(rng.lognormal(5.58, 0.75, n)), 2, 35213).astype(np.int64)
    ntail = max(1, int(64 * scale))
    tail = np.round(np.exp(rng.uniform(np.log(5000), np.log(35213), ntail))).astype(np.int64)
    lens = np.concatenate([lens, tail, [35213]])
    rng.shuffle(lens)
    total = int(lens.sum())
    p = np.zeros(32)
    for ch, f in COMPOSITION.items():
        p[ORDER.index(ch)] = f
    p = p / p.sum() * 0.999
    for code in (23, 20, 22, 24):  # X, B, Z and an unknown ('U' encodes to '*' in the reference)
        p[code] += 0.001 / 4
    codes = rng.choice(32, size=total, p=p / p.sum()).astype(np.uint8)
    offsets = np.zeros(len(lens) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    return codes, offsets


def synth_db_uniprot_scale(parts=10):
    """configs[4]: ~10 x Swiss-Prot: `parts` Swiss-Prot-shaped databases (seeds 1785, 1786, ...) back to back."""
    all_codes, all_lens = [], []
    for k in range(parts):
        c, o = synth_db(seed=1785 + k)
        all_codes.append(c)
        all_lens.append(np.diff(o.astype(np.int64)))
    codes = np.concatenate(all_codes)
    lens = np.concatenate(all_lens)
    offsets = np.zeros(len(lens) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    return codes, offsets


def synth_queries(n=1000, seed=1785):
    """configs[4]: queries drawn from the database length law, clipped to [30, 5478], Swiss-Prot composition"""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.round(rng.lognormal(5.58, 0.75, n)), 30, 5478).astype(np.int64)
    p = np.zeros(32)
    for ch, f in COMPOSITION.items():
        p[ORDER.index(ch)] = f
    p /= p.sum()
    return [rng.choice(32, size=int(l), p=p).astype(np.uint8) for l in lens]


def synth_config4(seed=1784):
    """configs[3] (SURVEY 8d "Config 4"): queries of 5,000 / 10,000 / 20,000 / 35,213 residues; database = 256 random
    targets, log-uniform in [5,000, 35,213], plus a 10 %-mutated copy of every query and the query itself (self score
    ~5.5 x L >> 32767: forces the int32 recompute). Uniform 20-letter residues."""
    rng = np.random.default_rng(seed)
    tl = np.round(np.exp(rng.uniform(np.log(5000), np.log(35213), 256))).astype(np.int64)
    seqs = [rng.integers(0, 20, int(l)).astype(np.uint8) for l in tl]
    qs = []
    for ql in (5000, 10000, 20000, 35213):
        q = rng.integers(0, 20, ql).astype(np.uint8)
        mutated = q.copy()
        pos = rng.choice(ql, ql // 10, replace=False)
        mutated[pos] = rng.integers(0, 20, len(pos))
        qs.append(q)
        seqs += [mutated, q.copy()]
    lens = np.array([len(x) for x in seqs], dtype=np.int64)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    return np.concatenate(seqs), offsets, qs


def load_queries(swb):
    qdir = os.path.join(ROOT, "tests", "golden", "queries")
    names = sorted(fn[:-6] for fn in os.listdir(qdir) if fn.endswith(".fasta"))
    qs = []
    for nme in names:
        txt = "".join(open(os.path.join(qdir, nme + ".fasta")).read().split("\n")[1:])
        qs.append(swb.encode(txt) if swb else txt)
    return names, qs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().split("\n"):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nme)
        busy = [s for s, pw in zip(sm, power) if pw > 200] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def host_threads():
    """threads for the CPU legs: every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which would
    turn the reference arm into a one-core run if the OpenMP default were used)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample_gcups(codes, offsets, qnames, qtexts, budget_s, threads=0, keep=None):
    """Oracle (port of the reference recurrence) on a bounded stride sample of the workload. keep: a list that
    receives (stride, scores of query k at the sampled sequences) for a parity check against the GPU's scores."""
    from oracle_lib import Oracle
    o = Oracle()
    cores = threads or host_threads()
    m = o.matrix("blosum50")
    n = len(offsets) - 1
    total_cells = float(sum(len(q) for q in qtexts)) * float(offsets[-1])
    target = cores * 0.65e9 * budget_s  # the port runs at ~0.7 GCUPS per host thread: budget_s seconds of CPU work
    stride = max(1, int(np.ceil(total_cells / target)))
    lens = np.diff(offsets.astype(np.int64))
    sample_res = int(lens[0::stride].sum())
    t0 = time.time()
    cells = 0
    for q in qtexts:
        qc = o.encode(q)
        sc = o.scan(qc, codes, offsets, m, 2, 0, stride, cores)
        if keep is not None:
            keep.append((stride, sc[0::stride].copy()))
        cells += len(qc) * sample_res
    dt = time.time() - t0
    return {"value": cells / dt * 1e-9, "unit": "GCUPS", "cores": cores, "kind": "port",
            "sample": "oracle/sw_oracle.c (OpenMP over sequences, %d threads), every %d-th of %d sequences x all %d "
                      "queries, %.3g cells in %.1f s" % (cores, stride, n, len(qtexts), cells, dt)}, dt


def run_reference(args):
    """--impl reference: the reference's CPU formulation (score-only port, all host threads) on this config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    codes, offsets = synth_db(scale=args.scale)
    names, qtexts = load_queries(None)
    budget = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    times, vals, last = [], [], None
    for it in range(args.warmup + args.steps):
        cb, dt = cpu_sample_gcups(codes, offsets, names, qtexts, budget)
        if it >= args.warmup:
            times.append(dt)
            vals.append(cb["value"])
        last = cb
    v = float(np.mean(vals))
    last["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(offsets, qtexts, args), "cpu_baseline": last,
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def ref_cuda_leg(qnames, budget_scale=0.25):
    """The reference's own CUDA solver (SWSolver.cu:201-264 behind oracle/_ref/ref_cuda_scan, compiled unmodified for
    sm_100a) on the same GPU in the same run, as a reported baseline: a 1/4-scale copy of the benchmark database (its
    fixed 400 MB residue buffer and batching do not take the full one in reasonable time) against the reference
    queries that fit its 1024-row limit (SWSolver.cu:85). Solver-call seconds exclude its FASTA parsing."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_cuda_scan")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_cuda_scan not built (needs /root/reference at build time)"}
    import tempfile
    letters = np.frombuffer(b"ARNDCQEGHILKMFPSTWYVBJZXU", dtype=np.uint8)
    codes, offsets = synth_db(scale=budget_scale)
    text = letters[np.minimum(codes, 24)].tobytes()
    out = {"kernel": "f_scoreSequenceTiledCoalesced (reference SWSolver.cu, unmodified, -arch=sm_100a)", "unit": "GCUPS",
           "db_sequences": int(len(offsets) - 1), "db_residues": int(offsets[-1]), "queries": {}}
    with tempfile.TemporaryDirectory() as tmp:
        dbpath = os.path.join(tmp, "db.fasta")
        with open(dbpath, "wb") as f:
            for i in range(len(offsets) - 1):
                f.write(b">s%d\n" % i)
                f.write(text[int(offsets[i]):int(offsets[i + 1])])
                f.write(b"\n")
        cells = secs = 0.0
        for nme in qnames:
            qpath = os.path.join(ROOT, "tests", "golden", "queries", nme + ".fasta")
            try:
                r = subprocess.run([exe, qpath, dbpath, "1"], capture_output=True, text=True, timeout=240)
            except subprocess.TimeoutExpired:
                out["queries"][nme] = "timeout"
                continue
            tl = [l for l in r.stdout.split("\n") if l.startswith("#TIME")]
            if r.returncode != 0 or not tl:
                out["queries"][nme] = "failed (exit %d)" % r.returncode
                continue
            f = dict(kv.split("=") for kv in tl[0].split()[1:])
            c = float(f["qlen"]) * float(offsets[-1])
            out["queries"][nme] = {"qlen": int(f["qlen"]), "solver_s": float(f["solver_s"]),
                                   "gcups": c / float(f["solver_s"]) * 1e-9}
            cells += c
            secs += float(f["solver_s"])
        out["value"] = cells / secs * 1e-9 if secs else None
        out["timing"] = "wall clock around smith_waterman_cuda() (pack + managed-memory upload + launches + gather)"
    return out


def workload_config(offsets, qs, args):
    cfg = _workload_config(offsets, qs, args)
    if getattr(args, "affine", ""):
        cfg["scoring"] = "BLOSUM50 ('*' zeroed), affine gaps open,extend = %s (side measurement)" % args.affine
    return cfg


def _workload_config(offsets, qs, args):
    if getattr(args, "workload", "config2") == "config4":
        return {"workload": "configs[3]: long-sequence path, 4 queries of 5,000..35,213 residues x 264 targets of 5k..35k "
                            "(256 random + mutated copies + the queries themselves)",
                "db_sequences": int(len(offsets) - 1), "db_residues": int(offsets[-1]), "queries": len(qs),
                "query_residues": int(sum(len(q) for q in qs)), "scoring": "BLOSUM50 ('*' zeroed), linear gap 2",
                "l2": "boundary scratch flushed through L2 between passes; inputs smaller than L2"}
    if getattr(args, "workload", "config2") == "config5":
        return {"workload": "configs[4]: 1,000 synthetic queries x UniProt-scale synthetic DB (10 Swiss-Prot-shaped parts)",
                "db_sequences": int(len(offsets) - 1), "db_residues": int(offsets[-1]), "queries": len(qs),
                "query_residues": int(sum(len(q) for q in qs)), "scoring": "BLOSUM50 ('*' zeroed), linear gap 2",
                "l2": "inputs larger than L2"}
    return {"workload": "configs[1]: Swiss-Prot-shaped synthetic DB (seed 1782) x reference 20-query set",
            "db_sequences": int(len(offsets) - 1), "db_residues": int(offsets[-1]), "queries": len(qs),
            "query_residues": int(sum(len(q) for q in qs)), "scoring": "BLOSUM50 ('*' zeroed), linear gap 2",
            "l2": "inputs larger than L2 (packed residues + boundary scratch >> 126 MB)", "scale": args.scale}


def sample_parity(swb, eng, codes, offsets, queries, local_of, fetch_row, threads, seconds):
    """Oracle against the engine on a strided sample of THIS rank's shard for the given queries (global indices ->
    code arrays). fetch_row(local query index) returns the shard's score vector. Returns (ok, description)."""
    from oracle_lib import Oracle
    o = Oracle()
    ids = eng.db_ids()
    if len(ids) == 0 or not queries:
        return True, "empty shard or no queries"
    lens = np.diff(offsets.astype(np.int64))[ids]
    cells = float(sum(len(q) for q in queries.values())) * float(lens.sum())
    stride = max(1, int(np.ceil(cells / (threads * 0.6e9 * seconds))))
    pos = np.arange(0, len(ids), stride)
    sel = ids[pos]
    sc, so = swb.pack_sequences([codes[int(offsets[i]):int(offsets[i + 1])] for i in sel])
    m = o.matrix("blosum50")
    ok = True
    for qi, q in queries.items():
        want = o.scan(q, sc, so, m, 2, 0, 1, threads)
        got = fetch_row(local_of[qi])[pos]
        ok = ok and bool(np.array_equal(got, want))
    return ok, "oracle vs GPU, every %d-th sequence of the rank's shard (%d sequences) x %d queries" % (
        stride, len(sel), len(queries))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--scale", type=float, default=1.0, help="database size factor (1.0 = the named workload)")
    ap.add_argument("--workload", default="config2", help="config2 (default, the headline), config4 (long sequences) or "
                    "config5 (1,000 queries x UniProt-scale database; meant for 8 GPUs)")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="warm end-to-end repetitions (default 2; config5: 1)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--parity-seconds", type=float, default=3.0, help="CPU budget of the per-rank oracle sample (N > 1)")
    ap.add_argument("--topk", type=int, default=10)
    ap.add_argument("--db-parts", type=int, default=0, help="database parts P of the P x R rank layout (0 = the layout "
                    "rule of the engine group: the largest divisor of N with >= 450,000 sequences per part, more parts when the "
                    "batch cannot fill the query groups evenly)")
    ap.add_argument("--group", action="store_true", help="one process drives all --gpus devices through the engine "
                    "group (swb_group_*) instead of one torchrun rank per GPU")
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--group-len", type=int, default=0)
    ap.add_argument("--group-order", type=int, default=0)
    ap.add_argument("--split", type=int, default=-1, help="pipelined passes for very long tiles: 1 on, 0 off, -1 "
                    "(default) = the engine's rule: on for small shards")
    ap.add_argument("--split-k", type=int, default=0, help="rows per lane of the pipelined groups: 0 auto, 8, 16")
    ap.add_argument("--direct-len", type=int, default=-1, help="tiles / queries at least this long are scored by the "
                    "rebased s16 policy at once (-1 = engine default 16000, 0 = never)")
    ap.add_argument("--exact", type=int, default=-1, help="exact passes: 0 rebased s16 (default), 1 int32")
    ap.add_argument("--batch-order", type=int, default=-1, help="0 longest query first (default), 1 as given")
    ap.add_argument("--chunk-rows", type=int, default=0)
    ap.add_argument("--xl-len", type=int, default=0)
    ap.add_argument("--synth-queries", type=int, default=0, help="N > 0: N synthetic queries of the configs[4] length law "
                    "instead of the 20 reference queries (a side measurement: short-query mix on the configs[1] DB)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference-CUDA-kernel leg (N = 1 default run)")
    ap.add_argument("--affine", default="", help="GO,GE: affine gaps instead of the reference's linear gap 2 (a side "
                    "measurement of the V16A kernels; not the headline metric, no CPU leg)")
    ap.add_argument("--per-query", action="store_true", help="also print device GCUPS per query")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    swb = importlib.import_module(PKG)
    side = False  # a side measurement (not the named workload): no CPU / reference legs
    if args.workload == "config5":
        codes, offsets = synth_db_uniprot_scale()
        qs = synth_queries()
        names = ["q%d" % i for i in range(len(qs))]
        side = True
    elif args.workload == "config4":
        codes, offsets, qs = synth_config4()
        names = ["L%d" % len(q) for q in qs]
        side = True
    else:
        codes, offsets = synth_db(scale=args.scale)
        names, qs = load_queries(swb)
        if args.synth_queries:
            qs = synth_queries(args.synth_queries)
            names = ["q%d" % i for i in range(len(qs))]
            side = True
    if args.affine:
        side = True
    if side:
        args.no_cpu = True
        args.no_ref_cuda = True
    big = args.workload == "config5"  # no nq x n host matrix at UniProt scale: device-side hit lists instead
    if args.e2e_steps < 0:
        args.e2e_steps = 1 if big else 2
    n_total = len(offsets) - 1
    total_cells = float(sum(len(q) for q in qs)) * float(offsets[-1])
    if args.group:
        return run_group(args, swb, codes, offsets, qs, total_cells)

    # rank layout: P database parts x R query groups (the engine group's rule); rank -> (part, group)
    _, qoffs_all = swb.pack_sequences(qs)
    parts = args.db_parts if args.db_parts else swb.layout_parts(n_total, world, qoffsets=qoffs_all)
    if world % parts:
        raise SystemExit("--db-parts must divide the number of ranks")
    groups = world // parts
    part, grp = rank % parts, rank // parts
    group_of = swb.layout_query_groups(qoffs_all, groups)
    mine = [qi for qi in range(len(qs)) if group_of[qi] == grp]  # global indices of this rank's queries
    local_of = {qi: k for k, qi in enumerate(mine)}
    qcodes, qoffs = swb.pack_sequences([qs[qi] for qi in mine])
    my_cells = float(sum(len(qs[qi]) for qi in mine))

    opts = {}
    for key, val, unset in (("streams", args.streams, 0), ("k", args.k, 0), ("group_len", args.group_len, 0),
                            ("group_order", args.group_order, 0), ("split", args.split, -1),
                            ("split_k", args.split_k, 0), ("direct_len", args.direct_len, -1), ("exact", args.exact, -1),
                            ("batch_order", args.batch_order, -1), ("chunk_rows", args.chunk_rows, 0),
                            ("xl_len", args.xl_len, 0)):
        if val != unset:
            opts[key] = val
    eng = swb.Engine(local, **opts)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    if args.affine:
        go, ge = (int(x) for x in args.affine.split(","))
        eng.set_scoring_affine(swb.scoring_matrix(swb.SWB_SCORING_BLOSUM50_REF)[0], go, ge)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # One step = one scan of the whole database by all queries, scores (config5: the k best per query, selected on the
    # device) copied back to host memory inside the step: SURVEY 8(d) times "first scoring kernel to last score D2H".
    out = None

    def step():
        if big:
            return eng.search_batch_topk(qcodes, qoffs, args.topk)
        eng.search_batch_packed(qcodes, qoffs, fetch=True, out=out)
        return None

    # cold end to end: the first swb_db_load (allocations, plan, upload, pack) + the first scan of this process
    barrier()
    t0 = time.perf_counter()
    eng.db_load(codes, offsets, part, parts)
    nloc = eng.db_count()
    if not big:
        out = np.zeros((len(mine), nloc), dtype=np.int32)
    step()
    torch.cuda.synchronize()
    cold_s = max_over_ranks(time.perf_counter() - t0)
    cold_load_ms = eng.stats()["load_ms"]

    for _ in range(max(0, args.warmup - 1)):  # the cold call above was the first warm-up step
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    launches = 0
    for _ in range(args.steps):
        step()
        launches += eng.stats()["kernel_launches"]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_max = max_over_ranks(ev0.elapsed_time(ev1))
    st = eng.stats()
    ms_per_step = ms_max / args.steps
    value = total_cells / (ms_per_step * 1e-3) * 1e-9
    lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
    pc = torch.tensor([float(st["padded_cells"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(lt)
        dist.all_reduce(pc)
    launches, padded_cells = int(lt.item()), float(pc.item())

    # ---- parity on every rank: oracle on a strided sample of the rank's shard ------------------------------------
    def fetch_row(k):
        return eng.fetch_scores(k) if big else out[k]

    threads = max(1, host_threads() // world)
    sample_ok, sample_desc = None, None
    if args.workload == "config4":
        # every query against its mutated copy and itself (the beyond-s16 hits) and 6 targets
        from oracle_lib import Oracle
        o = Oracle()
        ids = eng.db_ids()
        sel = [i for i in ids if i >= 256 or i < 6]
        pos = {int(g): k for k, g in enumerate(ids)}
        sc, so = swb.pack_sequences([codes[int(offsets[i]):int(offsets[i + 1])] for i in sel])
        sample_ok = True
        for qi in mine:
            got = fetch_row(local_of[qi])[[pos[int(i)] for i in sel]]
            sample_ok = sample_ok and bool(np.array_equal(got, o.scan(qs[qi], sc, so, o.matrix("blosum50"))))
        sample_desc = "oracle vs GPU: every query x (6 targets + the mutated copies + the queries themselves)"
    elif not args.affine and (world > 1 or big or args.no_cpu):
        chosen = mine if not big else mine[::max(1, len(mine) // 5)][:5]
        sample_ok, sample_desc = sample_parity(swb, eng, codes, offsets, {qi: qs[qi] for qi in chosen}, local_of,
                                               fetch_row, threads, args.parity_seconds)
    if sample_ok is not None and world > 1:
        t = torch.tensor([1 if sample_ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        sample_ok = bool(t.item())

    # ---- hit lists: device-side selection per rank, merged on the host, checked three ways ---------------------------
    #  (1) every rank: its device list == the host selection over the full vector of its shard
    #  (2) rank 0: merged lists == the k best of the merged FULL vector   (3) == oracle scores of those ids
    top_ok, top_desc = None, None
    if not args.affine and len(qs) and n_total:
        k = args.topk
        dev_ids, dev_top = eng.search_batch_topk(qcodes, qoffs, k)
        check = mine if not big else mine[::max(1, len(mine) // 3)][:3]
        ids_mine = eng.db_ids()
        ok1 = True
        fulls = {}
        for qi in check:
            row = eng.fetch_scores(local_of[qi])
            hi, ht = eng.topk(row, k)
            ok1 = ok1 and bool(np.array_equal(hi, dev_ids[local_of[qi]]) and np.array_equal(ht, dev_top[local_of[qi]]))
            fulls[qi] = row
        payload = (part, ids_mine, {qi: (dev_ids[local_of[qi]], dev_top[local_of[qi]]) for qi in mine}, fulls, ok1)
        gathered = [payload]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, payload)
        if rank == 0:
            from oracle_lib import Oracle
            o = Oracle()
            top_ok = all(g[4] for g in gathered)
            nmerged = 0
            for qi in range(len(qs)):
                holders = [g for g in gathered if qi in g[2]]
                mi, mt = swb.merge_topk([g[2][qi] for g in holders], k)
                top_ok = top_ok and len(holders) == parts and len(mi) == min(k, n_total)
                nmerged += 1
                full_holders = [g for g in holders if qi in g[3]]
                if len(full_holders) == parts:
                    full = swb.merge_shard_scores(n_total, [(g[1], g[3][qi]) for g in full_holders])
                    order = np.lexsort((np.arange(n_total), -full.astype(np.int64)))[:k]
                    top_ok = top_ok and bool(np.array_equal(mi, order.astype(np.uint32)) and np.array_equal(mt, full[order]))
                    sc, so = swb.pack_sequences([codes[int(offsets[i]):int(offsets[i + 1])] for i in mi])
                    top_ok = top_ok and bool(np.array_equal(o.scan(qs[qi], sc, so, o.matrix("blosum50")), mt))
            nfull = len(set(qi for g in gathered for qi in g[3]))
            top_desc = ("per-rank device top-%d == host selection over the shard's full vector (%d queries); merged "
                        "lists of %d queries (%d parts each); of those, %d also == top-%d of the merged full vector == "
                        "oracle scores of those ids" % (k, nfull, nmerged, parts, nfull, k))

    # ---- warm end to end through the C ABI with host buffers (database upload + scan + results back), wall clock ----
    e2e_times = []
    for it in range(args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        eng.db_load(codes, offsets, part, parts)
        step()
        torch.cuda.synchronize()
        e2e_times.append(max_over_ranks(time.perf_counter() - t0))
    e2e_s = float(np.mean(e2e_times)) if e2e_times else cold_s
    load_ms = eng.stats()["load_ms"]
    shard_bytes = int(eng.stats()["db_residues"]) if parts > 1 else int(codes.nbytes)
    d2h = 8 * args.topk * len(mine) if big else int(4 * len(mine) * nloc)
    bt = torch.tensor([float(shard_bytes + qcodes.nbytes + qoffs.nbytes), float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(bt)
    e2e = {"value": total_cells / e2e_s * 1e-9, "unit": "GCUPS",
           "h2d_bytes_per_step": int(bt[0].item()) + int(offsets.nbytes), "d2h_bytes_per_step": int(bt[1].item()),
           "seconds_per_step": e2e_s, "db_load_ms": load_ms,
           "warm": "mean of %d repetitions after the first (buffers already allocated)" % len(e2e_times) if e2e_times
                   else "not run: value is the cold call",
           "cold": {"value": total_cells / cold_s * 1e-9, "seconds_per_step": cold_s, "db_load_ms": cold_load_ms,
                    "what": "first swb_db_load + first scan of the process (allocations, module load)"},
           "api": "swb_db_load + %s (host buffers)" % ("swb_search_batch_topk" if big else "swb_search_batch")}

    if rank == 0:
        # roofline of the dominant kernel (swb_score_kernel<K,V16>): the ALU pipe (64 lanes/clk/SM). Per cell pair the
        # kernel issues prmt + vimax3.relu + 1/2 vimax3 on that pipe (2.5) and two vadd2, which the pairwise
        # microbenchmark shows issuing on another pipe (viaddmax+vadd2 runs at twice the single rate); if they did
        # not, the count would be 4.5. Peak = measured single-instruction issue rate / ALU instructions per cell.
        # (Until profiles/r2x the cell had 3.5 ALU-pipe instructions; frac_vs_3.5 keeps that ceiling for comparison.)
        alu = swb.microbench(local, 0)      # Glane-instr/s of viaddmax.relu alone = the ALU pipe's issue rate
        pair = swb.microbench(local, 11)    # viaddmax + vadd2 in the same loop
        mix = swb.microbench(local, 14)     # dependent-chain loop of the whole 4.5-instruction mix (a lower bound)
        per_kind = {swb.MICROBENCH_KINDS[k]: round(swb.microbench(local, k), 1) for k in (0, 1, 2, 3, 5, 6, 11, 12, 13)}
        vadd_off_alu = pair > 1.5 * alu
        instr_per_cell = (2.5 if vadd_off_alu else 4.5) / 2.0
        ach_padded = padded_cells / (ms_per_step * 1e-3) * 1e-9  # cells the kernels really execute, all ranks
        peak_gcups = alu / instr_per_cell
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # algorithmic bytes: one byte of database residue per query pass + 4 B of score per (query, sequence)
        alg_bytes = float(sum(1 for _ in qs)) * float(offsets[-1]) + 4.0 * len(qs) * n_total
        traffic, traffic_src = None, "no ncu --set full capture of this configuration is committed"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            key = "%s_scale%g_n%d" % (args.workload, args.scale, world)
            if key in tj:
                traffic, traffic_src = tj[key]["dram_bytes_per_launch"], tj[key]["source"]
        except Exception:
            pass
        roofline = {"bound": "int_alu", "kernel": "swb_score_kernel<K,V16>", "achieved": value / world, "peak": peak_gcups,
                    "unit": "GCUPS", "frac": (value / world) / peak_gcups,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "traffic_algorithmic_per_step": alg_bytes,
                    "achieved_incl_padding": ach_padded / world,
                    "frac_incl_padding": (ach_padded / world) / peak_gcups,
                    "alu_instr_per_cell_pair": {"used": 2 * instr_per_cell, "if_vadd2_on_alu_pipe": 4.5,
                                                "if_vadd2_off_alu_pipe": 2.5, "vadd2_off_alu_pipe": bool(vadd_off_alu),
                                                "frac_if_4.5": (value / world) / (alu / 2.25),
                                                "frac_vs_3.5": (value / world) / (alu / 1.75),
                                                "issue_slots_per_cell_pair": 5.0,
                                                "what": "2.5 = prmt + vimax3.relu + 1/2 vimax3; beside them 2 vadd2 on the "
                                                        "other 64-lane pipe and 1/2 LDS: 5.0 issue slots per cell pair "
                                                        "against one slot per clock and scheduler, so the issue-slot "
                                                        "ceiling (alu x 2 / 2.5 per cell) is practically the same number; "
                                                        "frac_vs_3.5 = against the ceiling of the round-1 cell "
                                                        "(prmt + 2 viaddmax + 1/2 vimax3)"},
                    "peak_source": "measured live: ALU pipe issues %.0f Glane-instr/s (swb_microbench, viaddmax.relu alone); "
                                   "%.1f ALU-pipe instructions per cell pair (prmt + vimax3.relu + 1/2 vimax3%s); "
                                   "a dependent-chain loop of the full mix reaches %.0f Glane-instr/s" % (
                                       alu, 2 * instr_per_cell,
                                       "; vadd2 issues on another pipe: viaddmax+vadd2 = %.0f" % pair if vadd_off_alu
                                       else " + 2 vadd2", mix),
                    "instr_rates_glane_per_s": per_kind,
                    "hbm": {"achieved": alg_bytes / (ms_per_step * 1e-3) * 1e-9, "peak": hbm_peak * world, "unit": "GB/s",
                            "frac": alg_bytes / (ms_per_step * 1e-3) * 1e-9 / (hbm_peak * world),
                            "what": "database streaming: algorithmic bytes per step / step time (the path does "
                                    "hundreds of cell updates per byte: HBM is not the bound)",
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
        # the one HBM-bound kernel of the path: swb_pack_kernel reads the shard's raw residues once and writes the
        # interleaved copy once per database load (device time from the engine's CUDA events, rank 0's shard)
        try:
            pack_us, pack_bytes = eng.pack_time(10)
            roofline["hbm_pack"] = {"kernel": "swb_pack_kernel", "achieved": pack_bytes / (pack_us * 1e-6) * 1e-9,
                                    "peak": hbm_peak, "unit": "GB/s",
                                    "frac": pack_bytes / (pack_us * 1e-6) * 1e-9 / hbm_peak,
                                    "bytes_per_launch": pack_bytes, "us": pack_us,
                                    "us_inside_db_load": st.get("pack_us"),
                                    "what": "residues read + packed residues written, once per swb_db_load; mean of 10 "
                                            "back-to-back launches (inside swb_db_load the same kernel follows an upload "
                                            "during which the SMs were idle)"}
        except Exception as ex:
            roofline["hbm_pack"] = {"unavailable": repr(ex)}
        cfg = workload_config(offsets, qs, args)
        cfg["layout"] = {"db_parts": parts, "query_groups": groups,
                         "rule": "P = largest divisor of N with >= 450,000 sequences per part; P = N when the batch "
                                 "cannot fill N / P query groups with >= 8 queries each, evenly (LPT, heaviest group "
                                 "<= 1.02 x mean)"}
        cfg["value_includes"] = ("device-side top-%d per query copied to the host" % args.topk if big
                                 else "all scores copied to the host (4 B x queries x sequences per step)")
        line = {"metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "s16x2 (exact rebased-s16 / int32 recompute on overflow)",
                "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "engine": {k: st[k] for k in ("tiles", "tiles_by_group", "last_k",
                                                                    "recomputed_tiles", "sm_count")},
                "topk_merge_ok": top_ok, "topk_check": top_desc, "sample_parity_ok": sample_ok,
                "sample_parity": sample_desc}
        if not args.no_cpu and world == 1:  # the CPU baseline is reported at N = 1 only
            names_t, qtexts = load_queries(None)
            kept = []
            line["cpu_baseline"], _ = cpu_sample_gcups(codes, offsets, names_t, qtexts, args.cpu_seconds, keep=kept)
            # the oracle's scores of that sample double as a parity check of the measured configuration (full-size
            # database, the engine's own choice of group_len and K): every query, every sampled sequence, bit-exact
            try:
                line["sample_parity_ok"] = bool(all(
                    np.array_equal(out[local_of[qi]][0::stride], want) for qi, (stride, want) in enumerate(kept)))
                line["sample_parity"] = "oracle vs GPU scores, every %d-th sequence x %d queries" % (kept[0][0], len(kept))
            except Exception as ex:  # never lose the bench line to the checker
                line["sample_parity"] = "check failed to run: %r" % (ex,)
        else:
            line["cpu_baseline"] = None
        if not args.no_ref_cuda and world == 1:
            try:
                line["ref_cuda_baseline"] = ref_cuda_leg(["P02232", "P01008", "P27895"])
            except Exception as ex:
                line["ref_cuda_baseline"] = {"unavailable": repr(ex)}
        if world == 1 and not args.affine and not big and top_ok is not None:
            # traceback of the hit lists (cpu.cpp:76-108): every top-k hit of every query in one
            # launch of swb_align_batch, wall clock incl. the copies; the aligned scores must equal the scan's
            try:
                hits = [(local_of[qi], int(sid)) for qi in mine for sid in dev_ids[local_of[qi]] if int(sid) != 0xFFFFFFFF]
                want = [int(v) for qi in mine for sid, v in zip(dev_ids[local_of[qi]], dev_top[local_of[qi]])
                        if int(sid) != 0xFFFFFFFF]
                lens = [int(offsets[sid + 1] - offsets[sid]) for _, sid in hits]
                mine_q = [qs[qi] for qi in mine]
                eng.align_batch(mine_q, hits, lens)  # allocations
                t0 = time.perf_counter()
                res = eng.align_batch(mine_q, hits, lens)
                dt = time.perf_counter() - t0
                acells = float(sum(len(mine_q[h[0]]) * l for h, l in zip(hits, lens)))
                line["align"] = {"hits": len(hits), "ms": dt * 1e3, "cells": acells, "gcups": acells / dt * 1e-9,
                                 "scores_equal_scan": bool([r[0] for r in res] == want),
                                 "api": "swb_align_batch: top-%d hits of %d queries, one launch, ops copied back" % (
                                     args.topk, len(mine))}
            except Exception as ex:
                line["align"] = {"unavailable": repr(ex)}
        if args.per_query:
            pq = {}
            for nme, q in zip(names, qs):
                eng.search_batch([q], fetch=False)
                eng.search_batch([q], fetch=False)
                s2 = eng.stats()
                pq[nme] = {"qlen": len(q), "gcups": s2["cells"] / (s2["device_ms"] * 1e-3) * 1e-9, "k": s2["last_k"],
                           "ms": s2["device_ms"]}
            line["per_query"] = pq
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def run_group(args, swb, codes, offsets, qs, total_cells):
    """--group: ONE process, every device behind the engine group (what the C++ drop-in uses). Device time = the
    slowest device's CUDA-event time per step (swb_group_stats), wall clock beside it."""
    g = swb.EngineGroup(args.gpus)
    big = args.workload == "config5"
    qcodes, qoffs = swb.pack_sequences(qs)
    # the bench knows its batch before it loads the database: the batch-aware layout rule, as for the ranks above
    g.set_option("db_parts", args.db_parts if args.db_parts else
                 swb.layout_parts(len(offsets) - 1, g.size(), qoffsets=qoffs))
    t0 = time.perf_counter()
    g.db_load(codes, offsets)
    out = None if big else np.zeros((len(qs), len(offsets) - 1), dtype=np.int32)

    def step():
        if big:
            return g.search_batch_topk(qcodes, qoffs, args.topk)
        return g.search_batch_packed(qcodes, qoffs, out=out)

    step()
    cold_s = time.perf_counter() - t0
    for _ in range(max(0, args.warmup - 1)):
        step()
    dev_ms, wall = 0.0, time.perf_counter()
    launches = 0
    for _ in range(args.steps):
        step()
        s = g.stats()
        dev_ms += s["device_ms"]
        launches += s["kernel_launches"]
    wall = time.perf_counter() - wall
    ok = None
    if not big:
        from oracle_lib import Oracle
        o = Oracle()
        stride = 997
        m = o.matrix("blosum50")
        ok = all(np.array_equal(out[qi][0::stride], o.scan(qs[qi], codes, offsets, m, 2, 0, stride, host_threads())[0::stride])
                 for qi in range(0, len(qs), max(1, len(qs) // 4)))
    t1 = time.perf_counter()
    g.db_load(codes, offsets)
    step()
    e2e_s = time.perf_counter() - t1
    cfg = workload_config(offsets, qs, args)
    cfg["layout"] = {"db_parts": g.db_parts(), "query_groups": g.size() // max(1, g.db_parts()), "process": "single (engine group)"}
    line = {"metric": METRIC, "value": total_cells / (dev_ms / args.steps * 1e-3) * 1e-9, "unit": "GCUPS",
            "n_gpus": g.size(), "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "wall_ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "s16x2 (exact rebased-s16 / int32 recompute on overflow)", "data": "synthetic", "config": cfg,
            "e2e": {"value": total_cells / e2e_s * 1e-9, "unit": "GCUPS", "seconds_per_step": e2e_s,
                    "cold": {"value": total_cells / cold_s * 1e-9, "seconds_per_step": cold_s},
                    "api": "swb_group_db_load + swb_group_search_batch%s (host buffers, one process)" % ("_topk" if big else "")},
            "gpu_launches": launches, "sample_parity_ok": ok, "mode": "engine group, one process"}
    print(json.dumps(line))
    g.close()


if __name__ == "__main__":
    main()
