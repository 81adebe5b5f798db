"""N > 1 host path on CPU: gloo ranks take their cell of the P database parts x R query groups layout (same planner and
layout rules the engine group uses), score it, and exchange only small results (per-shard (ids, scores) and top-k
lists) -- no data-path collective.
The merged vector / hit list must equal the unsharded oracle scan. Scores of a shard come from the host
emulation of the warp program (test infrastructure), since this container has no GPU."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, PKG, ROOT


def _worker(rank, world, port, q):
    """One rank of the P x R layout bench.py uses under torchrun (include/swb.h, engine group): rank -> database part
    rank % P, query group rank // P; the shard's scores come from the host emulation of the warp program."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu_lib
        from oracle_lib import Oracle, pack_db, read_fasta, read_query
        swb = importlib.import_module(PKG)
        o = Oracle()
        _, seqs = read_fasta(os.path.join(GOLDEN, "uniprot_subset.fasta"))
        codes, offs = pack_db([o.encode(s) for s in seqs])
        m = o.matrix("blosum50")
        names = ["P02232", "P05013", "P14942", "P01008", "P07327"]
        queries = [o.encode(read_query(os.path.join(GOLDEN, "queries", nme + ".fasta"))) for nme in names]
        _, qoffs = swb.pack_sequences(queries)
        # 60 sequences per part at least: world 2 -> P = 1 (two query groups), world 4 -> P = 1 too; forcing P = world
        # is the pure database sharding
        for parts in (swb.layout_parts(len(seqs), world, 60), world):
            groups = world // parts
            part, grp = rank % parts, rank // parts
            group_of = swb.layout_query_groups(qoffs, groups)
            mine_q = [qi for qi in range(len(queries)) if group_of[qi] == grp]
            info, _, ids = swb.plan_describe(offs, part, parts, want_ids=True)
            res = {}
            for qi in mine_q:
                sc, _ = emu_lib.search(codes, offs, m, queries[qi], K=0, shard=part, nshards=parts, n_out=info.n_local)
                order = np.lexsort((ids, -sc.astype(np.int64)))[:10]
                res[qi] = (sc.tolist(), ids[order].tolist(), sc[order].tolist())
            gathered = [None] * world
            dist.all_gather_object(gathered, (part, ids.tolist(), res))
            dist.barrier()
            if rank == 0:
                ok = sorted(qi for g in gathered for qi in g[2]) == sorted(list(range(len(queries))) * parts)
                for qi, query in enumerate(queries):
                    want = o.scan(query, codes, offs, m)
                    holders = [g for g in gathered if qi in g[2]]
                    full = swb.merge_shard_scores(len(seqs), [(g[1], g[2][qi][0]) for g in holders])
                    tid, ts = swb.merge_topk([(g[2][qi][1], g[2][qi][2]) for g in holders], 10)
                    worder = np.lexsort((np.arange(len(want)), -want.astype(np.int64)))[:10]
                    ok = ok and len(holders) == parts and bool(
                        np.array_equal(full, want) and np.array_equal(tid, worder.astype(np.uint32)) and
                        np.array_equal(ts, want[worder]))
                q.put(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_scan_gloo(world):
    so = os.path.join(ROOT, PKG, "lib", "libswbemu.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, PKG), "emu"], check=True, capture_output=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True  # the layout rule's choice of parts
    assert q.get(timeout=10) is True  # parts = world: pure database sharding


def test_merge_helpers():
    swb = importlib.import_module(PKG)
    full = swb.merge_shard_scores(5, [([0, 3], [7, 1]), ([1, 2, 4], [5, 9, 9])])
    assert full.tolist() == [7, 5, 9, 1, 9]
    with pytest.raises(swb.SwbError):
        swb.merge_shard_scores(5, [([0, 3], [7, 1]), ([1, 2], [5, 9])])
    ids, top = swb.merge_topk([([2, 0], [9, 7]), ([4, 1], [9, 5])], 3)
    assert ids.tolist() == [2, 4, 0] and top.tolist() == [9, 9, 7]
