"""N > 1 host path on CPU: two gloo ranks each take one residue-balanced shard (same planner the engine uses),
score it, and exchange only small results (per-shard (ids, scores) and top-k lists) -- no data-path collective.
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
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes
        from oracle_lib import Oracle, pack_db, read_fasta, read_query
        swb = importlib.import_module(PKG)
        o = Oracle()
        _, seqs = read_fasta(os.path.join(GOLDEN, "uniprot_subset.fasta"))
        codes, offs = pack_db([o.encode(s) for s in seqs])
        m = o.matrix("blosum50")
        query = o.encode(read_query(os.path.join(GOLDEN, "queries", "P02232.fasta")))
        info, _, ids = swb.plan_describe(offs, rank, world, want_ids=True)
        emu = ctypes.CDLL(os.path.join(ROOT, PKG, "lib", "libswbemu.so"))
        u8p, i8p = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int8)
        u64p, i32p = ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_int32)
        emu.swbemu_search.restype = ctypes.c_int
        emu.swbemu_search.argtypes = [u8p, u64p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, i8p,
                                      ctypes.c_int, u8p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                      ctypes.c_int, ctypes.c_uint32, i32p, ctypes.POINTER(ctypes.c_uint32), u8p,
                                      ctypes.c_uint32, i32p, ctypes.c_uint32]
        mine = np.zeros(info.n_local, dtype=np.int32)
        rc = ctypes.c_uint32()
        assert emu.swbemu_search(codes.ctypes.data_as(u8p), offs.ctypes.data_as(u64p), len(seqs), rank, world, 384,
                                 m.ctypes.data_as(i8p), 2, query.ctypes.data_as(u8p), len(query), 0, 0, 0, -1, 8192,
                                 mine.ctypes.data_as(i32p), ctypes.byref(rc), None, 0, None, 0) == 0
        order = np.lexsort((ids, -mine))[:10]
        parts = [None] * world
        dist.all_gather_object(parts, (ids.tolist(), mine.tolist(), ids[order].tolist(), mine[order].tolist()))
        dist.barrier()
        if rank == 0:
            full = swb.merge_shard_scores(len(seqs), [(p[0], p[1]) for p in parts])
            want = o.scan(query, codes, offs, m)
            tid, ts = swb.merge_topk([(p[2], p[3]) for p in parts], 10)
            worder = np.lexsort((np.arange(len(want)), -want))[:10]
            ok = bool(np.array_equal(full, want) and np.array_equal(tid, worder.astype(np.uint32)) and
                      np.array_equal(ts, want[worder]))
            q.put(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_sharded_scan_gloo(world):
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
    assert q.get(timeout=10) is True


def test_merge_helpers():
    swb = importlib.import_module(PKG)
    full = swb.merge_shard_scores(5, [([0, 3], [7, 1]), ([1, 2, 4], [5, 9, 9])])
    assert full.tolist() == [7, 5, 9, 1, 9]
    with pytest.raises(swb.SwbError):
        swb.merge_shard_scores(5, [([0, 3], [7, 1]), ([1, 2], [5, 9])])
    ids, top = swb.merge_topk([([2, 0], [9, 7]), ([4, 1], [9, 5])], 3)
    assert ids.tolist() == [2, 4, 0] and top.tolist() == [9, 9, 7]
