"""CPU validation of the kernel logic: tests/emu/swb_emu.cu runs the SAME warp program as the GPU
(csrc/swb_warp.cuh) lane by lane on the host -- same tiling plan, packing function, query chunking,
boundary scratch, lane-group wavefront, overflow flagging and int32 recompute -- and must reproduce the
oracle bit for bit. (The GPU run of the same code is covered by tests/test_gpu_parity.py.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT, random_db
from oracle_lib import pack_db

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i8p = ctypes.POINTER(ctypes.c_int8)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_i32p = ctypes.POINTER(ctypes.c_int32)


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(ROOT, PKG, "lib", "libswbemu.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, PKG), "emu"], check=True, capture_output=True)
    L = ctypes.CDLL(so)
    L.swbemu_search.restype = ctypes.c_int
    L.swbemu_search.argtypes = [_u8p, _u64p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _i8p,
                                ctypes.c_int, _u8p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                ctypes.c_int, ctypes.c_uint32, _i32p, ctypes.POINTER(ctypes.c_uint32), _u8p,
                                ctypes.c_uint32, _i32p, ctypes.c_uint32]

    def search(codes, offs, m, q, K=32, group_len=384, force_i32=0, chunk_rows=0, thr=-1, gap=2, shard=0, nshards=1,
               n_out=None, xl_len=8192, q2=None, chunk_rows_pair=0):
        """q2 given: a query-pair job; returns ((scores, scores2), recomputed_tiles)"""
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        if len(codes) == 0:
            codes = np.zeros(1, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        q = np.ascontiguousarray(q, dtype=np.uint8)
        m = np.ascontiguousarray(m, dtype=np.int8)
        n = len(offs) - 1
        out = np.full(n if n_out is None else n_out, -7, dtype=np.int32)
        out2 = np.full(len(out), -7, dtype=np.int32)
        rc = ctypes.c_uint32()
        if q2 is not None:
            q2 = np.ascontiguousarray(q2, dtype=np.uint8)
            if len(q2) == 0:
                q2 = np.zeros(1, dtype=np.uint8)[:0]
        q2p = None if q2 is None else (q2.ctypes.data_as(_u8p) if len(q2) else ctypes.cast(out2.ctypes.data, _u8p))
        r = L.swbemu_search(codes.ctypes.data_as(_u8p), offs.ctypes.data_as(_u64p), n, shard, nshards, group_len,
                            m.ctypes.data_as(_i8p), gap, q.ctypes.data_as(_u8p) if len(q) else None, len(q), K,
                            force_i32, chunk_rows, thr, xl_len, out.ctypes.data_as(_i32p), ctypes.byref(rc), q2p,
                            0 if q2 is None else len(q2), out2.ctypes.data_as(_i32p), chunk_rows_pair)
        assert r == 0
        if q2 is not None:
            return (out, out2), rc.value
        return out, rc.value

    return search


@pytest.mark.parametrize("K,group_len", [(0, 384), (8, 2000), (16, 384), (32, 384), (32, 64), (16, 16), (0, 16), (0, 64)])
def test_subset_all_shapes(emu, oracle, subset, queries, K, group_len):
    """K = 0 lets the planner pick the rows per lane per group size (the product default: one launch group per
    distinct K)"""
    m = oracle.matrix("blosum50")
    q = oracle.encode(queries["P02232"])
    want = oracle.scan(q, subset["codes"], subset["offsets"], m)
    got, _ = emu(subset["codes"], subset["offsets"], m, q, K=K, group_len=group_len)
    assert np.array_equal(got, want)


def test_query_chunks_and_forced_recompute(emu, oracle, subset, queries):
    m = oracle.matrix("blosum50")
    q = oracle.encode(queries["P04775"])  # 2005 rows -> two chunks of 1024
    want = oracle.scan(q, subset["codes"], subset["offsets"], m)
    got, rc = emu(subset["codes"], subset["offsets"], m, q, K=32, group_len=384, chunk_rows=1024)
    assert np.array_equal(got, want) and rc == 0
    # threshold override: every tile whose best exceeds 300 is re-scored by the int32 path
    got, rc = emu(subset["codes"], subset["offsets"], m, q, K=32, group_len=384, chunk_rows=1024, thr=300)
    assert np.array_equal(got, want) and rc >= 1
    got, rc = emu(subset["codes"], subset["offsets"], m, q, K=16, group_len=128, force_i32=1)
    assert np.array_equal(got, want)


def test_real_s16_overflow_is_caught(emu, oracle):
    """W x 2300 against itself scores 34500 > 32767: the s16 pass must flag it and int32 must fix it"""
    m = oracle.matrix("blosum50")
    w = np.full(2300, 17, dtype=np.uint8)
    rng = np.random.default_rng(2)
    enc = [w, rng.integers(0, 20, 300).astype(np.uint8), w[:2200].copy(), rng.integers(0, 20, 50).astype(np.uint8)]
    codes, offs = pack_db(enc)
    want = oracle.scan(w, codes, offs, m)
    assert want[0] == 34500 and want[2] == 33000
    got, rc = emu(codes, offs, m, w, K=0, group_len=384)
    assert np.array_equal(got, want) and rc >= 1


def test_edges_ident3_and_shards(emu, oracle, subset, queries):
    rng = np.random.default_rng(5)
    m = oracle.matrix("blosum50")
    lens = [0, 1, 2, 3, 4, 5, 7, 8, 9, 0, 33, 64, 65, 127, 128, 129, 1, 0, 300]
    enc = random_db(rng, lens)
    codes, offs = pack_db(enc)
    for ql in (1, 8, 9, 33):
        q = rng.integers(0, 24, ql).astype(np.uint8)
        want = oracle.scan(q, codes, offs, m)
        for K, gl in ((8, 8), (16, 16), (32, 384), (0, 16)):
            got, _ = emu(codes, offs, m, q, K=K, group_len=gl)
            assert np.array_equal(got, want), (ql, K, gl)
    got, _ = emu(codes, offs, m, np.zeros(0, np.uint8), K=32)
    assert np.array_equal(got, np.zeros(len(lens), np.int32))
    # +3/-3 scheme
    mi = oracle.matrix("ident3")
    ic, io = pack_db([oracle.encode(s, "ident3") for s in subset["seqs"][:40]])
    q = oracle.encode(queries["P02232"], "ident3")
    got, _ = emu(ic, io, mi, q, K=16, group_len=128)
    assert np.array_equal(got, oracle.scan(q, ic, io, mi))
    # shards: union of the per-shard outputs equals the unsharded scan
    q = oracle.encode(queries["P02232"])
    want = oracle.scan(q, subset["codes"], subset["offsets"], m)
    import importlib
    swb = importlib.import_module(PKG)
    merged = np.full(111, -1, np.int32)
    for s in range(3):
        info, _, ids = swb.plan_describe(subset["offsets"], s, 3, want_ids=True)
        got, _ = emu(subset["codes"], subset["offsets"], m, q, K=32, shard=s, nshards=3, n_out=info.n_local)
        merged[ids] = got
    assert np.array_equal(merged, want)


def test_pipelined_passes_of_very_long_tiles(emu, oracle):
    """xl_len: 32-lane tiles wider than it hand out their passes as separate, pipelined work items (boundary row
    through the scratch + progress counters, scores merged with atomicMax). In the emulation the items run one after
    the other in hand-out order, so a pass that had to wait for data that is not there yet would spin forever."""
    rng = np.random.default_rng(21)
    m = oracle.matrix("blosum50")
    lens = [1500, 1333, 900, 801, 640, 300, 280, 120, 64, 30, 7, 1200]
    enc = random_db(rng, lens, alphabet=20)
    codes, offs = pack_db(enc)
    for ql in (100, 256, 257, 700, 1100):
        q = rng.integers(0, 20, ql).astype(np.uint8)
        want = oracle.scan(q, codes, offs, m)
        for gl, xl in ((16, 600), (16, 100), (32, 1000)):
            got, _ = emu(codes, offs, m, q, K=0, group_len=gl, xl_len=xl)
            assert np.array_equal(got, want), (ql, gl, xl)
    # chunked query through split tiles, and an s16 overflow inside a split tile (flag -> int32 recompute of the tile)
    q = rng.integers(0, 20, 2300).astype(np.uint8)
    want = oracle.scan(q, codes, offs, m)
    got, _ = emu(codes, offs, m, q, K=0, group_len=16, xl_len=500, chunk_rows=1024)
    assert np.array_equal(got, want)
    w = np.full(2300, 17, dtype=np.uint8)
    c2, o2 = pack_db([w, w[:2250].copy(), enc[0], enc[1]])
    want = oracle.scan(w, c2, o2, m)
    got, rc = emu(c2, o2, m, w, K=0, group_len=16, xl_len=500)
    assert want[0] == 34500 and np.array_equal(got, want) and rc >= 1
    # the int32 recompute of split tiles is pipelined too (scores cleared first, combined with atomicMax): a low
    # threshold sends nearly every tile through it, over several query chunks; force_i32 runs it alone
    want = oracle.scan(q, codes, offs, m)
    got, rc = emu(codes, offs, m, q, K=0, group_len=16, xl_len=500, chunk_rows=1024, thr=40)
    assert np.array_equal(got, want) and rc >= 3
    got, _ = emu(codes, offs, m, q, K=0, group_len=16, xl_len=300, force_i32=1)
    assert np.array_equal(got, want)


def test_query_pair_jobs(emu, oracle, subset, queries):
    """V16Q: two queries in the halves of the s16x2 lanes against one DB sequence per lane; the tile is two work
    items (first / second sequence of every pair). Unequal lengths, chunked profile, overflow of one query only."""
    m = oracle.matrix("blosum50")
    codes, offs = subset["codes"], subset["offsets"]
    for na, nb, kw in (("P02232", "P05013", dict(K=0, group_len=384)), ("P01008", "P02232", dict(K=0, group_len=64)),
                       ("P14942", "P14942", dict(K=16, group_len=16)),
                       ("P27895", "P07327", dict(K=0, group_len=128, chunk_rows_pair=512))):
        qa, qb = oracle.encode(queries[na]), oracle.encode(queries[nb])
        (ga, gb), _ = emu(codes, offs, m, qa, q2=qb, **kw)
        assert np.array_equal(ga, oracle.scan(qa, codes, offs, m)), (na, nb)
        assert np.array_equal(gb, oracle.scan(qb, codes, offs, m)), (na, nb)
    rng = np.random.default_rng(8)
    w = np.full(2300, 17, dtype=np.uint8)
    enc = [w, rng.integers(0, 20, 300).astype(np.uint8), w[:2200].copy(), rng.integers(0, 20, 50).astype(np.uint8),
           np.zeros(0, np.uint8)]
    c2, o2 = pack_db(enc)
    qb = rng.integers(0, 20, 700).astype(np.uint8)
    (ga, gb), rc = emu(c2, o2, m, w, q2=qb, K=0, group_len=384)
    assert np.array_equal(ga, oracle.scan(w, c2, o2, m)) and ga[0] == 34500 and rc >= 1
    assert np.array_equal(gb, oracle.scan(qb, c2, o2, m))
    (ga, gb), _ = emu(c2, o2, m, qb[:9], q2=np.zeros(0, np.uint8), K=0)
    assert np.array_equal(ga, oracle.scan(qb[:9], c2, o2, m)) and not gb.any()


# ---- affine gaps ("define affine penalty ?", SWSolver.cu:8) -------------------------------------------
@pytest.fixture(scope="module")
def emu_affine(emu):
    L = ctypes.CDLL(os.path.join(ROOT, PKG, "lib", "libswbemu.so"))
    L.swbemu_search_affine.restype = ctypes.c_int
    L.swbemu_search_affine.argtypes = [_u8p, _u64p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                       _i8p, ctypes.c_int, ctypes.c_int, _u8p, ctypes.c_uint32, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_uint32, ctypes.c_int, _i32p,
                                       ctypes.POINTER(ctypes.c_uint32)]

    def search(codes, offs, m, q, go, ge, K=0, group_len=384, force_i32=0, chunk_rows=0, thr=-1):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        q = np.ascontiguousarray(q, dtype=np.uint8)
        m = np.ascontiguousarray(m, dtype=np.int8)
        n = len(offs) - 1
        out = np.full(n, -7, dtype=np.int32)
        rc = ctypes.c_uint32()
        r = L.swbemu_search_affine(codes.ctypes.data_as(_u8p), offs.ctypes.data_as(_u64p), n, 0, 1, group_len,
                                   m.ctypes.data_as(_i8p), go, ge, q.ctypes.data_as(_u8p), len(q), K, force_i32,
                                   chunk_rows, thr, out.ctypes.data_as(_i32p), ctypes.byref(rc))
        assert r == 0
        return out, rc.value

    return search


def test_oracle_affine_degenerates_to_linear(oracle, subset, queries):
    m = oracle.matrix("blosum50")
    q = oracle.encode(queries["P02232"])
    lin = oracle.scan(q, subset["codes"], subset["offsets"], m)
    assert np.array_equal(oracle.scan_affine(q, subset["codes"], subset["offsets"], m, 2, 2), lin)
    # a dearer opening can only lower a score, and never below the gap-free diagonal score
    aff = oracle.scan_affine(q, subset["codes"], subset["offsets"], m, 10, 2)
    assert (aff <= lin).all() and (aff < lin).any()
    assert (oracle.scan_affine(q, subset["codes"], subset["offsets"], m, 64, 64) <= aff).all()


@pytest.mark.parametrize("go,ge,K,group_len", [(10, 2, 0, 384), (10, 2, 8, 64), (12, 1, 16, 16), (5, 0, 0, 16), (11, 1, 32, 96),
                                               (3, 2, 8, 2000)])
def test_affine_subset(emu_affine, oracle, subset, queries, go, ge, K, group_len):
    m = oracle.matrix("blosum50")
    q = oracle.encode(queries["P02232"])
    want = oracle.scan_affine(q, subset["codes"], subset["offsets"], m, go, ge)
    got, _ = emu_affine(subset["codes"], subset["offsets"], m, q, go, ge, K=K, group_len=group_len)
    assert np.array_equal(got, want)


def test_affine_chunks_recompute_and_overflow(emu_affine, oracle, subset, queries):
    m = oracle.matrix("blosum50")
    q = oracle.encode(queries["P04775"])[:1100]  # two chunks of 1024 rows
    n = 36  # the first 36 sequences of the subset keep the emulation short
    offs = subset["offsets"][:n + 1]
    codes = subset["codes"][:int(offs[-1])]
    want = oracle.scan_affine(q, codes, offs, m, 10, 2)
    got, rc = emu_affine(codes, offs, m, q, 10, 2, chunk_rows=1024)
    assert np.array_equal(got, want) and rc == 0
    got, rc = emu_affine(codes, offs, m, q, 10, 2, chunk_rows=1024, thr=60, group_len=128)
    assert np.array_equal(got, want) and rc >= 1
    got, rc = emu_affine(codes, offs, m, q, 10, 2, force_i32=1, group_len=64)
    assert np.array_equal(got, want)
    # a real s16 overflow (a matrix with W:W = 100 gets there with a short sequence; the GPU test uses BLOSUM50)
    m2 = m.copy()
    m2[17, 17] = 100
    w = np.full(400, 17, dtype=np.uint8)
    rng = np.random.default_rng(2)
    codes, offs = pack_db([w, rng.integers(0, 20, 300).astype(np.uint8), w[:390].copy()])
    want = oracle.scan_affine(w, codes, offs, m2, 10, 2)
    assert want[0] == 40000 and want[2] == 39000
    got, rc = emu_affine(codes, offs, m2, w, 10, 2)
    assert np.array_equal(got, want) and rc >= 1


def test_affine_edges(emu_affine, oracle):
    rng = np.random.default_rng(8)
    m = oracle.matrix("blosum50")
    lens = [0, 1, 2, 3, 4, 5, 7, 8, 9, 0, 33, 64, 65, 127, 128, 129, 1, 0, 300, 700]
    codes, offs = pack_db(random_db(rng, lens))
    for ql in (1, 8, 9, 33, 130):
        # low-complexity queries make gapped paths win often
        q = rng.integers(0, 4, ql).astype(np.uint8)
        for go, ge in ((10, 2), (4, 1), (2, 0)):
            want = oracle.scan_affine(q, codes, offs, m, go, ge)
            for K, gl in ((8, 8), (16, 16), (0, 384)):
                got, _ = emu_affine(codes, offs, m, q, go, ge, K=K, group_len=gl)
                assert np.array_equal(got, want), (ql, go, ge, K, gl)


def test_pipelined_passes_with_full_blocks(emu, oracle):
    """split launches with enough work items run K = 16 / 32 strips over the block-staged chunk (the engine raises K
    while the items still fill the GPU's warp slots; here split_fill = 1 always picks the largest K that applies)"""
    L = ctypes.CDLL(os.path.join(ROOT, PKG, "lib", "libswbemu.so"))
    L.swbemu_search_split.restype = ctypes.c_int
    L.swbemu_search_split.argtypes = [_u8p, _u64p, ctypes.c_uint32, ctypes.c_uint32, _i8p, ctypes.c_int, _u8p,
                                      ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32,
                                      ctypes.c_uint32, _i32p, ctypes.POINTER(ctypes.c_uint32)]

    def run(codes, offs, m, q, group_len, xl_len, fill, chunk_rows=0, thr=-1, force_i32=0):
        out = np.full(len(offs) - 1, -7, dtype=np.int32)
        rc = ctypes.c_uint32()
        r = L.swbemu_search_split(codes.ctypes.data_as(_u8p), offs.ctypes.data_as(_u64p), len(offs) - 1, group_len,
                                  np.ascontiguousarray(m, dtype=np.int8).ctypes.data_as(_i8p), 2,
                                  q.ctypes.data_as(_u8p), len(q), force_i32, chunk_rows, thr, xl_len, fill,
                                  out.ctypes.data_as(_i32p), ctypes.byref(rc))
        assert r == 0
        return out, rc.value

    rng = np.random.default_rng(23)
    m = oracle.matrix("blosum50")
    lens = [1500, 1333, 900, 801, 640, 300, 280, 120, 64, 30, 7, 1200]
    codes, offs = pack_db(random_db(rng, lens, alphabet=20))
    for ql in (100, 700, 1100):
        q = rng.integers(0, 20, ql).astype(np.uint8)
        want = oracle.scan(q, codes, offs, m)
        for gl, xl, fill in ((16, 100, 1), (16, 600, 40), (32, 1000, 1)):
            got, _ = run(codes, offs, m, q, gl, xl, fill)
            assert np.array_equal(got, want), (ql, gl, xl, fill)
    q = rng.integers(0, 20, 2300).astype(np.uint8)
    want = oracle.scan(q, codes, offs, m)
    got, rc = run(codes, offs, m, q, 16, 500, 1, chunk_rows=1024, thr=40)  # chunks + int32 recompute (K = 16 split)
    assert np.array_equal(got, want) and rc >= 3
    got, _ = run(codes, offs, m, q, 16, 300, 1, force_i32=1)
    assert np.array_equal(got, want)


def test_pipelined_passes_every_lane_group_class(oracle):
    """a split set that reaches down to 2 lanes per pair: tiles with several pairs per warp (slots) publish one
    progress value for all of them, items map to (tile, pass) through the per-class tables"""
    L = ctypes.CDLL(os.path.join(ROOT, PKG, "lib", "libswbemu.so"))
    L.swbemu_search_split.restype = ctypes.c_int
    L.swbemu_search_split.argtypes = [_u8p, _u64p, ctypes.c_uint32, ctypes.c_uint32, _i8p, ctypes.c_int, _u8p,
                                      ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32,
                                      ctypes.c_uint32, _i32p, ctypes.POINTER(ctypes.c_uint32)]
    m = oracle.matrix("blosum50")
    rng = np.random.default_rng(99)
    lens = [17, 18, 20, 25, 31, 32, 33, 40, 60, 64, 65, 100, 128, 129, 200, 256, 257, 300, 500, 700, 5, 9, 12, 16]
    rng.shuffle(lens)
    codes, offs = pack_db(random_db(rng, lens, alphabet=20))
    mm = np.ascontiguousarray(m, dtype=np.int8)
    for ql in (1, 9, 33, 257):
        q = rng.integers(0, 20, ql).astype(np.uint8)
        want = oracle.scan(q, codes, offs, m)
        for gl, xl in ((16, 16), (16, 40), (8, 8), (32, 32)):
            for thr in (-1, 30):  # 30: most tiles also go through the pipelined int32 recompute
                out = np.full(len(offs) - 1, -7, dtype=np.int32)
                rc = ctypes.c_uint32()
                r = L.swbemu_search_split(codes.ctypes.data_as(_u8p), offs.ctypes.data_as(_u64p), len(offs) - 1, gl,
                                          mm.ctypes.data_as(_i8p), 2, q.ctypes.data_as(_u8p), len(q), 0, 0, thr, xl, 0,
                                          out.ctypes.data_as(_i32p), ctypes.byref(rc))
                assert r == 0 and np.array_equal(out, want), (ql, gl, xl, thr)


def test_randomized_configurations(emu, oracle):
    """40 random (database, query, group_len, K, xl_len, overflow threshold) draws through the host emulation of the
    warp program, each against the oracle (fixed seed: the draw is part of the test)"""
    rng = np.random.default_rng(20261018)
    m = oracle.matrix("blosum50")
    for it in range(40):
        nseq = int(rng.integers(1, 40))
        top = int(rng.choice([12, 60, 300, 900]))
        lens = rng.integers(0, top, nseq)
        codes, offs = pack_db(random_db(rng, lens, alphabet=int(rng.choice([4, 20, 25]))))
        q = rng.integers(0, 24, int(rng.choice([1, 5, 8, 31, 64, 130, 400]))).astype(np.uint8)
        gl = int(rng.choice([8, 16, 32, 64, 384]))
        K = int(rng.choice([0, 8, 16, 32]))
        xl = int(rng.choice([0, 16, 64, 256, 8192]))
        thr = int(rng.choice([-1, -1, 10, 40, 200]))
        want = oracle.scan(q, codes, offs, m)
        got, _ = emu(codes, offs, m, q, K=K, group_len=gl, xl_len=xl, thr=thr)
        assert np.array_equal(got, want), (it, nseq, top, len(q), gl, K, xl, thr)
