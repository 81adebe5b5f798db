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
    import emu_lib
    emu_lib.load()
    return emu_lib.search


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


# ---- affine gaps ("define affine penalty ?", SWSolver.cu:8) -------------------------------------------
@pytest.fixture(scope="module")
def emu_affine(emu):
    def search(codes, offs, m, q, go, ge, K=0, group_len=384, force_i32=0, chunk_rows=0, thr=-1):
        return emu(codes, offs, m, q, K=K, group_len=group_len, force_i32=force_i32, chunk_rows=chunk_rows, thr=thr,
                   gap=go, gap_extend=ge, xl_len=0)

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


def test_pipelined_passes_with_16_row_strips(emu, oracle):
    """split groups with 16 rows per lane (the engine picks them while the items still fill the GPU): s16 pass, the
    exact recompute in both forms (rebased s16 and int32), several query chunks"""
    rng = np.random.default_rng(23)
    m = oracle.matrix("blosum50")
    lens = [1500, 1333, 900, 801, 640, 300, 280, 120, 64, 30, 7, 1200]
    codes, offs = pack_db(random_db(rng, lens, alphabet=20))
    for ql in (100, 700, 1100):
        q = rng.integers(0, 20, ql).astype(np.uint8)
        want = oracle.scan(q, codes, offs, m)
        for gl, xl, sk in ((16, 100, 16), (16, 600, 16), (32, 1000, 16), (16, 300, 32)):
            got, _ = emu(codes, offs, m, q, K=0, group_len=gl, xl_len=xl, split_k=sk)
            assert np.array_equal(got, want), (ql, gl, xl, sk)
    q = rng.integers(0, 20, 2300).astype(np.uint8)
    want = oracle.scan(q, codes, offs, m)
    for exact_i32 in (0, 1):
        got, rc = emu(codes, offs, m, q, K=0, group_len=16, xl_len=500, chunk_rows=1024, thr=40, split_k=16,
                      exact_i32=exact_i32, rebase_shift=6)
        assert np.array_equal(got, want) and rc >= 3, exact_i32
        got, _ = emu(codes, offs, m, q, K=0, group_len=16, xl_len=300, force_i32=1, split_k=16, exact_i32=exact_i32,
                     rebase_shift=7)
        assert np.array_equal(got, want), exact_i32
    got, rc = emu(codes, offs, m, q, K=0, group_len=16, xl_len=500, thr=40, split_k=32, direct_len=1000, rebase_shift=6)
    assert np.array_equal(got, want)


def test_pipelined_passes_every_lane_group_class(emu, oracle):
    """a split set that reaches down to 2 lanes per pair: tiles with several pairs per warp (slots) publish one
    progress value for all of them, items map to (tile, pass) through the per-class tables"""
    m = oracle.matrix("blosum50")
    rng = np.random.default_rng(99)
    lens = [17, 18, 20, 25, 31, 32, 33, 40, 60, 64, 65, 100, 128, 129, 200, 256, 257, 300, 500, 700, 5, 9, 12, 16]
    rng.shuffle(lens)
    codes, offs = pack_db(random_db(rng, lens, alphabet=20))
    for ql in (1, 9, 33, 257):
        q = rng.integers(0, 20, ql).astype(np.uint8)
        want = oracle.scan(q, codes, offs, m)
        for gl, xl in ((16, 16), (16, 40), (8, 8), (32, 32)):
            for thr in (-1, 30):  # 30: most tiles also go through the pipelined exact recompute
                for exact_i32 in (0, 1):
                    got, _ = emu(codes, offs, m, q, K=0, group_len=gl, xl_len=xl, thr=thr, exact_i32=exact_i32,
                                 rebase_shift=6)
                    assert np.array_equal(got, want), (ql, gl, xl, thr, exact_i32)


def test_rebased_s16_is_exact_far_beyond_the_s16_range(emu, oracle):
    """V16R (s16x2 relative to a base that follows the columns): the exact pass over every tile and the "direct" path
    for long-against-long tiles, with true scores of 66,000 (4,400 W-W matches at 15: beyond even an unsigned 16-bit
    range; the GPU suite repeats this at 151,500), several blocks per pass, several passes per tile, pairs of very
    different length, one lane per pair up to 32 lanes per pair"""
    m = oracle.matrix("blosum50")
    rng = np.random.default_rng(5)
    w = np.full(4400, 17, dtype=np.uint8)  # W: 15 per match
    noisy = w.copy()
    noisy[rng.choice(len(w), 300, replace=False)] = rng.integers(0, 20, 300)
    small = [rng.integers(0, 20, 1800).astype(np.uint8), w[:700].copy(), rng.integers(0, 20, 90).astype(np.uint8),
             np.zeros(0, np.uint8)]
    codes, offs = pack_db([w, noisy] + small)
    want = oracle.scan(w, codes, offs, m)
    assert want[0] == 66000 and want[1] > 45000
    # direct: the long tiles never see the plain s16 pass (nothing is flagged there); the short ones go the normal way
    got, rc = emu(codes, offs, m, w, K=0, group_len=128, xl_len=1000, split_k=16, direct_len=1500, rebase_shift=8)
    assert np.array_equal(got, want)
    # flagged -> rebased recompute, pipelined (split_k 8) and not (xl_len 0), small blocks
    q = w[:2400]
    codes, offs = pack_db([q.copy(), noisy[:2300].copy()] + small)
    wq = oracle.scan(q, codes, offs, m)
    assert wq.max() == 36000
    got, rc = emu(codes, offs, m, q, K=0, group_len=128, xl_len=1000, split_k=8, rebase_shift=7)
    assert np.array_equal(got, wq) and rc >= 1
    got, rc = emu(codes, offs, m, q, K=0, group_len=512, xl_len=0, rebase_shift=6)
    assert np.array_equal(got, wq) and rc >= 1
    # one lane per pair through the rebased policy (its G = 1 case): group_len above every sequence
    c3, o3 = pack_db(small)
    got, rc = emu(c3, o3, m, q, K=0, group_len=16384, xl_len=0, force_i32=1, rebase_shift=8)
    assert np.array_equal(got, oracle.scan(q, c3, o3, m))
    c1, o1 = pack_db([w[:2400].copy(), w[:2300].copy()])
    got, rc = emu(c1, o1, m, q, K=0, group_len=16384, xl_len=0, rebase_shift=6)  # flagged one-lane tile -> V16R
    assert np.array_equal(got, oracle.scan(q, c1, o1, m)) and got[0] == 36000 and rc == 1


def test_rebased_s16_padding_lanes_at_the_clamped_floor(emu, oracle):
    """Regression (found on the configs[3] workload): behind the last column of a tile the first lane keeps running on
    padding with "zero" from above; once the base has passed 32000 that zero is the clamped floor, and a rebase inside
    the padding used to wrap it into a huge positive value that reached the running maximum (score + ~32000). The pair
    below is two configs[3] targets of ~21,400 residues against the first 30,000 rows of the 35,213-row query."""
    import bench
    codes, offs, qs = bench.synth_config4()
    m = oracle.matrix("blosum50")
    c2, o2 = pack_db([codes[int(offs[i]):int(offs[i + 1])] for i in (68, 1)])
    q = qs[3][:30000]
    want = oracle.scan(q, c2, o2, m)
    assert want.min() > 32767
    got, _ = emu(c2, o2, m, q, K=0, group_len=384, xl_len=3072, split_k=16, direct_len=3000)
    assert np.array_equal(got, want)


def test_rebased_s16_random_and_ident3(emu, oracle):
    """V16R as the only pass (force) over random databases, both scoring presets, every K / group size mix"""
    rng = np.random.default_rng(77)
    for preset, alphabet in (("blosum50", 20), ("ident3", 4)):
        m = oracle.matrix(preset)
        lens = [900, 640, 333, 300, 280, 120, 64, 30, 7, 1, 0, 450]
        codes, offs = pack_db(random_db(rng, lens, alphabet=alphabet))
        for ql in (1, 40, 333, 1030):
            q = rng.integers(0, alphabet, ql).astype(np.uint8)
            want = oracle.scan(q, codes, offs, m)
            for K, gl, xl, shift in ((0, 64, 0, 6), (8, 16, 200, 6), (16, 32, 64, 7), (0, 384, 0, 6)):
                got, _ = emu(codes, offs, m, q, K=K, group_len=gl, xl_len=xl, force_i32=1, rebase_shift=shift)
                assert np.array_equal(got, want), (preset, ql, K, gl, xl, shift)


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
        sk = int(rng.choice([8, 16, 32]))
        ex = int(rng.choice([0, 0, 1]))
        dl = int(rng.choice([0, 0, 50, 300]))
        want = oracle.scan(q, codes, offs, m)
        got, _ = emu(codes, offs, m, q, K=K, group_len=gl, xl_len=xl, thr=thr, split_k=sk, exact_i32=ex, direct_len=dl,
                     rebase_shift=6)
        assert np.array_equal(got, want), (it, nseq, top, len(q), gl, K, xl, thr, sk, ex, dl)


@pytest.mark.parametrize("variant", ["", "_blk"])
def test_one_lane_tiles_in_column_blocks_and_pass_groups(oracle, subset, queries, variant):
    """One-lane-per-pair tiles run SWB_PASS_GROUP passes together over column blocks of SWB_BLOCK_CHUNKS chunks, parking
    their row state between blocks (swb_run_tile). Checked with the product's parameters (4 passes x 16 chunks) and with
    a build of the same source that uses 3 x 3, so that every border case occurs on small inputs: widths below, at and
    above a block, pass counts that are not a multiple of the group, query chunks, the int32 / affine policies (their
    parked elements are 8 and 16 bytes), a real s16 overflow re-scored by V32."""
    import emu_lib
    emu_lib.load(variant)
    m = oracle.matrix("blosum50")
    rng = np.random.default_rng(23)
    lens = [0, 1, 3, 4, 5, 11, 12, 13, 23, 24, 25, 47, 48, 49, 95, 96, 97, 130, 191, 192, 193, 260, 401, 777]
    codes, offs = pack_db(random_db(rng, lens))
    for ql in (1, 31, 32, 33, 96, 97, 130, 257):
        q = rng.integers(0, 24, ql).astype(np.uint8)
        want = oracle.scan(q, codes, offs, m)
        for K in (8, 16, 32):
            got, _ = emu_lib.search(codes, offs, m, q, K=K, group_len=4096, variant=variant)
            assert np.array_equal(got, want), (variant, ql, K)
        got, _ = emu_lib.search(codes, offs, m, q, K=16, group_len=4096, force_i32=1, exact_i32=1, variant=variant)
        assert np.array_equal(got, want), (variant, ql, "V32")
        wa = oracle.scan_affine(q, codes, offs, m, 10, 2)
        got, _ = emu_lib.search(codes, offs, m, q, K=0, group_len=4096, gap=10, gap_extend=2, variant=variant)
        assert np.array_equal(got, wa), (variant, ql, "V16A")
        got, _ = emu_lib.search(codes, offs, m, q, K=8, group_len=4096, gap=10, gap_extend=2, force_i32=1, variant=variant)
        assert np.array_equal(got, wa), (variant, ql, "V32A")
    # query chunks (the boundary row of a chunk is the top row of the next launch) with one-lane tiles only
    q = oracle.encode(queries["P04775"])  # 2005 rows -> two chunks of 1024
    want = oracle.scan(q, subset["codes"], subset["offsets"], m)
    got, rc = emu_lib.search(subset["codes"], subset["offsets"], m, q, K=32, group_len=4096, chunk_rows=1024, variant=variant)
    assert np.array_equal(got, want) and rc == 0
    # a real s16 overflow in a one-lane tile, re-scored by the int32 policy
    w = np.full(2300, 17, dtype=np.uint8)
    codes, offs = pack_db([w, rng.integers(0, 20, 300).astype(np.uint8), w[:2200].copy()])
    want = oracle.scan(w, codes, offs, m)
    assert want[0] == 34500
    got, rc = emu_lib.search(codes, offs, m, w, K=0, group_len=4096, exact_i32=1, variant=variant)
    assert np.array_equal(got, want) and rc >= 1


def test_running_maximum_over_diagonal_sums(emu, emu_affine, oracle):
    """The s16 cells keep max(d) with d = H(i-1,j-1) + S instead of max(H) (csrc/swb_warp.cuh, SWB_V16_FORM 1): a maximal
    H is never the end of a gap. Structured cases where the best cell sits in the first row / first column, right behind
    a gap, or nowhere (score 0), linear and affine, every strip height and lane-group size."""
    m = oracle.matrix("blosum50")
    A, W, C, P = 0, 17, 4, 14
    q = np.array([W] * 6 + [A] * 3 + [C] * 5, dtype=np.uint8)
    enc = [
        np.array([W], dtype=np.uint8),                                   # best in the first column
        np.array([P] * 40 + [W], dtype=np.uint8),                        # best in the last column, first query row
        np.array([W] * 6 + [P] * 2 + [C] * 5, dtype=np.uint8),           # gap in the target between two blocks
        np.array([W] * 6 + [C] * 5, dtype=np.uint8),                     # gap in the query
        np.array([P] * 70, dtype=np.uint8),                              # nothing aligns: 0
        np.array([C] * 5 + [P] * 30 + [W] * 6 + [A] * 3 + [C] * 5 + [P] * 9, dtype=np.uint8),
        np.zeros(0, dtype=np.uint8),
    ]
    rng = np.random.default_rng(11)
    enc += [np.concatenate([rng.integers(0, 20, int(n)).astype(np.uint8), q[::-1], rng.integers(0, 20, 3).astype(np.uint8)])
            for n in (0, 5, 130, 700)]
    codes, offs = pack_db(enc)
    want = oracle.scan(q, codes, offs, m)
    assert want[0] == 15 and want[4] == 0 and want[2] > want[3] > 0
    for K, gl in ((8, 8), (16, 16), (32, 384), (0, 64), (32, 32)):
        got, _ = emu(codes, offs, m, q, K=K, group_len=gl)
        assert np.array_equal(got, want), (K, gl)
    for go, ge in ((10, 2), (3, 1)):
        wanta = oracle.scan_affine(q, codes, offs, m, go, ge)
        for K, gl in ((8, 8), (16, 64), (32, 384)):
            got, _ = emu_affine(codes, offs, m, q, go, ge, K=K, group_len=gl)
            assert np.array_equal(got, wanta), (go, ge, K, gl)
