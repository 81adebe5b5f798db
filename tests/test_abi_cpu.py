"""CPU-side checks of the product library: it loads, exports every symbol include/swb.h declares, its
presets equal the oracle's, the database plan (sort / shard / tile) is sane, and it refuses to run
without a GPU instead of falling back."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_exports_match_header(swb):
    header = open(os.path.join(ROOT, "include", "swb.h")).read()
    declared = sorted(set(re.findall(r"\b(swb_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(swb.ABI_SYMBOLS)
    L = swb.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_presets_equal_oracle(swb, oracle):
    m, gap = swb.scoring_matrix(swb.SWB_SCORING_BLOSUM50_REF)
    assert gap == 2 and np.array_equal(m, oracle.matrix("blosum50"))
    m, gap = swb.scoring_matrix(swb.SWB_SCORING_IDENT3)
    assert gap == 2 and np.array_equal(m, oracle.matrix("ident3"))
    text = bytes(range(1, 256))
    assert np.array_equal(swb.encode(text, swb.SWB_SCORING_BLOSUM50_REF), oracle.encode(text, "blosum50"))
    assert np.array_equal(swb.encode(text, swb.SWB_SCORING_IDENT3), oracle.encode(text, "ident3"))


def test_no_cpu_fallback(swb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(swb.SwbError) as ei:
        swb.Engine(0)
    assert "no CPU fallback" in str(ei.value)


def _offsets(lens):
    o = np.zeros(len(lens) + 1, dtype=np.uint64)
    o[1:] = np.cumsum(lens)
    return o


def test_plan_sort_and_tiles(swb):
    rng = np.random.default_rng(11)
    lens = np.clip(np.round(rng.lognormal(5.58, 0.75, 20000)), 2, 35213).astype(np.int64)
    lens[:3] = [35213, 0, 1]
    offs = _offsets(lens)
    info, sorted_ids, shard_ids = swb.plan_describe(offs, want_ids=True)
    assert info.n_total == info.n_local == len(lens) and info.max_len == 35213
    assert info.residues_local == info.residues_total == lens.sum()
    sl = lens[sorted_ids]
    assert (np.diff(sl) <= 0).all()  # longest first
    # stable: equal lengths keep database order
    for L in np.unique(sl)[:50]:
        ids = sorted_ids[sl == L]
        assert (np.diff(ids.astype(np.int64)) > 0).all()
    assert np.array_equal(shard_ids, np.arange(len(lens), dtype=np.uint32))
    assert sum(info.tiles_by_group) == info.tiles and info.tiles_by_group[5] >= 1
    assert info.res_bytes == info.padded_cols and info.bnd_elems * 2 == info.padded_cols
    # padding overhead of the tiling stays small on a Swiss-Prot-shaped length mix
    assert info.padded_cols < 1.06 * lens.sum()
    # group_len large enough -> every tile is one lane per pair
    info2 = swb.plan_describe(offs, group_len=1 << 20)
    assert info2.tiles_by_group[0] == info2.tiles


@pytest.mark.parametrize("nshards", [2, 3, 8])
def test_plan_sharding_is_balanced_partition(swb, nshards):
    rng = np.random.default_rng(5)
    lens = np.clip(np.round(rng.lognormal(5.58, 0.75, 30001)), 2, 35213).astype(np.int64)
    offs = _offsets(lens)
    seen = np.zeros(len(lens), dtype=int)
    res = []
    for s in range(nshards):
        info, sorted_ids, shard_ids = swb.plan_describe(offs, s, nshards, want_ids=True)
        seen[shard_ids] += 1
        assert (np.diff(shard_ids.astype(np.int64)) > 0).all()
        assert sorted(sorted_ids.tolist()) == shard_ids.tolist()
        res.append(info.residues_local)
        assert info.residues_local == lens[shard_ids].sum()
    assert (seen == 1).all()  # every sequence in exactly one shard
    assert (max(res) - min(res)) <= 2 * lens.max()  # residue-balanced to within one pair
    assert abs(max(res) / (lens.sum() / nshards) - 1) < 0.01


def test_plan_rejects_bad_input(swb):
    bad = np.array([0, 10, 5], dtype=np.uint64)
    with pytest.raises(swb.SwbError):
        swb.plan_describe(bad)
    info = swb.plan_describe(np.array([0], dtype=np.uint64))
    assert info.n_local == 0 and info.tiles == 0


def test_group_refuses_without_gpu(swb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(swb.SwbError) as ei:
        swb.EngineGroup()
    assert "no CPU fallback" in str(ei.value)


def test_layout_rules(swb):
    """the P x R layout of the engine group / of bench.py's ranks: P = the largest divisor of the device count that
    keeps min_part sequences per part; query groups by longest-processing-time first"""
    assert swb.layout_parts(570065, 1) == 1
    assert swb.layout_parts(570065, 2) == 1
    assert swb.layout_parts(570065, 8) == 1
    assert swb.layout_parts(2 * 570065, 8) == 2
    assert swb.layout_parts(5700650, 8) == 8
    assert swb.layout_parts(100, 8) == 1
    assert swb.layout_parts(900, 6, 300) == 3
    # the 20 reference queries over 4 groups: within half a percent of equal
    lens = [144, 189, 222, 375, 464, 567, 657, 729, 850, 1000, 1500, 2005, 2504, 3005, 3564, 4061, 4548, 4743, 5147, 5478]
    rng = np.random.default_rng(1)
    rng.shuffle(lens)
    offs = _offsets(lens)
    for groups in (1, 2, 4):
        g = swb.layout_query_groups(offs, groups)
        load = np.bincount(g, weights=np.array(lens, dtype=float), minlength=groups)
        assert len(g) == 20 and g.max() < groups
        assert load.max() / (sum(lens) / groups) < 1.005, (groups, load)
    # 1,000 queries of the configs[4] length law over 8 groups
    qlens = np.clip(np.round(np.random.default_rng(1785).lognormal(5.58, 0.75, 1000)), 30, 5478).astype(np.int64)
    g = swb.layout_query_groups(_offsets(qlens), 8)
    load = np.bincount(g, weights=qlens.astype(float), minlength=8)
    assert load.max() / load.mean() < 1.001
    # with the batch in view: 20 reference queries fill 2 or 4 groups evenly, not 8 (the longest query alone is 5 % above
    # an eighth of the rows) -> the database is split eight ways instead; 1,000 queries fill 8 groups
    # ... and four groups would hold five queries each, too few to fill the tails of each other's launches
    assert swb.layout_parts(570065, 2, qoffsets=offs) == 1
    assert swb.layout_parts(570065, 4, qoffsets=offs) == 4
    assert swb.layout_parts(570065, 8, qoffsets=offs) == 8
    assert swb.layout_parts(570065, 4, qoffsets=_offsets(list(lens) * 2)) == 1
    assert swb.layout_parts(570065, 8, qoffsets=_offsets(qlens)) == 1
    assert swb.layout_parts(5700650, 8, qoffsets=_offsets(qlens)) == 8
    assert swb.layout_parts(570065, 8, qoffsets=_offsets([100])) == 8
    # more groups than queries: the extra groups stay empty, nothing is lost
    g = swb.layout_query_groups(_offsets([5, 9]), 4)
    assert sorted(g.tolist()) == [0, 1] or len(set(g.tolist())) == 2
    assert len(swb.layout_query_groups(_offsets([]), 3)) == 0


def test_rebase_block_of_the_reference_scheme():
    """V16R's exactness window for BLOSUM50 / gap 2 (swb_warp.cuh): neighbouring cells differ by at most
    maxS + g = 17, so a pass of 512 rows over a block of 1,024 columns spans (512 + 1024 + 2) * 17 = 26,146 < 31,000
    -- the rule swb_rebase_shift implements (restated here; the emulation tests run the policy itself)"""
    max_s, min_s, gap, rows = 15, -5, 2, 512
    step = max_s + gap
    cols = (31000 - (-min_s) - gap - max_s) // step - rows - 2
    shift = 0
    while (2 << shift) <= cols:
        shift += 1
    assert shift == 10 and (rows + (1 << shift) + 2) * step + 5 + 2 + 15 <= 31000
