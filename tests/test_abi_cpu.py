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
