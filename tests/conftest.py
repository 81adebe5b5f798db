import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "ece1782-smith-waterman-cuda_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def swb():
    mod = importlib.import_module(PKG)
    if not os.path.exists(mod.LIB_PATH):
        mod.build()
    return mod


@pytest.fixture(scope="session")
def subset(oracle):
    """The 111-entry Swiss-Prot subset of the reference (tests/golden/uniprot_subset.fasta)."""
    from oracle_lib import pack_db, read_fasta
    heads, seqs = read_fasta(os.path.join(GOLDEN, "uniprot_subset.fasta"))
    codes, offsets = pack_db([oracle.encode(s) for s in seqs])
    return {"heads": heads, "seqs": seqs, "codes": codes, "offsets": offsets}


@pytest.fixture(scope="session")
def queries():
    from oracle_lib import read_query
    qdir = os.path.join(GOLDEN, "queries")
    return {fn[:-6]: read_query(os.path.join(qdir, fn)) for fn in sorted(os.listdir(qdir))}


@pytest.fixture(scope="session")
def survey_exp():
    return json.load(open(os.path.join(GOLDEN, "survey_exp_blosum50.json")))


@pytest.fixture(scope="session")
def engine(swb):
    """One engine shared by the GPU tests."""
    e = swb.Engine(0)
    yield e
    e.close()


def random_db(rng, lens, alphabet=25):
    return [rng.integers(0, alphabet, size=int(l)).astype(np.uint8) for l in lens]
