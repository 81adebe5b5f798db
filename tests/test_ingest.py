"""Database ingest and the encoded on-disk database (CPU only): swb_read_fasta must cut records exactly like the
reference parser (checked against the expectations printed by /root/reference/src/FASTAParsers.h itself, minus its '/'
padding), swb_read_uniprot_dat must find the SQ blocks the reference's parse.py recipe finds, and a written
database file must map back bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, PKG, ROOT
from oracle_lib import read_fasta

EXPECT = os.path.join(GOLDEN, "parser_expect")


def _expected_records(name):
    """(ids, sequences without '/' padding) in FILE order from the reference parser's dump"""
    lines = open(os.path.join(EXPECT, name + ".db.txt"), "rb").read().decode("latin-1").split("\n")[4:]
    recs = []
    for l in lines:
        if not l.strip() and l == "":
            continue
        sid, plen, seq = (l.split(" ", 2) + [""])[:3]
        recs.append((int(sid), seq.rstrip("/") if seq.endswith("/") else seq))
    recs.sort(key=lambda r: r[0])
    return recs


@pytest.mark.parametrize("name", ["uniprot_subset.fasta", "test.dat", "tricky.fasta", "empty.fasta", "noeol.fasta"])
def test_read_fasta_matches_reference_parser(swb, name):
    codes, offs, first_id = swb.read_fasta(os.path.join(GOLDEN, name))
    recs = _expected_records(name)
    assert len(offs) - 1 == len(recs)
    assert first_id == recs[0][0]
    for k, (sid, seq) in enumerate(recs):
        assert sid == first_id + k
        got = codes[int(offs[k]):int(offs[k + 1])]
        want = swb.encode(seq)
        # the reference pads with '/', which it cannot tell from a real trailing '/' either: compare up to padding
        assert np.array_equal(got[:len(want)], want) and (got[len(want):] == 24).all(), (name, k)


def test_read_fasta_missing_file_is_one_empty_record(swb):
    codes, offs, first_id = swb.read_fasta("/nonexistent/file.fasta")
    assert len(codes) == 0 and offs.tolist() == [0, 0] and first_id == -1


def test_read_uniprot_dat(swb):
    codes, offs = swb.read_uniprot_dat(os.path.join(GOLDEN, "uniprot_head5.dat"))
    _, seqs = read_fasta(os.path.join(GOLDEN, "uniprot_subset.fasta"))
    assert len(offs) - 1 == 5  # the fifth entry has no trailing "//"
    for k in range(5):
        assert np.array_equal(codes[int(offs[k]):int(offs[k + 1])], swb.encode(seqs[k])), k
    with pytest.raises(swb.SwbError):
        swb.read_uniprot_dat("/nonexistent/file.dat")


def test_dbfile_roundtrip_and_rejects_garbage(swb, tmp_path):
    rng = np.random.default_rng(3)
    lens = [0, 5, 1, 300, 0, 17]
    codes = rng.integers(0, 25, sum(lens)).astype(np.uint8)
    offs = np.zeros(len(lens) + 1, np.uint64)
    offs[1:] = np.cumsum(lens)
    path = str(tmp_path / "db.swbdb")
    swb.dbfile_write(path, codes, offs, first_id=0)
    c2, o2, fid = swb.dbfile_read(path)
    assert np.array_equal(c2, codes) and np.array_equal(o2, offs) and fid == 0
    assert os.path.getsize(path) == 32 + 8 * (len(lens) + 1) + len(codes)
    swb.dbfile_write(path, np.zeros(0, np.uint8), np.array([0], np.uint64), first_id=-1)
    c3, o3, fid = swb.dbfile_read(path)
    assert len(c3) == 0 and o3.tolist() == [0] and fid == -1
    bad = str(tmp_path / "bad.swbdb")
    open(bad, "wb").write(b"not a database file, definitely longer than the header")
    with pytest.raises(swb.SwbError):
        swb.dbfile_read(bad)
    with pytest.raises(swb.SwbError):
        swb.dbfile_read(str(tmp_path / "missing.swbdb"))
    # a crafted header whose sizes wrap around to the file size (n = 2^28, residues chosen so that
    # 32 + 8 * (n + 1) + residues == 40 modulo 2^64) must be rejected, not dereferenced
    import struct
    n_evil = 0x10000000
    residues = (40 - 32 - 8 * (n_evil + 1)) % (1 << 64)
    open(bad, "wb").write(b"SWBDB\x00\x01\x00" + struct.pack("<IiQQ", n_evil, 0, residues, 0) + b"\0" * 8)
    assert os.path.getsize(bad) == 40
    with pytest.raises(swb.SwbError):
        swb.dbfile_read(bad)
    # offsets that decrease or run past the residues are rejected too
    swb.dbfile_write(path, codes, offs, first_id=0)
    raw = bytearray(open(path, "rb").read())
    for evil in (struct.pack("<Q", 400), struct.pack("<Q", 4)):  # offs[3] (6): beyond the end / below offs[2] = 5
        broken = bytearray(raw)
        broken[32 + 24:32 + 32] = evil
        open(bad, "wb").write(bytes(broken))
        with pytest.raises(swb.SwbError):
            swb.dbfile_read(bad)


def test_mkdb_tool(swb, tmp_path):
    tool = os.path.join(ROOT, PKG, "bin", "swb_mkdb")
    if not os.path.exists(tool):
        swb.build()
    out = str(tmp_path / "subset.swbdb")
    r = subprocess.run([tool, os.path.join(GOLDEN, "uniprot_subset.fasta"), out], capture_output=True, text=True)
    assert r.returncode == 0 and "111 sequences, 26299 residues" in r.stdout
    codes, offs, fid = swb.dbfile_read(out)
    c0, o0, f0 = swb.read_fasta(os.path.join(GOLDEN, "uniprot_subset.fasta"))
    assert np.array_equal(codes, c0) and np.array_equal(offs, o0) and fid == f0 == 0
    out2 = str(tmp_path / "head5.swbdb")
    r = subprocess.run([tool, "--uniprot-dat", os.path.join(GOLDEN, "uniprot_head5.dat"), out2], capture_output=True,
                       text=True)
    assert r.returncode == 0 and r.stdout.startswith("5 sequences")
    assert subprocess.run([tool], capture_output=True).returncode == 2


def _ref_fasta_records(data):
    """the record rule restated line by line (FASTAParsers.h:73-136 without the '/' padding)"""
    recs, cur, seen = [], bytearray(), False
    for line in data.split(b"\n"):
        if line[:1] == b">":
            if seen:
                recs.append(bytes(cur))
            cur = bytearray()
            seen = True
        else:
            cur += line
    recs.append(bytes(cur))
    return recs, (0 if seen else -1)


@pytest.mark.parametrize("slice_bytes", [1, 7, 64, 1000])
def test_parallel_slices_cut_like_the_sequential_rule(swb, tmp_path, monkeypatch, slice_bytes):
    """the readers cut a file into one slice per host thread at record boundaries; forcing tiny slices on files full
    of edge cases ('>' inside a line, blank lines, CR, text before the first header, no header, no final newline)
    must give the records of the line-by-line rule"""
    monkeypatch.setenv("SWB_INGEST_SLICE_BYTES", str(slice_bytes))
    rng = np.random.default_rng(slice_bytes)
    letters = np.frombuffer(b"ARNDCQEGHILKMFPSTWYVBJZX*UO", dtype=np.uint8)
    cases = [b"junk before\nMORE\n>a\nAC>D\n\n>b\n\n>c\nAC\r\nD\n>\n>d x\nWW", b"NOHEADER\nLINES\n\nONLY\n", b">only\n", b"\n\n>x\nA\n"]
    big = bytearray()
    for i in range(300):
        big += b">sp|%d| desc > with gt\n" % i
        n = int(rng.integers(0, 200))
        s = letters[rng.integers(0, len(letters), n)].tobytes()
        for k in range(0, n, 60):
            big += s[k:k + 60] + (b"\r\n" if i % 17 == 0 else b"\n")
        if i % 23 == 0:
            big += b"\n"
    cases.append(bytes(big))
    for ci, data in enumerate(cases):
        path = tmp_path / ("c%d.fasta" % ci)
        path.write_bytes(data)
        codes, offs, first_id = swb.read_fasta(str(path))
        recs, want_first = _ref_fasta_records(data)
        assert first_id == want_first and len(offs) - 1 == len(recs), ci
        for k, r in enumerate(recs):
            assert np.array_equal(codes[int(offs[k]):int(offs[k + 1])], swb.encode(r.decode("latin-1"))), (ci, k)
    # flat file: "//" inside other lines, the last entry without its "//"
    dat = bytearray()
    want = []
    for i in range(120):
        n = int(rng.integers(1, 300))
        s = letters[rng.integers(0, 24, n)].tobytes()
        dat += b"ID   X%d\nDE   // not a terminator here\nSQ   SEQUENCE %d AA;\n" % (i, n)
        for k in range(0, n, 60):
            row = s[k:k + 60]
            dat += b"     " + b" ".join(row[j:j + 10] for j in range(0, len(row), 10)) + b"\n"
        want.append(s)
        if i != 119:
            dat += b"//\n"
    path = tmp_path / "x.dat"
    path.write_bytes(bytes(dat))
    codes, offs = swb.read_uniprot_dat(str(path))
    assert len(offs) - 1 == len(want)
    for k, s in enumerate(want):
        assert np.array_equal(codes[int(offs[k]):int(offs[k + 1])], swb.encode(s.decode("latin-1"))), k
