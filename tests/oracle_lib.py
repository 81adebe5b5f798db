"""ctypes view of oracle/liboracle.so (the CPU oracle).  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i8p = ctypes.POINTER(ctypes.c_int8)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_i32p = ctypes.POINTER(ctypes.c_int32)


def _ptr(a, t):
    return a.ctypes.data_as(t)


class Oracle:
    def __init__(self):
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
        L = self.lib = ctypes.CDLL(so)
        L.swo_score.restype = ctypes.c_int32
        L.swo_score.argtypes = [_u8p, ctypes.c_uint32, _u8p, ctypes.c_uint32, _i8p, ctypes.c_int32]
        L.swo_scan.restype = None
        L.swo_scan.argtypes = [_u8p, ctypes.c_uint32, _u8p, _u64p, ctypes.c_uint32, _i8p, ctypes.c_int32,
                               _i32p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
        L.swo_scan_affine.restype = None
        L.swo_scan_affine.argtypes = [_u8p, ctypes.c_uint32, _u8p, _u64p, ctypes.c_uint32, _i8p, ctypes.c_int32,
                                      ctypes.c_int32, _i32p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
        L.swo_align.restype = ctypes.c_int32
        L.swo_align.argtypes = [_u8p, ctypes.c_char_p, ctypes.c_uint32, _u8p, ctypes.c_char_p, ctypes.c_uint32,
                                _i8p, ctypes.c_int32, ctypes.c_char_p, ctypes.c_char_p,
                                ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
        L.swo_max_threads.restype = ctypes.c_int
        for f in (L.swo_matrix_blosum50, L.swo_matrix_ident3):
            f.restype = None
            f.argtypes = [_i8p]
        for f in (L.swo_encode_blosum, L.swo_encode_ident):
            f.restype = None
            f.argtypes = [ctypes.c_char_p, ctypes.c_size_t, _u8p]

    def matrix(self, name):
        m = np.zeros((32, 32), dtype=np.int8)
        {"blosum50": self.lib.swo_matrix_blosum50, "ident3": self.lib.swo_matrix_ident3}[name](_ptr(m, _i8p))
        return m

    def encode(self, s, scheme="blosum50"):
        if isinstance(s, str):
            s = s.encode("latin-1")
        out = np.zeros(len(s), dtype=np.uint8)
        f = self.lib.swo_encode_blosum if scheme == "blosum50" else self.lib.swo_encode_ident
        f(s, len(s), _ptr(out, _u8p))
        return out

    def score(self, q, d, m, gap=2):
        q = np.ascontiguousarray(q, dtype=np.uint8)
        d = np.ascontiguousarray(d, dtype=np.uint8)
        m = np.ascontiguousarray(m, dtype=np.int8)
        return int(self.lib.swo_score(_ptr(q, _u8p), len(q), _ptr(d, _u8p), len(d), _ptr(m, _i8p), gap))

    def scan(self, q, codes, offsets, m, gap=2, start=0, stride=1, threads=0, out=None):
        q = np.ascontiguousarray(q, dtype=np.uint8)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        m = np.ascontiguousarray(m, dtype=np.int8)
        n = len(offsets) - 1
        if out is None:
            out = np.full(n, -1, dtype=np.int32)
        self.lib.swo_scan(_ptr(q, _u8p), len(q), _ptr(codes, _u8p), _ptr(offsets, _u64p), n, _ptr(m, _i8p), gap,
                          _ptr(out, _i32p), start, stride, threads)
        return out

    def scan_affine(self, q, codes, offsets, m, gap_open, gap_extend, start=0, stride=1, threads=0):
        q = np.ascontiguousarray(q, dtype=np.uint8)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        m = np.ascontiguousarray(m, dtype=np.int8)
        n = len(offsets) - 1
        out = np.full(n, -1, dtype=np.int32)
        self.lib.swo_scan_affine(_ptr(q, _u8p), len(q), _ptr(codes, _u8p), _ptr(offsets, _u64p), n, _ptr(m, _i8p),
                                 gap_open, gap_extend, _ptr(out, _i32p), start, stride, threads)
        return out

    def align(self, qtxt, dtxt, scheme="ident3", gap=2):
        q = self.encode(qtxt, scheme)
        d = self.encode(dtxt, scheme)
        m = self.matrix(scheme)
        a = ctypes.create_string_buffer(len(qtxt) + len(dtxt) + 1)
        b = ctypes.create_string_buffer(len(qtxt) + len(dtxt) + 1)
        ei, ej = ctypes.c_uint32(), ctypes.c_uint32()
        s = self.lib.swo_align(_ptr(q, _u8p), qtxt.encode(), len(q), _ptr(d, _u8p), dtxt.encode(), len(d),
                               _ptr(m, _i8p), gap, a, b, ctypes.byref(ei), ctypes.byref(ej))
        return int(s), a.value.decode(), b.value.decode(), (ei.value, ej.value)

    def align_affine(self, qtxt, dtxt, go, ge, scheme="blosum50"):
        """Gotoh alignment with traceback (swo_align_affine): (score, aligned a, aligned b, (end_i, end_j))"""
        q = self.encode(qtxt, scheme)
        d = self.encode(dtxt, scheme)
        m = self.matrix(scheme)
        a = ctypes.create_string_buffer(len(qtxt) + len(dtxt) + 1)
        b = ctypes.create_string_buffer(len(qtxt) + len(dtxt) + 1)
        ei, ej = ctypes.c_uint32(), ctypes.c_uint32()
        self.lib.swo_align_affine.restype = ctypes.c_int32
        self.lib.swo_align_affine.argtypes = [_u8p, ctypes.c_char_p, ctypes.c_uint32, _u8p, ctypes.c_char_p, ctypes.c_uint32,
                                              _i8p, ctypes.c_int32, ctypes.c_int32, ctypes.c_char_p, ctypes.c_char_p,
                                              ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
        s = self.lib.swo_align_affine(_ptr(q, _u8p), qtxt.encode(), len(q), _ptr(d, _u8p), dtxt.encode(), len(d),
                                      _ptr(m, _i8p), go, ge, a, b, ctypes.byref(ei), ctypes.byref(ej))
        return int(s), a.value.decode(), b.value.decode(), (ei.value, ej.value)

    def max_threads(self):
        return int(self.lib.swo_max_threads())


def read_fasta(path):
    """Plain multi-FASTA reader for tests: returns (headers, sequences)."""
    heads, seqs = [], []
    for line in open(path):
        line = line.rstrip("\r\n")
        if line.startswith(">"):
            heads.append(line[1:])
            seqs.append("")
        elif seqs:
            seqs[-1] += line
    return heads, seqs


def read_query(path):
    return "".join(open(path).read().split("\n")[1:])


def pack_db(encoded):
    """list of uint8 arrays -> (codes, offsets[n+1] uint64)"""
    offsets = np.zeros(len(encoded) + 1, dtype=np.uint64)
    if encoded:
        offsets[1:] = np.cumsum([len(e) for e in encoded], dtype=np.uint64)
    codes = np.concatenate(encoded) if encoded and offsets[-1] else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(codes, dtype=np.uint8), offsets
