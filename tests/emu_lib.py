"""ctypes view of lib/libswbemu.so (tests/emu/swb_emu.cu): the host emulation of the warp program. Test infrastructure."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "ece1782-smith-waterman-cuda_b200"
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i8p = ctypes.POINTER(ctypes.c_int8)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_ARGS = [_u8p, _u64p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _i8p, ctypes.c_int,
         ctypes.c_int, _u8p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_int,
         ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, _i32p,
         ctypes.POINTER(ctypes.c_uint32)]
_libs = {}


def load(variant=""):
    """variant "": the warp program as the product compiles it; "_blk": the same source with tiny column blocks and pass
    groups of three (-DSWB_BLOCK_CHUNKS=3 -DSWB_PASS_GROUP=3), so that small test inputs cross many block and group borders"""
    if variant not in _libs:
        so = os.path.join(ROOT, PKG, "lib", "libswbemu%s.so" % variant)
        if not os.path.exists(so):
            subprocess.run(["make", "-C", os.path.join(ROOT, PKG), "emu"], check=True, capture_output=True)
        lib = ctypes.CDLL(so)
        lib.swbemu_search.restype = ctypes.c_int
        lib.swbemu_search.argtypes = _ARGS
        _libs[variant] = lib
    return _libs[variant]


def search(codes, offs, m, q, K=32, group_len=384, force_i32=0, chunk_rows=0, thr=-1, gap=2, gap_extend=None,
           shard=0, nshards=1, n_out=None, xl_len=8192, split_k=0, exact_i32=0, direct_len=0, rebase_shift=0, variant=""):
    """One query through the emulated engine flow; returns (scores of the shard, recomputed tiles).
    force_i32: 1 = the exact pass alone over every tile (V16R, or V32 with exact_i32=1)."""
    L = load(variant)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    if len(codes) == 0:
        codes = np.zeros(1, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    q = np.ascontiguousarray(q, dtype=np.uint8)
    m = np.ascontiguousarray(m, dtype=np.int8)
    n = len(offs) - 1
    out = np.full(n if n_out is None else n_out, -7, dtype=np.int32)
    rc = ctypes.c_uint32()
    r = L.swbemu_search(codes.ctypes.data_as(_u8p), offs.ctypes.data_as(_u64p), n, shard, nshards, group_len,
                        m.ctypes.data_as(_i8p), gap, gap if gap_extend is None else gap_extend,
                        q.ctypes.data_as(_u8p) if len(q) else None, len(q), K, force_i32, chunk_rows, thr, xl_len,
                        split_k, exact_i32, direct_len, rebase_shift, out.ctypes.data_as(_i32p), ctypes.byref(rc))
    assert r == 0, r
    return out, rc.value
