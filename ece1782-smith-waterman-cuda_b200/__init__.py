"""Python view of libswb.so (C ABI in include/swb.h) plus a host-side mirror of the reference interface.

The package directory name contains hyphens, so import it with
``importlib.import_module("ece1782-smith-waterman-cuda_b200")`` (tests, bench.py and
__graft_entry__.py do). Everything here is plumbing: ctypes marshalling of plain pointers. All
scoring happens in the CUDA library; there is no CPU fallback -- creating an ``Engine`` without a
GPU, or without the built library, raises.

Mirror of the reference surface (reference file:line):
  FASTAQuery / FASTADatabase  -> src/FASTAParsers.h:33-138 (same parsing rules, padding to 8 with '/')
  smith_waterman_cuda         -> src/SWSolver.h:9, src/SWSolver.cu:266-404 (appends (id, score) pairs in
                                 descending padded length, file order inside a length bucket)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SWB_LIB: another build of the library (A/B measurements of kernel versions; tools/)
LIB_PATH = os.environ.get("SWB_LIB") or os.path.join(_HERE, "lib", "libswb.so")

SWB_SCORING_BLOSUM50_REF = 0
SWB_SCORING_IDENT3 = 1

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i8p = ctypes.POINTER(ctypes.c_int8)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_i32p = ctypes.POINTER(ctypes.c_int32)


class SwbStats(ctypes.Structure):
    _fields_ = [
        ("device_ms", ctypes.c_double),
        ("load_ms", ctypes.c_double),
        ("cells", ctypes.c_uint64),
        ("padded_cells", ctypes.c_uint64),
        ("db_residues", ctypes.c_uint64),
        ("db_residues_total", ctypes.c_uint64),
        ("db_sequences", ctypes.c_uint32),
        ("tiles", ctypes.c_uint32),
        ("tiles_by_group", ctypes.c_uint32 * 6),
        ("recomputed_tiles", ctypes.c_uint32),
        ("kernel_launches", ctypes.c_uint32),
        ("last_k", ctypes.c_uint32),
        ("sm_count", ctypes.c_uint32),
        ("pack_us", ctypes.c_uint32),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


class SwbPlanInfo(ctypes.Structure):
    _fields_ = [
        ("n_total", ctypes.c_uint32), ("n_local", ctypes.c_uint32), ("tiles", ctypes.c_uint32),
        ("max_len", ctypes.c_uint32),
        ("residues_local", ctypes.c_uint64), ("residues_total", ctypes.c_uint64), ("res_bytes", ctypes.c_uint64),
        ("bnd_elems", ctypes.c_uint64), ("padded_cols", ctypes.c_uint64),
        ("tiles_by_group", ctypes.c_uint32 * 6),
    ]


# every symbol include/swb.h declares (tests check the library exports exactly these)
ABI_SYMBOLS = [
    "swb_create", "swb_destroy", "swb_last_error", "swb_set_option", "swb_set_stream", "swb_set_scoring",
    "swb_set_scoring_preset", "swb_set_scoring_affine", "swb_scoring_matrix", "swb_encode", "swb_db_load", "swb_db_count", "swb_db_ids",
    "swb_search", "swb_search_batch", "swb_search_batch_scatter", "swb_search_batch_topk", "swb_fetch_scores",
    "swb_topk", "swb_stats", "swb_plan_describe",
    "swb_group_create", "swb_group_create_env", "swb_group_destroy", "swb_group_last_error", "swb_group_size", "swb_group_engine",
    "swb_group_db_parts", "swb_group_set_option", "swb_group_set_scoring", "swb_group_set_scoring_preset",
    "swb_group_set_scoring_affine", "swb_group_db_load", "swb_group_search_batch", "swb_group_search_batch_topk",
    "swb_group_stats", "swb_layout_parts", "swb_layout_parts_batch", "swb_layout_query_groups",
    "swb_microbench", "swb_pack_time", "swb_align", "swb_align_batch", "swb_read_fasta", "swb_read_uniprot_dat", "swb_free", "swb_dbfile_write",
    "swb_dbfile_open", "swb_dbfile_count", "swb_dbfile_first_id", "swb_dbfile_offsets", "swb_dbfile_codes",
    "swb_dbfile_close",
]

_lib = None


def build(verbose=False):
    """Compiles lib/libswb.so (and the drop-in C++ surface) for sm_100a with nvcc."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libswb.so failed")


def lib():
    """Loads libswb.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("CUDA library %s is missing: run __graft_entry__.build() (there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    L.swb_create.restype = ctypes.c_int
    L.swb_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
    L.swb_destroy.restype = None
    L.swb_destroy.argtypes = [vp]
    L.swb_last_error.restype = ctypes.c_char_p
    L.swb_last_error.argtypes = [vp]
    L.swb_set_option.restype = ctypes.c_int
    L.swb_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int64]
    L.swb_set_stream.restype = ctypes.c_int
    L.swb_set_stream.argtypes = [vp, vp]
    L.swb_set_scoring.restype = ctypes.c_int
    L.swb_set_scoring.argtypes = [vp, _i8p, ctypes.c_int, ctypes.c_int]
    L.swb_set_scoring_affine.restype = ctypes.c_int
    L.swb_set_scoring_affine.argtypes = [vp, _i8p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.swb_set_scoring_preset.restype = ctypes.c_int
    L.swb_set_scoring_preset.argtypes = [vp, ctypes.c_int]
    L.swb_scoring_matrix.restype = ctypes.c_int
    L.swb_scoring_matrix.argtypes = [ctypes.c_int, _i8p, ctypes.POINTER(ctypes.c_int)]
    L.swb_encode.restype = ctypes.c_int
    L.swb_encode.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, _u8p]
    L.swb_db_load.restype = ctypes.c_int
    L.swb_db_load.argtypes = [vp, _u8p, _u64p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
    L.swb_db_count.restype = ctypes.c_uint32
    L.swb_db_count.argtypes = [vp]
    L.swb_db_ids.restype = ctypes.c_int
    L.swb_db_ids.argtypes = [vp, _u32p]
    L.swb_search.restype = ctypes.c_int
    L.swb_search.argtypes = [vp, _u8p, ctypes.c_uint32, _i32p]
    L.swb_search_batch.restype = ctypes.c_int
    L.swb_search_batch.argtypes = [vp, _u8p, _u64p, ctypes.c_uint32, _i32p]
    L.swb_search_batch_scatter.restype = ctypes.c_int
    L.swb_search_batch_scatter.argtypes = [vp, _u8p, _u64p, ctypes.c_uint32, _i32p, ctypes.c_uint64]
    L.swb_search_batch_topk.restype = ctypes.c_int
    L.swb_search_batch_topk.argtypes = [vp, _u8p, _u64p, ctypes.c_uint32, ctypes.c_uint32, _u32p, _i32p]
    L.swb_group_create.restype = ctypes.c_int
    L.swb_group_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    L.swb_group_destroy.restype = None
    L.swb_group_destroy.argtypes = [vp]
    L.swb_group_last_error.restype = ctypes.c_char_p
    L.swb_group_last_error.argtypes = [vp]
    L.swb_group_size.restype = ctypes.c_int
    L.swb_group_size.argtypes = [vp]
    L.swb_group_engine.restype = vp
    L.swb_group_engine.argtypes = [vp, ctypes.c_int]
    L.swb_group_db_parts.restype = ctypes.c_int
    L.swb_group_db_parts.argtypes = [vp]
    L.swb_group_set_option.restype = ctypes.c_int
    L.swb_group_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int64]
    L.swb_group_set_scoring.restype = ctypes.c_int
    L.swb_group_set_scoring.argtypes = [vp, _i8p, ctypes.c_int, ctypes.c_int]
    L.swb_group_set_scoring_preset.restype = ctypes.c_int
    L.swb_group_set_scoring_preset.argtypes = [vp, ctypes.c_int]
    L.swb_group_set_scoring_affine.restype = ctypes.c_int
    L.swb_group_set_scoring_affine.argtypes = [vp, _i8p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.swb_group_db_load.restype = ctypes.c_int
    L.swb_group_db_load.argtypes = [vp, _u8p, _u64p, ctypes.c_uint32]
    L.swb_group_search_batch.restype = ctypes.c_int
    L.swb_group_search_batch.argtypes = [vp, _u8p, _u64p, ctypes.c_uint32, _i32p]
    L.swb_group_search_batch_topk.restype = ctypes.c_int
    L.swb_group_search_batch_topk.argtypes = [vp, _u8p, _u64p, ctypes.c_uint32, ctypes.c_uint32, _u32p, _i32p]
    L.swb_group_stats.restype = ctypes.c_int
    L.swb_group_stats.argtypes = [vp, ctypes.POINTER(SwbStats)]
    L.swb_layout_parts.restype = ctypes.c_int
    L.swb_layout_parts.argtypes = [ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32]
    L.swb_layout_parts_batch.restype = ctypes.c_int
    L.swb_layout_parts_batch.argtypes = [ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, _u64p, ctypes.c_uint32]
    L.swb_layout_query_groups.restype = ctypes.c_int
    L.swb_layout_query_groups.argtypes = [_u64p, ctypes.c_uint32, ctypes.c_int, _u32p]
    L.swb_fetch_scores.restype = ctypes.c_int
    L.swb_fetch_scores.argtypes = [vp, ctypes.c_uint32, _i32p]
    L.swb_topk.restype = ctypes.c_int
    L.swb_topk.argtypes = [vp, _i32p, ctypes.c_uint32, _u32p, _i32p]
    L.swb_stats.restype = ctypes.c_int
    L.swb_stats.argtypes = [vp, ctypes.POINTER(SwbStats)]
    L.swb_plan_describe.restype = ctypes.c_int
    L.swb_plan_describe.argtypes = [_u64p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                    ctypes.POINTER(SwbPlanInfo), _u32p, _u32p]
    L.swb_align.restype = ctypes.c_int
    L.swb_align.argtypes = [vp, _u8p, ctypes.c_uint32, ctypes.c_uint32, _i32p, _u32p, _u32p, _u8p, ctypes.c_uint32, _u32p]
    L.swb_align_batch.restype = ctypes.c_int
    L.swb_align_batch.argtypes = [vp, _u8p, _u64p, ctypes.c_uint32, _u32p, _u32p, ctypes.c_uint32, _i32p, _u32p, _u32p,
                                  _u8p, _u64p, _u32p]
    L.swb_read_fasta.restype = ctypes.c_int
    L.swb_read_fasta.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(_u8p), ctypes.POINTER(_u64p), _u32p,
                                 _i32p]
    L.swb_read_uniprot_dat.restype = ctypes.c_int
    L.swb_read_uniprot_dat.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(_u8p), ctypes.POINTER(_u64p), _u32p]
    L.swb_free.restype = None
    L.swb_free.argtypes = [vp]
    L.swb_dbfile_write.restype = ctypes.c_int
    L.swb_dbfile_write.argtypes = [ctypes.c_char_p, _u8p, _u64p, ctypes.c_uint32, ctypes.c_int32]
    L.swb_dbfile_open.restype = ctypes.c_int
    L.swb_dbfile_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
    L.swb_dbfile_count.restype = ctypes.c_uint32
    L.swb_dbfile_count.argtypes = [vp]
    L.swb_dbfile_first_id.restype = ctypes.c_int32
    L.swb_dbfile_first_id.argtypes = [vp]
    L.swb_dbfile_offsets.restype = _u64p
    L.swb_dbfile_offsets.argtypes = [vp]
    L.swb_dbfile_codes.restype = _u8p
    L.swb_dbfile_codes.argtypes = [vp]
    L.swb_dbfile_close.restype = None
    L.swb_dbfile_close.argtypes = [vp]
    L.swb_pack_time.restype = ctypes.c_int
    L.swb_pack_time.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]
    L.swb_microbench.restype = ctypes.c_int
    L.swb_microbench.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                 ctypes.POINTER(ctypes.c_double)]
    _lib = L
    return L


class SwbError(RuntimeError):
    pass


# ---- host helpers (no GPU needed) -------------------------------------------------------------
def scoring_matrix(preset=SWB_SCORING_BLOSUM50_REF):
    m = np.zeros((32, 32), dtype=np.int8)
    gap = ctypes.c_int()
    rc = lib().swb_scoring_matrix(preset, m.ctypes.data_as(_i8p), ctypes.byref(gap))
    if rc != 0:
        raise SwbError("unknown preset %r" % (preset,))
    return m, gap.value


def encode(text, preset=SWB_SCORING_BLOSUM50_REF):
    if isinstance(text, str):
        text = text.encode("latin-1")
    out = np.zeros(len(text), dtype=np.uint8)
    rc = lib().swb_encode(preset, text, len(text), out.ctypes.data_as(_u8p))
    if rc != 0:
        raise SwbError("swb_encode failed")
    return out


def pack_sequences(encoded):
    """list of uint8 code arrays -> (codes, offsets[n+1] uint64) as swb_db_load / swb_search_batch take them"""
    offsets = np.zeros(len(encoded) + 1, dtype=np.uint64)
    if len(encoded):
        offsets[1:] = np.cumsum([len(e) for e in encoded], dtype=np.uint64)
    if len(encoded) and offsets[-1] > 0:
        codes = np.ascontiguousarray(np.concatenate(encoded), dtype=np.uint8)
    else:
        codes = np.zeros(1, dtype=np.uint8)
    return codes, offsets


def _take(codes_p, offs_p, n):
    """copies malloc'ed (codes, offsets) out of the library and releases them"""
    L = lib()
    offsets = np.ctypeslib.as_array(offs_p, shape=(n + 1,)).copy()
    total = int(offsets[-1])
    codes = np.ctypeslib.as_array(codes_p, shape=(max(total, 1),))[:total].copy()
    L.swb_free(ctypes.cast(codes_p, ctypes.c_void_p))
    L.swb_free(ctypes.cast(offs_p, ctypes.c_void_p))
    return codes, offsets


def read_fasta(path, preset=SWB_SCORING_BLOSUM50_REF):
    """(codes, offsets, first_id): the records of a FASTA file as the reference parser cuts them, without '/' padding"""
    cp, op, n, fid = _u8p(), _u64p(), ctypes.c_uint32(), ctypes.c_int32()
    rc = lib().swb_read_fasta(os.fsencode(path), preset, ctypes.byref(cp), ctypes.byref(op), ctypes.byref(n),
                              ctypes.byref(fid))
    if rc != 0:
        raise SwbError("swb_read_fasta(%s) failed (%d)" % (path, rc))
    codes, offsets = _take(cp, op, n.value)
    return codes, offsets, fid.value


def read_uniprot_dat(path, preset=SWB_SCORING_BLOSUM50_REF):
    cp, op, n = _u8p(), _u64p(), ctypes.c_uint32()
    rc = lib().swb_read_uniprot_dat(os.fsencode(path), preset, ctypes.byref(cp), ctypes.byref(op), ctypes.byref(n))
    if rc != 0:
        raise SwbError("swb_read_uniprot_dat(%s) failed (%d)" % (path, rc))
    return _take(cp, op, n.value)


def dbfile_write(path, codes, offsets, first_id=0):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    cpz = codes if len(codes) else np.zeros(1, np.uint8)
    rc = lib().swb_dbfile_write(os.fsencode(path), cpz.ctypes.data_as(_u8p), offsets.ctypes.data_as(_u64p),
                                len(offsets) - 1, first_id)
    if rc != 0:
        raise SwbError("swb_dbfile_write(%s) failed (%d)" % (path, rc))


def dbfile_read(path):
    """(codes, offsets, first_id) copied out of a memory-mapped encoded database"""
    L = lib()
    h = ctypes.c_void_p()
    rc = L.swb_dbfile_open(os.fsencode(path), ctypes.byref(h))
    if rc != 0:
        raise SwbError("swb_dbfile_open(%s) failed (%d)" % (path, rc))
    try:
        n = L.swb_dbfile_count(h)
        offsets = np.ctypeslib.as_array(L.swb_dbfile_offsets(h), shape=(n + 1,)).copy()
        total = int(offsets[-1])
        codes = np.ctypeslib.as_array(L.swb_dbfile_codes(h), shape=(max(total, 1),))[:total].copy() if total else \
            np.zeros(0, np.uint8)
        return codes, offsets, int(L.swb_dbfile_first_id(h))
    finally:
        L.swb_dbfile_close(h)


MIN_PART_SEQUENCES = 450000  # SWB_MIN_PART_SEQUENCES


def layout_parts(n, ndev, min_part_sequences=MIN_PART_SEQUENCES, qoffsets=None):
    """database parts P of the P x R device grid (include/swb.h, engine group); with the offsets of a batch, P also
    makes sure the R = ndev / P query groups can be balanced (swb_layout_parts_batch)"""
    if qoffsets is None:
        return int(lib().swb_layout_parts(int(n), int(ndev), int(min_part_sequences)))
    qoffsets = np.ascontiguousarray(qoffsets, dtype=np.uint64)
    return int(lib().swb_layout_parts_batch(int(n), int(ndev), int(min_part_sequences), qoffsets.ctypes.data_as(_u64p),
                                            len(qoffsets) - 1))


def layout_query_groups(qoffsets, groups):
    """group (0..groups-1) of every query of a batch: equal total length, longest-processing-time first"""
    qoffsets = np.ascontiguousarray(qoffsets, dtype=np.uint64)
    nq = len(qoffsets) - 1
    out = np.zeros(max(nq, 1), dtype=np.uint32)
    rc = lib().swb_layout_query_groups(qoffsets.ctypes.data_as(_u64p), nq, int(groups), out.ctypes.data_as(_u32p))
    if rc != 0:
        raise SwbError("swb_layout_query_groups failed")
    return out[:nq]


def plan_describe(offsets, shard=0, nshards=1, group_len=0, want_ids=False):
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    info = SwbPlanInfo()
    rc = lib().swb_plan_describe(offsets.ctypes.data_as(_u64p), n, shard, nshards, group_len, ctypes.byref(info),
                                 None, None)
    if rc != 0:
        raise SwbError("swb_plan_describe failed")
    if not want_ids:
        return info
    sorted_ids = np.zeros(info.n_local, dtype=np.uint32)
    shard_ids = np.zeros(info.n_local, dtype=np.uint32)
    lib().swb_plan_describe(offsets.ctypes.data_as(_u64p), n, shard, nshards, group_len, ctypes.byref(info),
                            sorted_ids.ctypes.data_as(_u32p), shard_ids.ctypes.data_as(_u32p))
    return info, sorted_ids, shard_ids


MICROBENCH_KINDS = ["viaddmax_s16x2_relu", "vimax3_s16x2", "vadd2", "prmt", "v16_mix", "viaddmax+imad", "imad",
                    "scalar_addmax", "v16b_mix", "hmnmx2+mask", "viaddmax+hmnmx2", "viaddmax+vadd2", "viaddmax+prmt",
                    "viaddmax+vimax3", "v16_mix_vimax3"]


def microbench(device=0, kind=4):
    """giga lane-instructions/s of one instruction kind over the whole GPU (measurement support)"""
    g = ctypes.c_double()
    ms = ctypes.c_double()
    rc = lib().swb_microbench(device, kind, ctypes.byref(g), ctypes.byref(ms))
    if rc != 0:
        raise SwbError("swb_microbench failed (%d)" % rc)
    return g.value


# ---- engine -----------------------------------------------------------------------------------
class Engine:
    """One GPU, one resident (shard of a) database."""

    def __init__(self, device=0, **options):
        self._h = ctypes.c_void_p()
        self._L = lib()
        rc = self._L.swb_create(ctypes.byref(self._h), device)
        if rc != 0:
            raise SwbError("swb_create failed: %s" % self._L.swb_last_error(None).decode())
        for k, v in options.items():
            self.set_option(k, v)

    def _check(self, rc, what):
        if rc != 0:
            raise SwbError("%s failed (%d): %s" % (what, rc, self._L.swb_last_error(self._h).decode()))

    def close(self):
        if self._h:
            self._L.swb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        self._check(self._L.swb_set_option(self._h, key.encode(), int(value)), "swb_set_option(%s)" % key)

    def set_stream(self, cuda_stream_ptr):
        self._check(self._L.swb_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)), "swb_set_stream")

    def set_scoring(self, matrix, gap):
        m = np.ascontiguousarray(matrix, dtype=np.int8)
        assert m.ndim == 2 and m.shape[0] == m.shape[1]
        self._check(self._L.swb_set_scoring(self._h, m.ctypes.data_as(_i8p), m.shape[0], gap), "swb_set_scoring")

    def set_scoring_affine(self, matrix, gap_open, gap_extend):
        """Gotoh gaps: a gap of length L costs gap_open + (L-1)*gap_extend (gap_open == gap_extend: linear)."""
        m = np.ascontiguousarray(matrix, dtype=np.int8)
        assert m.ndim == 2 and m.shape[0] == m.shape[1]
        self._check(self._L.swb_set_scoring_affine(self._h, m.ctypes.data_as(_i8p), m.shape[0], gap_open, gap_extend),
                    "swb_set_scoring_affine")

    def set_scoring_preset(self, preset):
        self._check(self._L.swb_set_scoring_preset(self._h, preset), "swb_set_scoring_preset")

    def db_load(self, codes, offsets, shard=0, nshards=1):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._check(self._L.swb_db_load(self._h, codes.ctypes.data_as(_u8p), offsets.ctypes.data_as(_u64p),
                                        len(offsets) - 1, shard, nshards), "swb_db_load")

    def db_count(self):
        return int(self._L.swb_db_count(self._h))

    def db_ids(self):
        ids = np.zeros(self.db_count(), dtype=np.uint32)
        if len(ids):
            self._check(self._L.swb_db_ids(self._h, ids.ctypes.data_as(_u32p)), "swb_db_ids")
        return ids

    def search(self, query_codes, out=None):
        q = np.ascontiguousarray(query_codes, dtype=np.uint8)
        if out is None:
            out = np.zeros(self.db_count(), dtype=np.int32)
        qp = q.ctypes.data_as(_u8p) if len(q) else None
        self._check(self._L.swb_search(self._h, qp, len(q), out.ctypes.data_as(_i32p)), "swb_search")
        return out

    def search_batch(self, queries, fetch=True, out=None):
        """queries: list of code arrays. Returns an (nq, db_count) int32 array, or None when fetch=False
        (results stay on the device; see fetch_scores)."""
        qcodes, qoffs = pack_sequences([np.ascontiguousarray(q, dtype=np.uint8) for q in queries])
        return self.search_batch_packed(qcodes, qoffs, fetch=fetch, out=out)

    def search_batch_packed(self, qcodes, qoffs, fetch=True, out=None):
        nq = len(qoffs) - 1
        if fetch and out is None:
            out = np.zeros((nq, self.db_count()), dtype=np.int32)
        outp = out.ctypes.data_as(_i32p) if fetch else None
        self._check(self._L.swb_search_batch(self._h, qcodes.ctypes.data_as(_u8p), qoffs.ctypes.data_as(_u64p), nq,
                                             outp), "swb_search_batch")
        return out if fetch else None

    def search_batch_scatter(self, qcodes, qoffs, out):
        """scores by database id into `out` (nq x n_total, shared by the engines of all shards)"""
        nq = len(qoffs) - 1
        assert out.dtype == np.int32 and out.flags.c_contiguous and out.shape[0] == nq
        self._check(self._L.swb_search_batch_scatter(self._h, qcodes.ctypes.data_as(_u8p), qoffs.ctypes.data_as(_u64p),
                                                     nq, out.ctypes.data_as(_i32p), out.shape[1]),
                    "swb_search_batch_scatter")
        return out

    def search_batch_topk(self, qcodes, qoffs, k):
        """device-side hit lists: (ids, scores), each nq x k, score descending / database id ascending"""
        nq = len(qoffs) - 1
        ids = np.zeros((nq, k), dtype=np.uint32)
        top = np.zeros((nq, k), dtype=np.int32)
        self._check(self._L.swb_search_batch_topk(self._h, qcodes.ctypes.data_as(_u8p), qoffs.ctypes.data_as(_u64p),
                                                  nq, k, ids.ctypes.data_as(_u32p), top.ctypes.data_as(_i32p)),
                    "swb_search_batch_topk")
        return ids, top

    def fetch_scores(self, query_index):
        out = np.zeros(self.db_count(), dtype=np.int32)
        self._check(self._L.swb_fetch_scores(self._h, query_index, out.ctypes.data_as(_i32p)), "swb_fetch_scores")
        return out

    def topk(self, scores, k):
        scores = np.ascontiguousarray(scores, dtype=np.int32)
        ids = np.zeros(k, dtype=np.uint32)
        top = np.zeros(k, dtype=np.int32)
        self._check(self._L.swb_topk(self._h, scores.ctypes.data_as(_i32p), k, ids.ctypes.data_as(_u32p),
                                     top.ctypes.data_as(_i32p)), "swb_topk")
        return ids, top

    def align(self, query_codes, db_id, subject_len):
        """traceback alignment against one database sequence: (score, end_i, end_j, ops) -- see swb_align"""
        q = np.ascontiguousarray(query_codes, dtype=np.uint8)
        cap = len(q) + int(subject_len) + 1
        ops = np.zeros(cap, dtype=np.uint8)
        score = ctypes.c_int32()
        ei, ej, n = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        qp = q.ctypes.data_as(_u8p) if len(q) else None
        self._check(self._L.swb_align(self._h, qp, len(q), int(db_id), ctypes.byref(score), ctypes.byref(ei),
                                      ctypes.byref(ej), ops.ctypes.data_as(_u8p), cap, ctypes.byref(n)), "swb_align")
        return int(score.value), int(ei.value), int(ej.value), ops[:n.value].copy()

    def align_batch(self, queries, hits, subject_lens):
        """traceback alignments of a list of hits in one launch (swb_align_batch). queries: list of code arrays; hits:
        list of (query index, db id); subject_lens: the subject length of every hit (sizes the ops buffers). Returns a
        list of (score, end_i, end_j, ops)."""
        qcodes, qoffs = pack_sequences([np.ascontiguousarray(q, dtype=np.uint8) for q in queries])
        nh = len(hits)
        hq = np.ascontiguousarray([h[0] for h in hits], dtype=np.uint32)
        hd = np.ascontiguousarray([h[1] for h in hits], dtype=np.uint32)
        qlen = [len(q) for q in queries]
        room = np.asarray([(qlen[int(h[0])] if 0 <= int(h[0]) < len(qlen) else 0) + int(l) + 1
                           for h, l in zip(hits, subject_lens)], dtype=np.uint64)  # a bad index is the C side's error to report
        ooff = np.zeros(nh + 1, dtype=np.uint64)
        ooff[1:] = np.cumsum(room)
        ops = np.zeros(max(int(ooff[-1]), 1), dtype=np.uint8)
        scores = np.zeros(max(nh, 1), dtype=np.int32)
        ei = np.zeros(max(nh, 1), dtype=np.uint32)
        ej = np.zeros(max(nh, 1), dtype=np.uint32)
        cnt = np.zeros(max(nh, 1), dtype=np.uint32)
        self._check(self._L.swb_align_batch(self._h, qcodes.ctypes.data_as(_u8p), qoffs.ctypes.data_as(_u64p), len(queries),
                                            hq.ctypes.data_as(_u32p), hd.ctypes.data_as(_u32p), nh,
                                            scores.ctypes.data_as(_i32p), ei.ctypes.data_as(_u32p),
                                            ej.ctypes.data_as(_u32p), ops.ctypes.data_as(_u8p),
                                            ooff.ctypes.data_as(_u64p), cnt.ctypes.data_as(_u32p)), "swb_align_batch")
        return [(int(scores[h]), int(ei[h]), int(ej[h]), ops[int(ooff[h]):int(ooff[h]) + int(cnt[h])].copy())
                for h in range(nh)]

    def pack_time(self, reps=10):
        """(microseconds per launch, bytes read + written per launch) of the pack kernel of the loaded database"""
        us, nbytes = ctypes.c_double(), ctypes.c_uint64()
        self._check(self._L.swb_pack_time(self._h, reps, ctypes.byref(us), ctypes.byref(nbytes)), "swb_pack_time")
        return us.value, int(nbytes.value)

    def stats(self):
        s = SwbStats()
        self._check(self._L.swb_stats(self._h, ctypes.byref(s)), "swb_stats")
        return s.as_dict()


class EngineGroup:
    """Every GPU of the box in one process (swb_group_*): P database parts x R query groups."""

    def __init__(self, devices=None, **options):
        self._h = ctypes.c_void_p()
        self._L = lib()
        if devices is None:
            rc = self._L.swb_group_create(ctypes.byref(self._h), None, 0)
        elif isinstance(devices, int):
            rc = self._L.swb_group_create(ctypes.byref(self._h), None, devices)
        else:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self._L.swb_group_create(ctypes.byref(self._h), arr, len(devices))
        if rc != 0:
            raise SwbError("swb_group_create failed: %s" % self._L.swb_group_last_error(None).decode())
        self.n_total = 0
        for k, v in options.items():
            self.set_option(k, v)

    def _check(self, rc, what):
        if rc != 0:
            raise SwbError("%s failed (%d): %s" % (what, rc, self._L.swb_group_last_error(self._h).decode()))

    def close(self):
        if self._h:
            self._L.swb_group_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return int(self._L.swb_group_size(self._h))

    def db_parts(self):
        return int(self._L.swb_group_db_parts(self._h))

    def set_option(self, key, value):
        self._check(self._L.swb_group_set_option(self._h, key.encode(), int(value)), "swb_group_set_option(%s)" % key)

    def set_scoring_preset(self, preset):
        self._check(self._L.swb_group_set_scoring_preset(self._h, preset), "swb_group_set_scoring_preset")

    def set_scoring_affine(self, matrix, gap_open, gap_extend):
        m = np.ascontiguousarray(matrix, dtype=np.int8)
        self._check(self._L.swb_group_set_scoring_affine(self._h, m.ctypes.data_as(_i8p), m.shape[0], gap_open,
                                                         gap_extend), "swb_group_set_scoring_affine")

    def db_load(self, codes, offsets):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._check(self._L.swb_group_db_load(self._h, codes.ctypes.data_as(_u8p), offsets.ctypes.data_as(_u64p),
                                              len(offsets) - 1), "swb_group_db_load")
        self.n_total = len(offsets) - 1

    def search_batch_packed(self, qcodes, qoffs, out=None):
        nq = len(qoffs) - 1
        if out is None:
            out = np.zeros((nq, self.n_total), dtype=np.int32)
        self._check(self._L.swb_group_search_batch(self._h, qcodes.ctypes.data_as(_u8p), qoffs.ctypes.data_as(_u64p),
                                                   nq, out.ctypes.data_as(_i32p)), "swb_group_search_batch")
        return out

    def search_batch(self, queries, out=None):
        qcodes, qoffs = pack_sequences([np.ascontiguousarray(q, dtype=np.uint8) for q in queries])
        return self.search_batch_packed(qcodes, qoffs, out=out)

    def search_batch_topk(self, qcodes, qoffs, k):
        nq = len(qoffs) - 1
        ids = np.zeros((nq, k), dtype=np.uint32)
        top = np.zeros((nq, k), dtype=np.int32)
        self._check(self._L.swb_group_search_batch_topk(self._h, qcodes.ctypes.data_as(_u8p),
                                                        qoffs.ctypes.data_as(_u64p), nq, k, ids.ctypes.data_as(_u32p),
                                                        top.ctypes.data_as(_i32p)), "swb_group_search_batch_topk")
        return ids, top

    def stats(self):
        s = SwbStats()
        self._check(self._L.swb_group_stats(self._h, ctypes.byref(s)), "swb_group_stats")
        return s.as_dict()


# ---- host-side mirror of the reference interface ----------------------------------------------
TILE_SIZE = 8  # FASTAParsers.h:12


def round_up(n, multiple):
    """FASTAParsers.h:21-31"""
    if multiple == 0:
        return n
    r = n % multiple
    return n if r == 0 else n + multiple - r


def _getlines(path):
    """std::getline semantics on a file that may not exist (FASTAParsers.h:40-47, 75-88): a missing
    file yields no lines; the final line needs no terminator; '\\r' stays in the line."""
    try:
        data = open(path, "rb").read()
    except OSError:
        return []
    if not data:
        return []
    lines = data.split(b"\n")
    if lines[-1] == b"":
        lines.pop()
    return [l.decode("latin-1") for l in lines]


class FASTAQuery:
    """FASTAParsers.h:33-63: the first line is dropped, the rest concatenated."""

    def __init__(self, filepath, is_query=True):
        self.is_query = is_query
        self.buffer = "".join(_getlines(filepath)[1:])

    def get_buffer(self):
        return self.buffer

    def print_buffer(self):
        print(self.buffer)


class FASTADatabase:
    """FASTAParsers.h:65-138: records split at lines starting with '>', each sequence padded with '/' to a
    multiple of 8 and filed under its padded length; ids are 0-based record ordinals (-1 for text before
    the first '>')."""

    def __init__(self, filepath):
        self.parsedDB = {}  # padded length -> list of (id, padded sequence), file order
        self.largestSubjectLength = 0
        self.numSubjects = 0
        self.subjectLengthSum = 0
        cur, _id, first = [], -1, True
        for line in _getlines(filepath):
            if line[:1] == ">":
                if not first:
                    self._add(_id, "".join(cur))
                first = False
                cur = []
                _id += 1
            else:
                cur.append(line)
        self._add(_id, "".join(cur))

    def _add(self, _id, seq):
        seq = seq + "/" * (round_up(len(seq), TILE_SIZE) - len(seq))
        self.parsedDB.setdefault(len(seq), []).append((_id, seq))
        self.subjectLengthSum += len(seq)
        self.largestSubjectLength = max(self.largestSubjectLength, len(seq))
        self.numSubjects += 1

    def ordered(self):
        """(id, padded sequence) in the order SWSolver.cu:383-390 reports results."""
        out = []
        for length in sorted(self.parsedDB, reverse=True):
            out.extend(self.parsedDB[length])
        return out


_solver_engines = {}


def smith_waterman_cuda(query, db, result, device=0):
    """Same contract as the reference entry point (SWSolver.h:9): appends one (id, score) pair per database
    sequence to `result`, BLOSUM50 ('*' zeroed) with linear gap 2, in the reference's result order. The packed
    database is cached on the GPU per FASTADatabase object."""
    eng = _solver_engines.get(device)
    if eng is None:
        eng = _solver_engines[device] = {"engine": Engine(device), "db": None}
    e = eng["engine"]
    ordered = db.ordered()
    if eng["db"] is not db:
        e.set_scoring_preset(SWB_SCORING_BLOSUM50_REF)
        codes, offsets = pack_sequences([encode(s) for _, s in ordered])
        e.db_load(codes, offsets)
        eng["db"] = db
    scores = e.search(encode(query.get_buffer()))
    for k, (sid, _) in enumerate(ordered):
        result.append((sid, int(scores[k])))
    return result


def render_alignment(query_text, subject_text, end_i, end_j, ops):
    """The two aligned strings cpu.cpp prints (cpu.cpp:80-108) from swb_align's result."""
    i = end_i - sum(1 for o in ops if o != 1)
    j = end_j - sum(1 for o in ops if o != 2)
    a, b = [], []
    for o in ops:
        if o == 1:
            a.append("-")
            b.append(subject_text[j])
            j += 1
        elif o == 2:
            a.append(query_text[i])
            b.append("-")
            i += 1
        else:
            a.append(query_text[i])
            b.append(subject_text[j])
            i += 1
            j += 1
    return "".join(a), "".join(b)


# ---- multi-GPU host side: shards are independent, only small results are exchanged -------------
def merge_shard_scores(n_total, parts):
    """parts: iterable of (ids, scores) per shard (ids = database ids of that shard). Returns the full
    score vector in database order; every id must be covered exactly once."""
    out = np.full(n_total, np.iinfo(np.int32).min, dtype=np.int32)
    seen = np.zeros(n_total, dtype=np.uint8)
    for ids, scores in parts:
        ids = np.asarray(ids, dtype=np.int64)
        out[ids] = np.asarray(scores, dtype=np.int32)
        seen[ids] += 1
    if not (seen == 1).all():
        raise SwbError("shards do not partition the database")
    return out


def merge_topk(parts, k):
    """parts: iterable of (ids, scores) hit lists (one per GPU). k best overall: score descending, id ascending."""
    hits = [(int(s), int(i)) for ids, scores in parts for i, s in zip(ids, scores) if int(i) != 0xFFFFFFFF]
    hits.sort(key=lambda h: (-h[0], h[1]))
    hits = hits[:k]
    return np.array([h[1] for h in hits], dtype=np.uint32), np.array([h[0] for h in hits], dtype=np.int32)


def sharded_search(engine, queries, rank, world, k=10, gather=None):
    """One rank's part of a sharded scan: scores of this rank's shard for every query plus its top-k lists; with
    `gather` (e.g. torch.distributed.all_gather_object wrapped to return the list) the lists are merged and the
    merged (ids, scores) per query are returned as well."""
    scores = engine.search_batch(queries)
    mine = [engine.topk(scores[i], k) for i in range(len(queries))]
    if gather is None:
        return scores, mine, None
    everyone = gather([(ids.tolist(), top.tolist()) for ids, top in mine])
    merged = [merge_topk([everyone[r][qi] for r in range(world)], k) for qi in range(len(queries))]
    return scores, mine, merged
