"""ctypes binding of the parity oracle (oracle/bsq_oracle.{h,c}).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from blazeseq_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbsq_oracle.so")

# FastxErrorCode (errors.mojo:43-56) + ORA_EMPTY_ERROR
OK, ID_NO_AT, SEP_NO_PLUS, SEQ_QUAL_LEN_MISMATCH, ASCII_INVALID, QUALITY_OUT_OF_RANGE = range(6)
EOF, UNEXPECTED_EOF, BUFFER_EXCEEDED, BUFFER_AT_MAX, OTHER, EMPTY_ERROR = range(6, 12)


class Config(C.Structure):
    _fields_ = [
        ("buffer_capacity", C.c_int64),
        ("buffer_max_capacity", C.c_int64),
        ("buffer_growth_enabled", C.c_int32),
        ("check_ascii", C.c_int32),
        ("check_quality", C.c_int32),
        ("q_lower", C.c_uint8),
        ("q_upper", C.c_uint8),
        ("q_offset", C.c_uint8),
        ("_pad", C.c_uint8),
        ("compat_simd_width", C.c_int32),
        ("reader_max_read", C.c_int64),
    ]


class View(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "header_start", "seq_start", "sep_start", "qual_start", "record_end",
        "id_start", "id_len", "seq_len", "qual_len")]


VIEW_DTYPE = np.dtype([(n, "<i8") for n, _ in View._fields_])


class Error(C.Structure):
    _fields_ = [
        ("code", C.c_int32),
        ("record_number", C.c_int64),
        ("line_number", C.c_int64),
        ("file_position", C.c_int64),
        ("message", C.c_char * 1024),
    ]

    @property
    def text(self) -> str:
        return self.message.decode("latin-1")


def build(force: bool = False) -> str:
    """Compile oracle/libbsq_oracle.so with the committed Makefile."""
    src = os.path.join(_HERE, "bsq_oracle.c")
    hdr = os.path.join(_HERE, "bsq_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_SO) for f in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p = C.c_void_p
        L.ora_default_config.argtypes = [C.POINTER(Config)]
        L.ora_parse_schema.argtypes = [C.c_char_p] + [C.POINTER(C.c_uint8)] * 3
        L.ora_parse_schema.restype = C.c_int
        L.ora_open.argtypes = [u8p, C.c_size_t, C.POINTER(Config)]
        L.ora_open.restype = C.c_void_p
        L.ora_close.argtypes = [C.c_void_p]
        L.ora_has_more.argtypes = [C.c_void_p]
        L.ora_has_more.restype = C.c_int
        for f in (L.ora_next_view, L.ora_next_record):
            f.argtypes = [C.c_void_p, C.POINTER(View), C.POINTER(Error)]
            f.restype = C.c_int
        L.ora_next_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64),
                                     C.POINTER(Error)]
        L.ora_next_batch.restype = C.c_int
        L.ora_build_batch.argtypes = [u8p, C.c_void_p, C.c_int64] + [C.c_void_p] * 5
        L.ora_parse_all.argtypes = [u8p, C.c_size_t, C.POINTER(Config), C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_int64), C.POINTER(Error)]
        L.ora_parse_all.restype = C.c_int64
        L.ora_baseline_mt.argtypes = [u8p, C.c_size_t, C.POINTER(Config), C.c_int, C.c_int64,
                                      C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.ora_baseline_mt.restype = C.c_int64
        L.ora_compute_num_reads_for_size.argtypes = [C.c_int64] * 3
        L.ora_compute_num_reads_for_size.restype = C.c_int64
        L.ora_synth_size.argtypes = [C.c_int64] * 3
        L.ora_synth_size.restype = C.c_int64
        L.ora_synth_generate.argtypes = [C.c_int64] * 7 + [C.c_uint8] * 3 + [C.c_int, C.c_void_p]
        L.ora_synth_generate.restype = C.c_int64
        L.ora_sha256.argtypes = [u8p, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


def schema(name: str):
    lo, up, off = C.c_uint8(), C.c_uint8(), C.c_uint8()
    unknown = lib().ora_parse_schema(name.encode(), C.byref(lo), C.byref(up), C.byref(off))
    return lo.value, up.value, off.value, bool(unknown)


def config(check_ascii=False, check_quality=False, schema_name="generic", buffer_capacity=None,
           buffer_growth_enabled=False, buffer_max_capacity=None, compat_simd_width=0,
           reader_max_read=0) -> Config:
    c = Config()
    lib().ora_default_config(C.byref(c))
    c.check_ascii = int(check_ascii)
    c.check_quality = int(check_quality)
    c.q_lower, c.q_upper, c.q_offset, _ = schema(schema_name)
    if buffer_capacity is not None:
        c.buffer_capacity = buffer_capacity
    if buffer_max_capacity is not None:
        c.buffer_max_capacity = buffer_max_capacity
    c.buffer_growth_enabled = int(buffer_growth_enabled)
    c.compat_simd_width = compat_simd_width
    c.reader_max_read = reader_max_read
    return c


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        assert data.dtype == np.uint8
        return np.ascontiguousarray(data)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data if a.size else 0)


def parse_all(data, cfg: Config | None = None, want_views=True):
    """Canonical whole-stream parse.  Returns (views ndarray, bases, Error)."""
    a = _as_u8(data)
    cfg = cfg or config()
    err = Error()
    bases = C.c_int64()
    cap = (a.size // 4 + 2) if want_views else 0  # a record has >= 4 bytes (4 newlines)
    views = np.zeros(cap, dtype=VIEW_DTYPE)
    n = lib().ora_parse_all(_ptr(a), a.size, C.byref(cfg), _ptr(views) if want_views else None,
                            cap, C.byref(bases), C.byref(err))
    return (views[:n] if want_views else n), bases.value, err


class StreamParser:
    """The streaming model: MemoryReader -> BufferedReader -> FastqParser."""

    def __init__(self, data, cfg: Config | None = None):
        self._a = _as_u8(data)
        self.cfg = cfg or config()
        self._h = lib().ora_open(_ptr(self._a), self._a.size, C.byref(self.cfg))

    def close(self):
        if self._h:
            lib().ora_close(self._h)
            self._h = None

    __del__ = close

    def has_more(self) -> bool:
        return bool(lib().ora_has_more(self._h))

    def _next(self, fn):
        v, e = View(), Error()
        rc = fn(self._h, C.byref(v), C.byref(e))
        return rc, v, e

    def next_view(self):
        return self._next(lib().ora_next_view)

    def next_record(self):
        return self._next(lib().ora_next_record)

    def next_batch(self, max_records=4096):
        views = np.zeros(max_records if max_records else 4096, dtype=VIEW_DTYPE)
        n, e = C.c_int64(), Error()
        rc = lib().ora_next_batch(self._h, max_records, _ptr(views), C.byref(n), C.byref(e))
        return rc, views[: n.value], e

    def fields(self, v):
        """(id, seq, qual) bytes of a View / VIEW_DTYPE row."""
        g = (lambda k: int(getattr(v, k))) if isinstance(v, View) else (lambda k: int(v[k]))
        b = self._a
        return (bytes(b[g("id_start"): g("id_start") + g("id_len")]),
                bytes(b[g("seq_start"): g("seq_start") + g("seq_len")]),
                bytes(b[g("qual_start"): g("qual_start") + g("qual_len")]))


def build_batch(data, views: np.ndarray):
    """FastqBatch SoA arrays (id, seq, qual bytes; id_ends, ends int64)."""
    a = _as_u8(data)
    n = len(views)
    views = np.ascontiguousarray(views)
    idb = np.zeros(int(views["id_len"].sum()), np.uint8)
    sqb = np.zeros(int(views["seq_len"].sum()), np.uint8)
    qlb = np.zeros(int(views["qual_len"].sum()), np.uint8)
    ide = np.zeros(n, np.int64)
    ends = np.zeros(n, np.int64)
    lib().ora_build_batch(_ptr(a), _ptr(views), n, _ptr(idb), _ptr(sqb), _ptr(qlb), _ptr(ide),
                          _ptr(ends))
    return idb, sqb, qlb, ide, ends


def baseline_mt(data, cfg: Config | None = None, mode=0, batch_size=4096, threads=1):
    a = _as_u8(data)
    cfg = cfg or config()
    bases, code = C.c_int64(), C.c_int32()
    n = lib().ora_baseline_mt(_ptr(a), a.size, C.byref(cfg), mode, batch_size, threads,
                              C.byref(bases), C.byref(code))
    return n, bases.value, code.value


def compute_num_reads_for_size(target, mn, mx) -> int:
    return lib().ora_compute_num_reads_for_size(target, mn, mx)


def synth_size(num_reads, mn, mx) -> int:
    return lib().ora_synth_size(num_reads, mn, mx)


def synth(num_reads, mn, mx, min_phred, max_phred, schema_name="generic", first=0, count=None,
          out: np.ndarray | None = None, gc_slots=-1) -> np.ndarray:
    """generate_synthetic_fastq_buffer (utils.mojo:831-917); optionally a slice of records."""
    lo, up, off, _ = schema(schema_name)
    count = num_reads - first if count is None else count
    if first == 0 and count == num_reads:
        size = synth_size(num_reads, mn, mx)
    else:
        digits = len(str(num_reads - 1)) if num_reads > 1 else 1
        size = count * (6 + digits + 1 + 2 * mx + 4)
    if out is None:
        out = np.empty(size, np.uint8)
    assert out.size >= size
    w = lib().ora_synth_generate(num_reads, first, count, mn, mx, min_phred, max_phred, lo, up,
                                 off, gc_slots, _ptr(out))
    return out[:w]


def fasta_parse(data, check_ascii=False):
    """FastaParser over the whole input.  Returns (ids [(bytes)], sequences [(bytes)], Error)."""
    a = _as_u8(data)
    cap = a.size // 2 + 2
    ids = np.zeros(2 * cap, np.int64)
    off = np.zeros(cap + 1, np.int64)
    seq = np.zeros(max(a.size, 1), np.uint8)
    err = Error()
    L = lib()
    L.ora_fasta_parse.restype = C.c_int64
    L.ora_fasta_parse.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    n = L.ora_fasta_parse(_ptr(a), a.size, int(check_ascii), _ptr(ids), _ptr(off), _ptr(seq), cap, C.byref(err))
    id_list = [bytes(a[ids[2 * i]: ids[2 * i] + ids[2 * i + 1]]) for i in range(n)]
    seq_list = [bytes(seq[off[i]: off[i + 1]]) for i in range(n)]
    return id_list, seq_list, err


def sha256(data) -> str:
    a = _as_u8(data)
    out = np.zeros(32, np.uint8)
    lib().ora_sha256(_ptr(a), a.size, _ptr(out))
    return bytes(out).hex()
