"""ctypes binding of include/blazeseq_gpu.h (the C ABI in blazeseq_b200/lib/libblazeseq_gpu.so).

This module never parses anything itself and has no fallback: if the shared library is missing
or no CUDA device is usable, creating a parser raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# BSQ_LIB: development override to A/B a differently tuned build of the same library (scripts/build_variants.sh)
LIB_PATH = os.environ.get("BSQ_LIB") or os.path.join(HERE, "lib", "libblazeseq_gpu.so")

# FastxErrorCode (blazeseq/errors.mojo:43-56) + library failures
OK, ID_NO_AT, SEP_NO_PLUS, SEQ_QUAL_LEN_MISMATCH, ASCII_INVALID, QUALITY_OUT_OF_RANGE = range(6)
EOF, UNEXPECTED_EOF, BUFFER_EXCEEDED, BUFFER_AT_MAX, OTHER, EMPTY_ERROR = range(6, 12)
E_CUDA, E_ARG, E_NO_DEVICE, E_NOMEM, E_STATE, E_IO = -1, -2, -3, -4, -5, -6
WANT_OFFSETS, WANT_BATCHES, WANT_WHOLE_BATCHES = 1, 2, 4

# every symbol include/blazeseq_gpu.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "bsq_default_config", "bsq_parse_schema", "bsq_abi_version", "bsq_create", "bsq_destroy",
    "bsq_last_error_text", "bsq_set_batch_size", "bsq_parse_device", "bsq_parse_host", "bsq_get_offsets", "bsq_get_batch",
    "bsq_get_soa", "bsq_batch_to_host", "bsq_offsets_to_host", "bsq_pass_device_input",
    "bsq_last_timing", "bsq_compute_num_reads_for_size", "bsq_synth_size", "bsq_synth_device",
    "bsq_summarize_device", "bsq_shard_prefix", "bsq_stream_open", "bsq_stream_next", "bsq_stream_region",
    "bsq_stream_get_stats", "bsq_stream_close", "bsq_quality_sums", "bsq_soa_to_host", "bsq_stream_region_info",
    "bsq_fasta_parse_device", "bsq_fasta_parse_host", "bsq_fasta_get", "bsq_fasta_to_host",
    "bsq_gzip_open", "bsq_gzip_read", "bsq_gzip_error", "bsq_gzip_close", "bsq_write_records",
]


class Config(C.Structure):
    _fields_ = [
        ("device_id", C.c_int32), ("check_ascii", C.c_int32), ("check_quality", C.c_int32),
        ("q_lower", C.c_uint8), ("q_upper", C.c_uint8), ("q_offset", C.c_uint8), ("_pad0", C.c_uint8),
        ("buffer_capacity", C.c_int64), ("buffer_max_capacity", C.c_int64),
        ("buffer_growth_enabled", C.c_int32), ("batch_size", C.c_int32),
        ("h2d_chunk_bytes", C.c_int64), ("force_id_slow_path", C.c_int32), ("inflate_threads", C.c_int32),
        ("compat_q5_width", C.c_int32), ("host_inflate", C.c_int32),
    ]


class Error(C.Structure):
    _fields_ = [
        ("code", C.c_int32), ("_pad", C.c_int32), ("record_number", C.c_int64),
        ("line_number", C.c_int64), ("file_position", C.c_int64), ("message", C.c_char * 1024),
    ]

    @property
    def text(self) -> str:
        return self.message.decode("latin-1")


class PassResult(C.Structure):
    _fields_ = [
        ("n_records", C.c_int64), ("n_bases", C.c_int64), ("bytes_consumed", C.c_int64),
        ("n_newlines", C.c_int64), ("n_batches", C.c_int64), ("n_windows", C.c_int32),
        ("id_slow_path", C.c_int32), ("stop", Error),
    ]


class OffsetsView(C.Structure):
    _fields_ = [
        ("stream_base", C.c_int64), ("first_record", C.c_int64), ("n_records", C.c_int64),
        ("line_ends", C.c_void_p), ("id_spans", C.c_void_p), ("window_bytes", C.c_void_p),
    ]


class BatchView(C.Structure):
    _fields_ = [
        ("num_records", C.c_int64), ("seq_len", C.c_int64), ("total_id_bytes", C.c_int64),
        ("quality_offset", C.c_uint8), ("_pad", C.c_uint8 * 7), ("sequence_bytes", C.c_int64),
        ("sequence_buffer", C.c_void_p), ("qual_buffer", C.c_void_p), ("id_buffer", C.c_void_p),
        ("ends", C.c_void_p), ("id_ends", C.c_void_p),
    ]


class Summary(C.Structure):
    _fields_ = [("w", C.c_uint32 * 16)]

    @property
    def count(self) -> int:
        return int(self.w[0])


class StreamStats(C.Structure):
    _fields_ = [("bytes_read", C.c_uint64), ("regions", C.c_uint64), ("reader_busy_s", C.c_double),
                ("parse_s", C.c_double), ("wait_reader_s", C.c_double), ("h2d_s", C.c_double), ("inflate_s", C.c_double),
                ("compressed_bytes", C.c_uint64), ("launch_s", C.c_double), ("wait_inflate_s", C.c_double)]


SOURCE_PLAIN, SOURCE_GZIP, SOURCE_AUTO = 0, 1, 2


class FastaResult(C.Structure):
    _fields_ = [("n_records", C.c_int64), ("n_bases", C.c_int64), ("n_lines", C.c_int64), ("stop", Error)]


class FastaView(C.Structure):
    _fields_ = [("n_records", C.c_int64), ("sequence_bytes", C.c_int64), ("sequence", C.c_void_p), ("seq_starts", C.c_void_p),
                ("id_start", C.c_void_p), ("id_len", C.c_void_p), ("input", C.c_void_p)]


class ShardStart(C.Structure):
    _fields_ = [("newline_rank", C.c_int64), ("first_record", C.c_int64), ("skip_bytes", C.c_int64),
                ("phase", C.c_int32), ("_pad", C.c_int32)]


_lib = None


def lib():
    """Loads the shared library (building it first if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    L = C.CDLL(LIB_PATH)
    vp, u8, i64, u64, i32, u32 = C.c_void_p, C.c_uint8, C.c_int64, C.c_uint64, C.c_int32, C.c_uint32
    L.bsq_default_config.argtypes = [C.POINTER(Config)]
    L.bsq_default_config.restype = None
    L.bsq_parse_schema.argtypes = [C.c_char_p] + [C.POINTER(u8)] * 3
    L.bsq_parse_schema.restype = i32
    L.bsq_abi_version.restype = u32
    L.bsq_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.bsq_destroy.argtypes = [vp]
    L.bsq_destroy.restype = None
    L.bsq_last_error_text.argtypes = [vp]
    L.bsq_last_error_text.restype = C.c_char_p
    L.bsq_set_batch_size.argtypes = [vp, i32]
    L.bsq_set_batch_size.restype = i32
    for f in (L.bsq_parse_device, L.bsq_parse_host):
        f.argtypes = [vp, vp, u64, i64, i64, i32, u32, C.POINTER(PassResult)]
        f.restype = i32
    L.bsq_get_offsets.argtypes = [vp, i32, C.POINTER(OffsetsView)]
    L.bsq_get_batch.argtypes = [vp, i64, C.POINTER(BatchView)]
    L.bsq_get_soa.argtypes = [vp, C.POINTER(BatchView)]
    L.bsq_batch_to_host.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.bsq_offsets_to_host.argtypes = [vp, i32, vp, vp]
    L.bsq_pass_device_input.argtypes = [vp]
    L.bsq_pass_device_input.restype = vp
    L.bsq_last_timing.argtypes = [vp, C.POINTER(C.c_float * 5), C.POINTER(i64)]
    L.bsq_compute_num_reads_for_size.argtypes = [i64] * 3
    L.bsq_compute_num_reads_for_size.restype = i64
    L.bsq_synth_size.argtypes = [i64] * 3
    L.bsq_synth_size.restype = i64
    L.bsq_synth_device.argtypes = [vp, vp, u64] + [i64] * 7 + [u8] * 3 + [C.POINTER(u64)]
    L.bsq_summarize_device.argtypes = [vp, vp, u64, C.POINTER(Summary)]
    L.bsq_shard_prefix.argtypes = [C.POINTER(Summary), C.POINTER(u64), i32, C.POINTER(ShardStart)]
    L.bsq_stream_open.argtypes = [vp, C.c_char_p, i32, u64, C.POINTER(vp)]
    L.bsq_stream_next.argtypes = [vp, u32, C.POINTER(PassResult)]
    L.bsq_stream_region.argtypes = [vp, C.POINTER(u64), C.POINTER(i64), C.POINTER(i64)]
    L.bsq_stream_region.restype = vp
    L.bsq_stream_region_info.argtypes = [vp, C.POINTER(u64), C.POINTER(i64), C.POINTER(i64)]
    L.bsq_stream_region_info.restype = None
    L.bsq_stream_get_stats.argtypes = [vp, C.POINTER(StreamStats)]
    L.bsq_stream_close.argtypes = [vp]
    L.bsq_stream_close.restype = None
    L.bsq_quality_sums.argtypes = [vp, i64, i64, vp, vp]
    L.bsq_soa_to_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.bsq_fasta_parse_device.argtypes = [vp, vp, u64, C.POINTER(FastaResult)]
    L.bsq_fasta_parse_host.argtypes = [vp, vp, u64, C.POINTER(FastaResult)]
    L.bsq_fasta_get.argtypes = [vp, C.POINTER(FastaView)]
    L.bsq_fasta_to_host.argtypes = [vp, vp, vp, vp, vp]
    for name in ("bsq_fasta_parse_device", "bsq_fasta_parse_host", "bsq_fasta_get", "bsq_fasta_to_host"):
        getattr(L, name).restype = i32
    L.bsq_soa_to_host.restype = i32
    L.bsq_quality_sums.restype = i32
    L.bsq_write_records.argtypes = [vp, i64, i64, vp, u64, vp, vp, C.POINTER(u64)]
    L.bsq_write_records.restype = i32
    L.bsq_gzip_open.argtypes = [C.c_char_p, i32, u64, C.POINTER(vp)]
    L.bsq_gzip_open.restype = i32
    L.bsq_gzip_read.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.bsq_gzip_read.restype = i32
    L.bsq_gzip_error.argtypes = [vp]
    L.bsq_gzip_error.restype = C.c_char_p
    L.bsq_gzip_close.argtypes = [vp]
    L.bsq_gzip_close.restype = None
    for name in ("bsq_stream_open", "bsq_stream_next", "bsq_stream_get_stats", "bsq_create", "bsq_get_offsets", "bsq_get_batch", "bsq_get_soa", "bsq_batch_to_host",
                 "bsq_offsets_to_host", "bsq_last_timing", "bsq_synth_device", "bsq_summarize_device",
                 "bsq_shard_prefix"):
        getattr(L, name).restype = i32
    _lib = L
    return L


class BsqLibraryError(RuntimeError):
    """A library failure (status < 0): CUDA error, bad argument, no device."""


def check(status: int, handle=None, what: str = "") -> int:
    if status < 0:
        detail = ""
        if handle:
            detail = lib().bsq_last_error_text(handle).decode("latin-1")
        names = {E_CUDA: "CUDA failure", E_ARG: "bad argument", E_NO_DEVICE: "no usable CUDA device",
                 E_NOMEM: "out of memory", E_STATE: "call sequence error", E_IO: "read / inflate failure"}
        raise BsqLibraryError(f"{what or 'blazeseq_gpu'}: {names.get(status, status)} {detail}".strip())
    return status


def parse_schema(name: str):
    lo, up, off = C.c_uint8(), C.c_uint8(), C.c_uint8()
    unknown = lib().bsq_parse_schema(name.encode(), C.byref(lo), C.byref(up), C.byref(off))
    return lo.value, up.value, off.value, bool(unknown)


def default_config() -> Config:
    c = Config()
    lib().bsq_default_config(C.byref(c))
    return c
