"""Host-side mirror of BlazeSeq's FASTQ parser API over the B200 C ABI.

Same names, argument meaning and error behaviour as the reference (paths relative to the
reference repository root):

    ParserConfig          blazeseq/fastq/parser.mojo:33-74
    FastqParser           blazeseq/fastq/parser.mojo:77-625  (has_more / next_view / next_record /
                          next_batch / views() / records() / batches())
    FastqView             blazeseq/fastq/record.mojo:431-550
    FastqRecord           blazeseq/fastq/record.mojo:230-428 (the subset the path produces)
    FastqBatch            blazeseq/fastq/record_batch.mojo:19-207
    DeviceFastqBatch      blazeseq/fastq/record_batch.mojo:210-244
    MemoryReader / FileReader / GZFile / RapidgzipReader   blazeseq/io/readers.mojo:86-443

The reference is Mojo and no Mojo toolchain exists in this image, so the host side is Python;
all byte work (scan, boundary resolution, validation, SoA packing) happens in the CUDA kernels
behind include/blazeseq_gpu.h.  Nothing here parses FASTQ on the CPU and there is no fallback:
without the shared library or a CUDA device, constructing a FastqParser raises.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
from dataclasses import dataclass
from typing import Iterator, Optional

import numpy as np

from . import _capi as capi

EOF = "EOF"                      # CONSTS.mojo:19
DEFAULT_CAPACITY = 256 * 1024    # CONSTS.mojo:26
MAX_CAPACITY = 2 ** 30           # CONSTS.mojo:27-28
DEFAULT_BATCH_SIZE = 4096        # CONSTS.mojo:31


class BlazeSeqError(Exception):
    """Mojo `Error(String)`; str(e) is the reference's message text."""

    def __init__(self, message: str, code: int = capi.OTHER, record_number: int = 0,
                 line_number: int = 0, file_position: int = 0):
        super().__init__(message)
        self.code, self.record_number = code, record_number
        self.line_number, self.file_position = line_number, file_position


class EOFError(BlazeSeqError):  # noqa: A001 - mirrors blazeseq/io/buffered.mojo:102-112
    def __init__(self):
        super().__init__(EOF, capi.EOF)


@dataclass
class ParserConfig:
    """parser.mojo:33-74."""
    buffer_capacity: int = DEFAULT_CAPACITY
    buffer_max_capacity: int = MAX_CAPACITY
    buffer_growth_enabled: bool = False
    check_ascii: bool = False
    check_quality: bool = False
    quality_schema: Optional[str] = None


@dataclass(frozen=True)
class QualitySchema:
    """quality_schema.mojo:9-31."""
    SCHEMA: str
    LOWER: int
    UPPER: int
    OFFSET: int


def parse_schema(name: str) -> QualitySchema:
    """utils.mojo:612-637 (unknown names print a warning and fall back to generic)."""
    lo, up, off, unknown = capi.parse_schema(name)
    if unknown:
        print("Unknown quality schema please choose one of 'sanger', 'solexa', 'illumina_1.3', "
              "'illumina_1.5' 'illumina_1.8', or 'generic'.\nParsing with generic schema.")
        name = "generic"
    return QualitySchema(name, lo, up, off)


# ------------------------------------------------------------------------------------------------
# byte sources (blazeseq/io/readers.mojo)
# ------------------------------------------------------------------------------------------------


class Reader:
    """trait Reader (readers.mojo:51-79): read_to_buffer returns bytes read, 0 = EOF."""

    def read_to_buffer(self, buf, amt: int, pos: int = 0) -> int:
        raise NotImplementedError

    @staticmethod
    def _check(buf, amt, pos):
        # readers.mojo:124-135,184-196
        if pos > len(buf):
            raise BlazeSeqError("Position is outside the buffer")
        if amt > len(buf) - pos:
            raise BlazeSeqError("Number of elements to read is bigger than the available space in the buffer")
        if amt < 0:
            raise BlazeSeqError("The amount to be read should be positive")


class MemoryReader(Reader):
    """readers.mojo:140-223."""

    def __init__(self, data):
        if isinstance(data, str):
            data = data.encode("latin-1")
        self.data = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        self.position = 0

    def read_to_buffer(self, buf, amt, pos=0):
        self._check(buf, amt, pos)
        k = min(amt, self.data.size - self.position)
        if k <= 0:
            return 0
        buf[pos:pos + k] = self.data[self.position:self.position + k]
        self.position += k
        return k

    def reset(self):
        self.position = 0


class FileReader(Reader):
    """readers.mojo:86-137."""

    def __init__(self, path):
        self.path = os.fspath(path)
        self._f = open(self.path, "rb", buffering=0)

    def read_to_buffer(self, buf, amt, pos=0):
        self._check(buf, amt, pos)
        got = self._f.readinto(memoryview(buf)[pos:pos + amt])
        return got or 0


class GZFile(Reader):
    """readers.mojo:283-377 (zlib gzread)."""

    def __init__(self, path, mode: str = "rb"):
        self.path = os.fspath(path)
        self._f = gzip.open(self.path, "rb")

    def read_to_buffer(self, buf, amt, pos=0):
        self._check(buf, amt, pos)
        chunk = self._f.read(amt)
        k = len(chunk)
        if k:
            buf[pos:pos + k] = np.frombuffer(chunk, dtype=np.uint8)
        return k


class RapidgzipReader(Reader):
    """readers.mojo:380-443: parallel gzip decompression, `parallelism` worker threads (0 = all cores).

    rapidgzip is not in this image; the native library carries its own two-stage speculative decoder
    (csrc/bsq_pgzip.h, `bsq_gzip_*`): chunks of the compressed file are inflated in parallel from the first
    deflate block found in each, references into the unknown window are resolved when the chunks are stitched,
    every member's CRC-32 is verified.  Through FastqParser's native pipeline a BGZF file is inflated on the
    device instead (k_inflate_members)."""

    def __init__(self, path, parallelism: int = 0, *, chunk_bytes: int = 0):
        self.path = os.fspath(path)
        self.parallelism = int(parallelism)
        self.chunk_bytes = int(chunk_bytes)       # compressed bytes per speculative chunk (0 = 2 MiB)
        self._h = C.c_void_p()
        self._closed = False
        with open(self.path, "rb"):      # "Raises: Error: If the file cannot be opened" (readers.mojo:405-407)
            pass

    def _open(self):
        # the decoder threads start on the first read: a FastqParser that takes the file through the library's own
        # stream pipeline never needs this handle
        st = capi.lib().bsq_gzip_open(self.path.encode(), self.parallelism, self.chunk_bytes, C.byref(self._h))
        if st != 0:
            raise OSError(f"RapidgzipReader: cannot open {self.path} as gzip")

    def read_to_buffer(self, buf, amt, pos=0):
        self._check(buf, amt, pos)
        if amt == 0:
            return 0
        if self._closed:
            raise ValueError("RapidgzipReader is closed")
        if not self._h:
            self._open()
        view = buf[pos:pos + amt]
        if not (isinstance(view, np.ndarray) and view.dtype == np.uint8 and view.flags["C_CONTIGUOUS"]):
            tmp = np.empty(amt, dtype=np.uint8)
            k = self._read(tmp)
            buf[pos:pos + k] = tmp[:k]
            return k
        return self._read(view)

    def _read(self, arr: np.ndarray) -> int:
        got = C.c_uint64(0)
        st = capi.lib().bsq_gzip_read(self._h, arr.ctypes.data_as(C.c_void_p), arr.size, C.byref(got))
        if st != 0:
            raise BlazeSeqError("Error reading from gzip file: " + capi.lib().bsq_gzip_error(self._h).decode("latin-1"))
        return int(got.value)

    def close(self):
        self._closed = True
        if getattr(self, "_h", None):
            capi.lib().bsq_gzip_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# records
# ------------------------------------------------------------------------------------------------


class FastqView:
    """record.mojo:431-550: zero-copy spans into the parser's region (valid until the parser moves on)."""

    __slots__ = ("_id", "_sequence", "_quality", "_phred_offset")

    def __init__(self, id, sequence, quality, phred_offset: int = 33):
        self._id, self._sequence, self._quality, self._phred_offset = id, sequence, quality, phred_offset

    def id(self) -> bytes:
        return bytes(self._id)

    def sequence(self) -> bytes:
        return bytes(self._sequence)

    def quality(self) -> bytes:
        return bytes(self._quality)

    def __len__(self) -> int:
        return len(self._sequence)

    def byte_len(self) -> int:
        return 1 + len(self._id) + len(self._sequence) + len(self._quality) + 5

    def phred_scores(self, offset: Optional[int] = None) -> list:
        off = self._phred_offset if offset is None else offset
        return [(b - off) & 0xFF for b in bytes(self._quality)]

    def write(self) -> bytes:
        return b"@" + self.id() + b"\n" + self.sequence() + b"\n+\n" + self.quality() + b"\n"


class FastqRecord:
    """record.mojo:230-428 (owned).  `id`, `sequence`, `quality`, `phred_scores` follow the
    Python binding (python/blazeseq_parser.mojo:348-420)."""

    __slots__ = ("_id", "_sequence", "_quality", "_phred_offset")

    def __init__(self, id, sequence, quality, phred_offset: int = 33):
        def b(x):
            return x.encode("latin-1") if isinstance(x, str) else bytes(x)
        self._id, self._sequence, self._quality, self._phred_offset = b(id), b(sequence), b(quality), phred_offset

    @property
    def id(self) -> str:
        return self._id.decode("latin-1")

    @property
    def sequence(self) -> str:
        return self._sequence.decode("latin-1")

    @property
    def quality(self) -> str:
        return self._quality.decode("latin-1")

    @property
    def phred_scores(self) -> list:
        return [(c - self._phred_offset) & 0xFF for c in self._quality]

    def __len__(self) -> int:
        return len(self._sequence)

    def __eq__(self, o) -> bool:
        return (isinstance(o, FastqRecord) and self._id == o._id and self._sequence == o._sequence
                and self._quality == o._quality)

    def __repr__(self):
        return f"FastqRecord(id={self.id!r}, len={len(self)})"

    def write(self) -> bytes:
        return b"@" + self._id + b"\n" + self._sequence + b"\n+\n" + self._quality + b"\n"


class DeviceFastqBatch:
    """record_batch.mojo:210-220: device-resident SoA (pointers into the parser's arena)."""

    def __init__(self, gpu: "GpuParser", index: int, view: capi.BatchView):
        self._gpu, self._index = gpu, index
        # the pointers below live in the parser's arena and are overwritten by its next pass: every accessor
        # checks that the parser is still on the pass this batch was cut from
        self._generation = gpu.generation
        self.num_records = int(view.num_records)
        self.seq_len = int(view.seq_len)
        self.quality_offset = int(view.quality_offset)
        self.total_id_bytes = int(view.total_id_bytes)
        self.qual_buffer = view.qual_buffer
        self.sequence_buffer = view.sequence_buffer
        self.id_buffer = view.id_buffer
        self.ends = view.ends
        self.id_ends = view.id_ends

    def valid(self) -> bool:
        return self._gpu.generation == self._generation

    def _check(self):
        if not self.valid():
            raise BlazeSeqError("DeviceFastqBatch is stale: the parser has run another pass since this batch was cut "
                                "(device batches are views into the parser's arena, valid until its next pass)")

    def pointers(self):
        """(sequence, quality, id, ends, id_ends) device addresses; raises if the parser has moved on."""
        self._check()
        return self.sequence_buffer, self.qual_buffer, self.id_buffer, self.ends, self.id_ends

    def copy_to_host(self) -> "FastqBatch":
        self._check()
        return FastqBatch._from_arrays(*self._gpu.batch_to_host(self._index), quality_offset=self.quality_offset)

    def write(self) -> bytes:
        """The batch's records as four-line FASTQ text (FastqRecord.write, record.mojo:390-402), serialised on the device."""
        self._check()
        m = self._gpu.cfg.batch_size
        text, _ = self._gpu.write_records(self._index * m, self.num_records)
        return text.tobytes()


class FastqBatch:
    """record_batch.mojo:19-207: five arrays, Int64 inclusive cumulative ends restarting per batch."""

    def __init__(self, batch_size: int = DEFAULT_BATCH_SIZE, avg_record_size: int = 150, quality_offset: int = 33):
        self._id_bytes = np.zeros(0, np.uint8)
        self._quality_bytes = np.zeros(0, np.uint8)
        self._sequence_bytes = np.zeros(0, np.uint8)
        self._id_ends = np.zeros(0, np.int64)
        self._ends = np.zeros(0, np.int64)
        self._quality_offset = quality_offset
        self._device: Optional[DeviceFastqBatch] = None

    @classmethod
    def _from_arrays(cls, seq, qual, idb, ends, id_ends, quality_offset=33, device=None):
        b = cls(quality_offset=quality_offset)
        b._sequence_bytes, b._quality_bytes, b._id_bytes, b._ends, b._id_ends = seq, qual, idb, ends, id_ends
        b._device = device
        return b

    def add(self, record) -> None:
        """record_batch.mojo:65-87."""
        q, s, i = (np.frombuffer(bytes(x), np.uint8) for x in (record._quality, record._sequence, record._id))
        self._quality_bytes = np.concatenate([self._quality_bytes, q])
        self._sequence_bytes = np.concatenate([self._sequence_bytes, s])
        self._id_bytes = np.concatenate([self._id_bytes, i])
        pe = int(self._ends[-1]) if self._ends.size else 0
        pi = int(self._id_ends[-1]) if self._id_ends.size else 0
        self._ends = np.append(self._ends, np.int64(pe + q.size))
        self._id_ends = np.append(self._id_ends, np.int64(pi + i.size))
        self._device = None

    def num_records(self) -> int:
        return int(self._ends.size)

    def seq_len(self) -> int:
        return int(self._ends[-1])

    def quality_offset(self) -> int:
        return self._quality_offset

    def __len__(self) -> int:
        return self.num_records()

    def __repr__(self) -> str:
        return f"FastqBatch(records={self.num_records()}, quality_offset={self._quality_offset})"

    def _range(self, ends, idx):
        return (0 if idx == 0 else int(ends[idx - 1])), int(ends[idx])

    def get_record(self, index: int) -> FastqRecord:
        if index < 0 or index >= self.num_records():
            raise BlazeSeqError("FastqBatch.get_record index out of range")
        a, b = self._range(self._id_ends, index)
        c, d = self._range(self._ends, index)
        return FastqRecord(self._id_bytes[a:b].tobytes(), self._sequence_bytes[c:d].tobytes(),
                           self._quality_bytes[c:d].tobytes(), self._quality_offset)

    def get_ref(self, index: int) -> FastqView:
        if index < 0 or index >= self.num_records():
            raise BlazeSeqError("FastqBatch.get_ref index out of range")
        a, b = self._range(self._id_ends, index)
        c, d = self._range(self._ends, index)
        return FastqView(self._id_bytes[a:b], self._sequence_bytes[c:d], self._quality_bytes[c:d],
                         self._quality_offset)

    def to_records(self) -> list:
        return [self.get_record(i) for i in range(self.num_records())]

    def __iter__(self) -> Iterator[FastqRecord]:
        return iter(self.to_records())

    def to_device(self) -> Optional[DeviceFastqBatch]:
        """record_batch.mojo:89-90.  Batches cut by the parser are already on the device."""
        return self._device


# ------------------------------------------------------------------------------------------------
# thin object over the C ABI
# ------------------------------------------------------------------------------------------------


class GpuParser:
    """One bsq_parser: one device, its streams and arenas."""

    def __init__(self, check_ascii=False, check_quality=False, schema: QualitySchema | None = None,
                 batch_size=DEFAULT_BATCH_SIZE, device_id=0, buffer_capacity=DEFAULT_CAPACITY,
                 buffer_max_capacity=MAX_CAPACITY, buffer_growth_enabled=False, h2d_chunk_bytes=None,
                 force_id_slow_path=False, inflate_threads=0, compat_q5_width=0, host_inflate=False):
        L = capi.lib()
        cfg = capi.default_config()
        cfg.device_id = device_id
        cfg.check_ascii, cfg.check_quality = int(check_ascii), int(check_quality)
        if schema is not None:
            cfg.q_lower, cfg.q_upper, cfg.q_offset = schema.LOWER, schema.UPPER, schema.OFFSET
        cfg.batch_size = batch_size
        cfg.buffer_capacity, cfg.buffer_max_capacity = buffer_capacity, buffer_max_capacity
        cfg.buffer_growth_enabled = int(buffer_growth_enabled)
        if h2d_chunk_bytes:
            cfg.h2d_chunk_bytes = h2d_chunk_bytes
        cfg.force_id_slow_path = int(force_id_slow_path)
        cfg.inflate_threads = int(inflate_threads)   # BGZF members of a stream are inflated by this many host threads (0 = all)
        # reproduce the reference's quality check as written for a SIMD width of W bytes (record.mojo:90-102)
        cfg.compat_q5_width = int(compat_q5_width)
        cfg.host_inflate = int(host_inflate)       # BGZF members inflated by host threads instead of k_inflate_members
        self.cfg = cfg
        self.generation = 0   # bumped by every pass: DeviceFastqBatch views of earlier passes are stale
        self._h = C.c_void_p()
        capi.check(L.bsq_create(C.byref(cfg), C.byref(self._h)), None, "bsq_create")
        self.result: Optional[capi.PassResult] = None

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().bsq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_batch_size(self, m: int):
        if m != self.cfg.batch_size:
            capi.check(capi.lib().bsq_set_batch_size(self._h, m), self._h)
            self.cfg.batch_size = m
            self.generation += 1

    def parse_host(self, data: np.ndarray, stream_offset=0, first_record=0, is_last=True,
                   want=capi.WANT_OFFSETS) -> capi.PassResult:
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        r = capi.PassResult()
        ptr = C.c_void_p(data.ctypes.data if data.size else 0)
        self.generation += 1
        capi.check(capi.lib().bsq_parse_host(self._h, ptr, data.size, stream_offset, first_record,
                                             int(is_last), want, C.byref(r)), self._h, "bsq_parse_host")
        self.result = r
        return r

    def parse_device(self, dev_ptr: int, n: int, stream_offset=0, first_record=0, is_last=True,
                     want=capi.WANT_BATCHES) -> capi.PassResult:
        r = capi.PassResult()
        self.generation += 1
        capi.check(capi.lib().bsq_parse_device(self._h, C.c_void_p(dev_ptr), n, stream_offset, first_record,
                                               int(is_last), want, C.byref(r)), self._h, "bsq_parse_device")
        self.result = r
        return r

    def offsets_to_host(self, window: int):
        v = capi.OffsetsView()
        capi.check(capi.lib().bsq_get_offsets(self._h, window, C.byref(v)), self._h, "bsq_get_offsets")
        n = int(v.n_records)
        le = np.zeros(4 * n + 1, np.uint32)
        sp = np.zeros(2 * n, np.uint32)
        capi.check(capi.lib().bsq_offsets_to_host(self._h, window, C.c_void_p(le.ctypes.data),
                                                  C.c_void_p(sp.ctypes.data if n else 0)), self._h)
        return v, le, sp

    def batch_view(self, index: int) -> capi.BatchView:
        v = capi.BatchView()
        capi.check(capi.lib().bsq_get_batch(self._h, index, C.byref(v)), self._h, "bsq_get_batch")
        return v

    def soa_view(self) -> capi.BatchView:
        v = capi.BatchView()
        capi.check(capi.lib().bsq_get_soa(self._h, C.byref(v)), self._h, "bsq_get_soa")
        return v

    def batch_to_host(self, index: int):
        v = self.batch_view(index)
        n = int(v.num_records)
        seq = np.zeros(int(v.sequence_bytes), np.uint8)
        qual = np.zeros(int(v.seq_len), np.uint8)
        idb = np.zeros(int(v.total_id_bytes), np.uint8)
        ends = np.zeros(n, np.int64)
        id_ends = np.zeros(n, np.int64)

        def p(a):
            return C.c_void_p(a.ctypes.data if a.size else 0)
        capi.check(capi.lib().bsq_batch_to_host(self._h, index, p(seq), p(qual), p(idb), p(ends), p(id_ends)),
                   self._h, "bsq_batch_to_host")
        return seq, qual, idb, ends, id_ends

    def soa_to_host(self, seq, qual, idb, ends, id_ends):
        """bsq_soa_to_host into caller arrays (numpy or torch CPU tensors, pinned for asynchronous copies)."""
        def p(a):
            if a is None:
                return C.c_void_p(0)
            return C.c_void_p(a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data)
        capi.check(capi.lib().bsq_soa_to_host(self._h, p(seq), p(qual), p(idb), p(ends), p(id_ends)), self._h,
                   "bsq_soa_to_host")

    def fasta_parse_host(self, data: np.ndarray) -> capi.FastaResult:
        """bsq_fasta_parse_host: FastaParser.next_record in a loop over `data` (fasta/parser.mojo:60-200)."""
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        r = capi.FastaResult()
        self.generation += 1
        capi.check(capi.lib().bsq_fasta_parse_host(self._h, C.c_void_p(data.ctypes.data if data.size else 0), data.size,
                                                   C.byref(r)), self._h, "bsq_fasta_parse_host")
        return r

    def fasta_parse_device(self, dev_ptr: int, n: int) -> capi.FastaResult:
        r = capi.FastaResult()
        self.generation += 1
        capi.check(capi.lib().bsq_fasta_parse_device(self._h, C.c_void_p(dev_ptr), n, C.byref(r)), self._h,
                   "bsq_fasta_parse_device")
        return r

    def fasta_view(self) -> capi.FastaView:
        v = capi.FastaView()
        capi.check(capi.lib().bsq_fasta_get(self._h, C.byref(v)), self._h, "bsq_fasta_get")
        return v

    def fasta_to_host(self):
        """(sequence bytes, seq_starts[n+1], id bytes, id_starts[n+1]) of the last FASTA pass."""
        v = self.fasta_view()
        n = int(v.n_records)
        seq = np.zeros(max(int(v.sequence_bytes), 1), np.uint8)
        ss = np.zeros(n + 1, np.uint64)
        # ids are at most as long as the input: size them from the device table
        ids = np.zeros(1, np.uint8)
        ist = np.zeros(n + 1, np.uint64)
        if n:
            capi.check(capi.lib().bsq_fasta_to_host(self._h, None, C.c_void_p(ss.ctypes.data), None, C.c_void_p(ist.ctypes.data)),
                       self._h, "bsq_fasta_to_host")
            ids = np.zeros(max(int(ist[n]), 1), np.uint8)
            capi.check(capi.lib().bsq_fasta_to_host(self._h, C.c_void_p(seq.ctypes.data), C.c_void_p(ss.ctypes.data),
                                                    C.c_void_p(ids.ctypes.data), C.c_void_p(ist.ctypes.data)), self._h,
                       "bsq_fasta_to_host")
        return seq, ss, ids, ist

    def write_records(self, first_record: int = 0, count: Optional[int] = None, out_device_ptr: int = 0, capacity: int = 0,
                      to_host: bool = True, want_offsets: bool = False):
        """bsq_write_records: records [first_record, first_record + count) of the last batches() pass as FASTQ text, written by
        the device.  Returns (text as np.uint8 or None, offsets as np.uint64[count + 1] or None); with out_device_ptr the
        text also (or only, to_host=False) lands in the caller's device buffer."""
        if count is None:
            count = int(self.result.n_records) - first_record
        n = C.c_uint64(0)
        offs = np.zeros(count + 1, np.uint64) if want_offsets else None
        op = C.c_void_p(offs.ctypes.data) if offs is not None else C.c_void_p(0)
        # size first, then the text
        capi.check(capi.lib().bsq_write_records(self._h, first_record, count, C.c_void_p(0), 0, C.c_void_p(0), op, C.byref(n)),
                   self._h, "bsq_write_records")
        text = np.empty(int(n.value), np.uint8) if to_host else None
        if to_host or out_device_ptr:
            capi.check(capi.lib().bsq_write_records(
                self._h, first_record, count, C.c_void_p(out_device_ptr), capacity,
                C.c_void_p(text.ctypes.data if (text is not None and text.size) else 0), op, C.byref(n)), self._h, "bsq_write_records")
        return text, offs

    def quality_sums(self, first_record: int = 0, count: Optional[int] = None, out_device_ptr: int = 0) -> np.ndarray:
        """Per-record sum of Phred scores of the last batches() pass, computed on the device from the SoA
        (bsq_quality_sums): the parse -> consumer hand-off without a host round trip."""
        if count is None:
            count = int(self.result.n_records) - first_record
        out = np.zeros(max(count, 0), np.int32)
        capi.check(capi.lib().bsq_quality_sums(self._h, first_record, count, C.c_void_p(out_device_ptr or None),
                                               C.c_void_p(out.ctypes.data if count > 0 else None)), self._h)
        return out

    def timing(self):
        ms = (C.c_float * 5)()
        n = C.c_int64()
        capi.check(capi.lib().bsq_last_timing(self._h, C.byref(ms), C.byref(n)), self._h)
        return list(ms), int(n.value)

    def synth_device(self, dev_ptr: int, capacity: int, num_reads, first, count, mn, mx, min_phred, max_phred,
                     schema: QualitySchema) -> int:
        w = C.c_uint64()
        capi.check(capi.lib().bsq_synth_device(self._h, C.c_void_p(dev_ptr), capacity, num_reads, first, count, mn,
                                               mx, min_phred, max_phred, schema.LOWER, schema.UPPER,
                                               schema.OFFSET, C.byref(w)), self._h, "bsq_synth_device")
        return int(w.value)

    def stream_open(self, path: str, kind: int = capi.SOURCE_AUTO, region_bytes: int = 0):
        h = C.c_void_p()
        capi.check(capi.lib().bsq_stream_open(self._h, os.fspath(path).encode(), kind, region_bytes, C.byref(h)),
                   self._h, "bsq_stream_open")
        return h

    def stream_next_result(self, stream, want: int):
        """bsq_stream_next without touching the region's bytes (a region inflated on the device stays there)."""
        r = capi.PassResult()
        self.generation += 1
        capi.check(capi.lib().bsq_stream_next(stream, want, C.byref(r)), self._h, "bsq_stream_next")
        self.result = r
        return r, None, None

    def stream_next(self, stream, want: int):
        """Parses the next region of a file stream; returns (PassResult, region bytes as a numpy view,
        stream offset of the region, records before it)."""
        r = capi.PassResult()
        self.generation += 1
        capi.check(capi.lib().bsq_stream_next(stream, want, C.byref(r)), self._h, "bsq_stream_next")
        self.result = r
        n, off, first = self.stream_region_info(stream)
        return r, self.stream_region_bytes(stream), off, first

    def stream_region_info(self, stream):
        """(bytes, stream offset, records before it) of the region last parsed, without fetching its bytes."""
        n, off, first = C.c_uint64(), C.c_int64(), C.c_int64()
        capi.lib().bsq_stream_region_info(stream, C.byref(n), C.byref(off), C.byref(first))
        return int(n.value), int(off.value), int(first.value)

    def stream_region_bytes(self, stream) -> np.ndarray:
        """The region's bytes as a numpy view (a device-inflated region is copied to the host by this call)."""
        n, off, first = C.c_uint64(), C.c_int64(), C.c_int64()
        ptr = capi.lib().bsq_stream_region(stream, C.byref(n), C.byref(off), C.byref(first))
        if ptr and n.value:
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n.value,))
        return np.zeros(0, np.uint8)

    def stream_stats(self, stream) -> capi.StreamStats:
        st = capi.StreamStats()
        capi.check(capi.lib().bsq_stream_get_stats(stream, C.byref(st)), self._h)
        return st

    def stream_close(self, stream):
        capi.lib().bsq_stream_close(stream)

    def summarize_device(self, dev_ptr: int, n: int) -> capi.Summary:
        s = capi.Summary()
        capi.check(capi.lib().bsq_summarize_device(self._h, C.c_void_p(dev_ptr), n, C.byref(s)), self._h)
        return s


def shard_prefix(summaries, shard_bytes):
    """bsq_shard_prefix: where each byte-range shard's first own record starts (host arithmetic)."""
    n = len(summaries)
    arr = (capi.Summary * n)(*summaries)
    sizes = (C.c_uint64 * n)(*[int(b) for b in shard_bytes])
    out = (capi.ShardStart * n)()
    capi.check(capi.lib().bsq_shard_prefix(arr, sizes, n, out), None, "bsq_shard_prefix")
    return list(out)


# ------------------------------------------------------------------------------------------------
# FastqParser
# ------------------------------------------------------------------------------------------------


class _Region:
    """One pass: the host bytes it covered and the tables it produced."""

    def __init__(self, data, stream_offset: int, first_record: int, result: capi.PassResult, want: int,
                 is_last: bool = True, size: Optional[int] = None):
        # `data`: the region's bytes, or a function that fetches them (a region inflated on the device is only
        # copied to the host when somebody looks at its bytes)
        self._data, self.stream_offset, self.first_record, self.want = data, stream_offset, first_record, want
        self.size = int(size if size is not None else data.size)
        self.is_last = is_last
        self.n = int(result.n_records)
        self.stop = result.stop
        self.consumed = int(result.bytes_consumed)
        self.n_windows = int(result.n_windows)
        self.offsets = None  # (start[5][n] int64 absolute in region, id_start, id_len)

    @property
    def data(self) -> np.ndarray:
        if callable(self._data):
            self._data = self._data()
        return self._data


class FastqParser:
    """FastqParser[R, config] (parser.mojo:77-625).

    FastqParser(reader)                                   parser.mojo:89-107
    FastqParser(reader, quality_schema)                   parser.mojo:109-123
    FastqParser(reader, batch_size=..., schema="generic") parser.mojo:125-145
    The compile-time `config` parameter is the keyword `config`.
    """

    def __init__(self, reader: Reader, quality_schema: Optional[str] = None, *, batch_size: Optional[int] = None,
                 schema: str = "generic", config: Optional[ParserConfig] = None, device_id: int = 0,
                 region_bytes: int = 1 << 30, _force_id_slow_path: bool = False, native_io: bool = True,
                 host_inflate: bool = False):
        self.config = config or ParserConfig()
        if quality_schema is not None:                      # parser.mojo:117
            self.quality_schema = parse_schema(quality_schema)
        elif self.config.quality_schema:                    # parser.mojo:96-99,134-137
            self.quality_schema = parse_schema(self.config.quality_schema)
        elif batch_size is not None:                        # parser.mojo:139
            self.quality_schema = parse_schema(schema)
        else:                                               # parser.mojo:101
            self.quality_schema = parse_schema("generic")
        self._batch_size = batch_size if batch_size is not None else DEFAULT_BATCH_SIZE
        self._reader = reader
        self._region_bytes = max(int(region_bytes), 1)
        self._gpu = GpuParser(self.config.check_ascii, self.config.check_quality, self.quality_schema,
                              self._batch_size, device_id, self.config.buffer_capacity,
                              self.config.buffer_max_capacity, self.config.buffer_growth_enabled,
                              force_id_slow_path=_force_id_slow_path,
                              inflate_threads=(1 if type(reader) is GZFile else int(getattr(reader, "parallelism", 0) or 0)),
                              host_inflate=host_inflate)
        self._carry = np.zeros(0, np.uint8)   # unconsumed tail of the previous region
        self._stream_pos = 0                  # stream offset of _carry[0]
        self._records_done = 0                # records of finished regions
        self._reader_eof = False
        self._eof_seen = False                # BufferedReader._is_eof (buffered.mojo:278-279)
        self._region: Optional[_Region] = None
        self._cursor = 0                      # next record of the current region
        # file-backed readers go through the library's own pipeline: a reader thread fills pinned
        # regions (inflating .gz with zlib) while the GPU parses the previous one (bsq_stream_*)
        self._stream = None
        self._pending = None
        path = getattr(reader, "path", None)
        if native_io and path is not None and type(reader) in (FileReader, GZFile, RapidgzipReader):
            kind = capi.SOURCE_PLAIN if type(reader) is FileReader else capi.SOURCE_GZIP
            self._stream = self._gpu.stream_open(path, kind, self._region_bytes)
            self._stream_done = False
            if os.path.getsize(path) == 0:
                self._eof_seen = True     # BufferedReader.__init__ already read 0 bytes
        else:
            self._first_fill()

    # -- input ---------------------------------------------------------------------------------

    def _read_upto(self, want: int) -> np.ndarray:
        """Next region = carry + fresh bytes (BufferedReader._compact_from + _fill_buffer)."""
        parts = [self._carry] if self._carry.size else []
        have = self._carry.size
        chunk = 8 << 20
        while have < want and not self._reader_eof:
            buf = np.empty(min(chunk, want - have), np.uint8)
            got = self._reader.read_to_buffer(buf, buf.size, 0)
            if got == 0:
                self._reader_eof = True
                break
            parts.append(buf[:got])
            have += got
        if not parts:
            return np.zeros(0, np.uint8)
        return np.concatenate(parts) if len(parts) != 1 else parts[0]

    def _first_fill(self):
        # BufferedReader.__init__ reads once (buffered.mojo:149): an empty source is at EOF at once
        self._pending = self._read_upto(self._region_bytes)
        if self._pending.size == 0 and self._reader_eof:
            self._eof_seen = True

    def _load_region(self, want: int):
        if self._stream is not None:
            self._gpu.set_batch_size(self._batch_size)
            res, _, _ = self._gpu.stream_next_result(self._stream, want)
            stream, gpu, gen = self._stream, self._gpu, self._gpu.generation
            n, off, first = gpu.stream_region_info(stream)

            def fetch():
                if gpu.generation != gen:
                    raise BlazeSeqError("the region's bytes are gone: the parser has moved to another region")
                return gpu.stream_region_bytes(stream)
            is_last = res.stop.code != capi.OK
            reg = _Region(fetch, off, first, res, want | (0 if is_last else (capi.WANT_OFFSETS if want & capi.WANT_BATCHES else 0)),
                          is_last, size=n)
            self._stream_done = is_last
            self._region = reg
            self._cursor = 0
            return
        if self._pending is None:
            self._pending = self._read_upto(max(self._region_bytes, 2 * self._carry.size))
        data = self._pending
        self._pending = None
        if not self._reader_eof:
            # is the source exhausted exactly at the region end?  peek one chunk ahead
            probe = np.empty(1 << 16, np.uint8)
            got = self._reader.read_to_buffer(probe, probe.size, 0)
            if got == 0:
                self._reader_eof = True
            else:
                data = np.concatenate([data, probe[:got]])
        is_last = self._reader_eof
        if not is_last:
            want |= capi.WANT_OFFSETS  # the cut between regions needs record offsets
        data = np.ascontiguousarray(data)
        self._gpu.set_batch_size(self._batch_size)
        res = self._gpu.parse_host(data, self._stream_pos, self._records_done, is_last, want)
        reg = _Region(data, self._stream_pos, self._records_done, res, want, is_last)
        if not is_last and reg.stop.code == capi.OK and (want & capi.WANT_BATCHES) and reg.n % self._batch_size:
            # keep batches whole across regions: the records of the trailing partial batch are
            # re-presented with the next region (their bytes go back into the carry)
            # (all of them when the region holds fewer than one batch: the next region is read larger)
            keep = reg.n - reg.n % self._batch_size
            if keep > 0:
                self._ensure_offsets(reg)
                reg.consumed = int(reg.offsets[0][keep])
            else:
                reg.consumed = 0
            reg.n = keep
        self._region = reg
        self._cursor = 0

    def _advance_region(self):
        reg = self._region
        if self._stream is not None:      # the library carries the unconsumed tail itself
            self._region = None
            return
        self._carry = reg.data[reg.consumed:]
        self._stream_pos += reg.consumed
        self._records_done += reg.n
        self._region = None

    def _ensure_offsets(self, reg: _Region):
        if reg.offsets is not None:
            return
        if not (reg.want & capi.WANT_OFFSETS):
            # the pass was cut for batches only; run it again for the offsets table
            res = self._gpu.parse_host(np.ascontiguousarray(reg.data), reg.stream_offset, reg.first_record,
                                       reg.is_last, reg.want | capi.WANT_OFFSETS)
            reg.want |= capi.WANT_OFFSETS
            reg.n_windows = int(res.n_windows)
        cols = [[] for _ in range(7)]
        for w in range(reg.n_windows):
            v, le, sp = self._gpu.offsets_to_host(w)
            n = int(v.n_records)
            if n == 0:
                continue
            base = int(v.stream_base) - reg.stream_offset
            q = le.reshape(-1)[: 4 * n + 1]
            for k in range(4):  # u32 arithmetic: the leading sentinel is begin-1 and may wrap
                cols[k].append((q[k:4 * n:4] + np.uint32(1)).astype(np.int64) + base)
            cols[4].append(q[4:4 * n + 1:4].astype(np.int64) + base)
            cols[5].append(sp[0::2].astype(np.int64) + base)
            cols[6].append(sp[1::2].astype(np.int64))
        if cols[0]:
            reg.offsets = [np.concatenate(c) for c in cols]
        else:
            reg.offsets = [np.zeros(0, np.int64) for _ in range(7)]
        # one extra entry so that offsets[0][n] is where the next record would start
        reg.offsets[0] = np.append(reg.offsets[0], reg.consumed)

    # -- the reference API -----------------------------------------------------------------------

    def has_more(self) -> bool:
        """parser.mojo:156-157: buffer.available() > 0 or not buffer.is_eof()."""
        if self._region is not None:
            if self._cursor < self._region.n:
                return True
            if self._region.stop.code == capi.OK:
                return True
            unconsumed = self._region.size - self._region.consumed
            return unconsumed > 0 or not self._eof_seen
        if self._stream is not None:
            return not self._eof_seen
        return self._carry.size > 0 or (self._pending is not None and self._pending.size > 0) or not self._eof_seen

    def _raise_stop(self, stop: capi.Error):
        if stop.code == capi.EOF:
            self._eof_seen = True
            raise EOFError()
        raise BlazeSeqError(stop.text, stop.code, stop.record_number, stop.line_number, stop.file_position)

    def _next_index(self, want: int) -> int:
        """Index (in the current region) of the next record, loading regions as needed."""
        while True:
            if self._region is None:
                if self._stream is not None:
                    if self._stream_done or self._eof_seen and os.path.getsize(self._reader.path) == 0:
                        raise EOFError()
                elif self._eof_seen and self._carry.size == 0 and (self._pending is None or self._pending.size == 0):
                    raise EOFError()
                self._load_region(want)
            reg = self._region
            if self._cursor < reg.n:
                i = self._cursor
                self._cursor += 1
                return i
            if reg.stop.code == capi.OK:      # region exhausted, more input follows
                self._advance_region()
                continue
            self._raise_stop(reg.stop)

    def _view_at(self, reg: _Region, i: int) -> FastqView:
        self._ensure_offsets(reg)
        o = reg.offsets
        d = reg.data
        seq_s, sep_s, qual_s, end = int(o[1][i]), int(o[2][i]), int(o[3][i]), int(o[4][i])
        ids, idl = int(o[5][i]), int(o[6][i])
        return FastqView(d[ids:ids + idl], d[seq_s:sep_s - 1], d[qual_s:end], self.quality_schema.OFFSET)

    def next_view(self) -> FastqView:
        """parser.mojo:160-170."""
        i = self._next_index(capi.WANT_OFFSETS)
        return self._view_at(self._region, i)

    def next_record(self) -> FastqRecord:
        """parser.mojo:189-211."""
        if not self.has_more():
            raise EOFError()
        v = self.next_view()
        return FastqRecord(v.id(), v.sequence(), v.quality(), self.quality_schema.OFFSET)

    next_ref_as_record = next_record  # python/blazeseq_parser.mojo:136-150

    def next_batch(self, max_records: int = DEFAULT_BATCH_SIZE) -> FastqBatch:
        """parser.mojo:239-251: up to max_records records; EOF ends the batch, any other error is
        re-raised (the partially filled batch is lost, like in the reference)."""
        limit = max_records if max_records else self._batch_size
        # whole device batches when the cut lines up with the pass: a region whose records are used up (more
        # input follows) is left here, so that the next one is parsed for batches too
        while True:
            reg = self._region
            if reg is not None and self._cursor >= reg.n and reg.stop.code == capi.OK:
                self._advance_region()
            if self._region is None and self.has_more():
                self._batch_size = limit
                self._load_region(capi.WANT_BATCHES)
                if self._region.n == 0 and self._region.stop.code == capi.OK:
                    continue              # fewer records than one batch: read a larger region
            break
        reg = self._region
        if (reg is not None and (reg.want & capi.WANT_BATCHES) and limit == self._gpu.cfg.batch_size
                and self._cursor % limit == 0 and self._cursor < reg.n):
            take = min(limit, reg.n - self._cursor)
            short = take < limit
            if short and reg.stop.code not in (capi.OK, capi.EOF):
                self._cursor = reg.n
                self._raise_stop(reg.stop)            # error inside the batch: batch is lost
            if not short or reg.stop.code == capi.EOF:
                b = self._cursor // limit
                dev = DeviceFastqBatch(self._gpu, b, self._gpu.batch_view(b))
                self._cursor += take
                if short:
                    self._eof_seen = True             # the loop hit EOF (parser.mojo:248-249)
                return dev.copy_to_host()._with_device(dev)
        # general path: record by record, like the reference's loop
        batch = FastqBatch(batch_size=limit)
        while len(batch) < limit and self.has_more():
            try:
                batch.add(self.next_view())
            except EOFError:
                break
        return batch

    def views(self) -> Iterator[FastqView]:
        """_FastqParserViewIter (parser.mojo:628-661): EOF ends the iteration, any other error is
        printed and ends it too."""
        while True:
            try:
                yield self.next_view()
            except EOFError:
                return
            except BlazeSeqError as e:
                print(str(e))
                return

    def records(self) -> Iterator[FastqRecord]:
        """_FastqParserRecordIter (parser.mojo:664-697)."""
        while self.has_more():
            try:
                yield self.next_record()
            except EOFError:
                return
            except BlazeSeqError as e:
                print(str(e))
                return

    def batches(self, max_records: Optional[int] = None) -> Iterator[FastqBatch]:
        """_FastqParserBatchIter (parser.mojo:700-735)."""
        limit = max_records if max_records is not None else self._batch_size
        while self.has_more():
            try:
                batch = self.next_batch(limit)
            except BlazeSeqError as e:
                if "Record number:" in str(e):
                    print(str(e))
                return
            if len(batch) == 0:
                return
            yield batch

    def __iter__(self):
        return self.records()

    def close(self):
        if getattr(self, "_stream", None) is not None:
            self._gpu.stream_close(self._stream)
            self._stream = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _with_device(self: FastqBatch, dev: DeviceFastqBatch) -> FastqBatch:
    self._device = dev
    return self


FastqBatch._with_device = _with_device


# ------------------------------------------------------------------------------------------------
# python/blazeseq surface (python/blazeseq_parser.mojo:80-114)
# ------------------------------------------------------------------------------------------------


class FastqGZParser(FastqParser):
    pass


def parser(path: str, quality_schema: str = "generic", parallelism: int = 4) -> FastqParser:
    """blazeseq.parser(path, quality_schema, parallelism): reader chosen by suffix."""
    p = os.fspath(path)
    if p.endswith((".fastq.gz", ".fq.gz", ".fastq.bgz", ".fq.bgz")):
        return FastqGZParser(RapidgzipReader(p, parallelism), quality_schema)
    if p.endswith((".fastq", ".fq")):
        return FastqParser(FileReader(p), quality_schema)
    raise BlazeSeqError("Unsupported file extension: expected .fastq, .fq, .fastq.gz or .fq.gz")


create_parser = parser
