"""Host-side mirror of BlazeSeq's FASTA parser (blazeseq/fasta/parser.mojo, fasta/record.mojo) over the B200 C ABI.

The reference's FastaParser is a line loop; here the whole input goes through one device pass (bsq_fasta_parse_host:
newline table, line classification, two prefix sums, sequence packing) and the records are handed out from its
result.  Method names, error texts and EOF behaviour follow the reference."""
from __future__ import annotations

from typing import Iterator, Optional

import numpy as np

from . import _capi as capi
from .host import BlazeSeqError, EOFError, GpuParser, Reader  # noqa: A004

_SPACES = bytes([9, 10, 11, 12, 13, 28, 29, 30, 32])   # is_posix_space, utils.mojo:266-289


def _strip_spaces(b: bytes) -> bytes:
    return b.strip(_SPACES)


class FastaParserConfig:
    """fasta/parser.mojo:23-33."""

    def __init__(self, check_ascii: bool = False):
        self.check_ascii = check_ascii


class FastaRecord:
    """fasta/record.mojo:11-144: id (without '>') and the sequence as one line."""

    def __init__(self, id, sequence):
        self._id = id.encode("latin-1") if isinstance(id, str) else bytes(id)
        self._sequence = sequence.encode("latin-1") if isinstance(sequence, str) else bytes(sequence)

    def id(self) -> bytes:
        return self._id

    def sequence(self) -> bytes:
        return self._sequence

    def definition(self):
        """record.mojo:82-95: (Id, Description) -- the first blank-separated token and the rest."""
        parts = self._id.split(b" ")
        ident = parts[0].strip()
        if len(parts) > 1:
            return ident, _strip_spaces(b"".join(parts[1:]))
        return ident, None

    def byte_len(self) -> int:
        return 1 + len(self._id) + 1 + len(self._sequence) + 1

    def write(self, line_width: int = 60) -> bytes:
        w = line_width if line_width > 0 else len(self._sequence)
        out = [b">" + self._id + b"\n"]
        for i in range(0, len(self._sequence), max(w, 1)):
            out.append(self._sequence[i:i + w] + b"\n")
        return b"".join(out)

    def __len__(self) -> int:
        return len(self._sequence)

    def __eq__(self, other) -> bool:
        return isinstance(other, FastaRecord) and self._sequence == other._sequence

    def __ne__(self, other) -> bool:
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._sequence)

    def __repr__(self) -> str:
        return self.write().decode("latin-1")


class FastaParser:
    """FastaParser[R, config] (fasta/parser.mojo:60-200): next_record(), records(), has_more(), iteration."""

    def __init__(self, reader: Reader, config: Optional[FastaParserConfig] = None, device_id: int = 0):
        self.config = config or FastaParserConfig()
        parts = []
        while True:
            buf = np.empty(8 << 20, np.uint8)
            got = reader.read_to_buffer(buf, buf.size, 0)
            if got == 0:
                break
            parts.append(buf[:got])
        data = np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(0, np.uint8)
        self._gpu = GpuParser(check_ascii=self.config.check_ascii, device_id=device_id)
        self._res = self._gpu.fasta_parse_host(data)
        self._seq, self._ss, self._ids, self._ist = self._gpu.fasta_to_host()
        self._n = int(self._res.n_records)
        self._i = 0
        self._raised = False
        self._gpu.close()

    def has_more(self) -> bool:
        """parser.mojo:103-105."""
        return self._i < self._n or (not self._raised and self._res.stop.code != capi.EOF)

    def next_record(self) -> FastaRecord:
        """parser.mojo:123-172."""
        if self._i < self._n:
            i = self._i
            self._i += 1
            return FastaRecord(self._ids[int(self._ist[i]):int(self._ist[i + 1])].tobytes(),
                               self._seq[int(self._ss[i]):int(self._ss[i + 1])].tobytes())
        self._raised = True
        stop = self._res.stop
        if stop.code == capi.EOF:
            raise EOFError()
        raise BlazeSeqError(stop.text, stop.code, stop.record_number, stop.line_number, stop.file_position)

    def records(self) -> Iterator[FastaRecord]:
        """_FastaParserRecordIter (parser.mojo:208-244): EOF ends the iteration, any other error is printed first."""
        while self.has_more():
            try:
                yield self.next_record()
            except EOFError:
                return
            except BlazeSeqError as e:
                print(str(e))
                return

    __iter__ = records
