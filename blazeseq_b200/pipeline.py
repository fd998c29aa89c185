"""Host bytes -> host FastqBatch arrays with both PCIe directions busy.

`FastqParser.batches()` of the reference hands the caller HOST batches (record_batch.mojo:19-87).  One parser handle
does H2D + passes, then the copy of the SoA back to the host -- one after the other.  Two handles that alternate
regions overlap them: while handle A copies the SoA of region k back (D2H), handle B takes region k+1 (H2D + passes).
Regions are cut where the previous one's last whole batch ended (`BSQ_WANT_WHOLE_BATCHES`, the
BufferedReader._compact_from contract), so the batches are exactly the batches of the whole stream.
"""
from __future__ import annotations

import threading
from typing import Callable, Optional

import numpy as np

from . import _capi as capi
from .host import GpuParser


class HostBatchPipeline:
    """`make_parser()` -> GpuParser; both handles must be configured alike (schema, validation, batch size)."""

    def __init__(self, make_parser: Callable[[], GpuParser], region_bytes: int = 1 << 30):
        self.parsers = [make_parser(), make_parser()]
        self.region_bytes = int(region_bytes)
        self.stop: Optional[capi.Error] = None

    def close(self):
        for g in self.parsers:
            g.close()
        self.parsers = []

    def run(self, data: np.ndarray, seq, qual, idb, ends, id_ends, stream_offset: int = 0, first_record: int = 0):
        """Parses `data` (host bytes, pinned for full speed) and fills the five FastqBatch arrays of the whole input
        (numpy arrays or CPU torch tensors; `ends` / `id_ends` int64, rebased at every batch like FastqBatch).
        Returns (records, sequence bytes, quality bytes, id bytes); `self.stop` is the stop reason (EOF when clean)."""
        n = int(data.size)
        cond = threading.Condition()
        st = {"turn": 0, "pos": 0, "rec": 0, "so": 0, "qo": 0, "io": 0, "done": n == 0, "err": None}
        self.stop = None
        want = capi.WANT_BATCHES | capi.WANT_WHOLE_BATCHES

        def at(a, off):
            return None if a is None else a[off:]

        def worker(i: int):
            gpu = self.parsers[i]
            k = i
            try:
                while True:
                    with cond:
                        cond.wait_for(lambda: st["turn"] == k or st["done"] or st["err"] is not None)
                        if st["done"] or st["err"] is not None:
                            return
                        pos, rec = st["pos"], st["rec"]
                    end = min(n, pos + self.region_bytes)
                    last = end == n
                    r = gpu.parse_host(data[pos:end], stream_offset + pos, first_record + rec, last, want)
                    nrec = int(r.n_records)
                    if not last and r.stop.code == capi.OK and int(r.bytes_consumed) == 0:
                        raise ValueError("region_bytes holds less than one batch")
                    v = gpu.soa_view() if nrec else None
                    with cond:
                        so, qo, io = st["so"], st["qo"], st["io"]
                        st["pos"] = pos + int(r.bytes_consumed)
                        st["rec"] = rec + nrec
                        if v is not None:
                            st["so"] += int(v.sequence_bytes); st["qo"] += int(v.seq_len); st["io"] += int(v.total_id_bytes)
                        st["turn"] = k + 1
                        if last or r.stop.code != capi.OK:
                            st["done"] = True
                            self.stop = r.stop
                        cond.notify_all()
                    if nrec:
                        gpu.soa_to_host(at(seq, so), at(qual, qo), at(idb, io), at(ends, rec), at(id_ends, rec))
                    k += 2
            except BaseException as e:   # noqa: BLE001 - handed to the caller
                with cond:
                    st["err"] = e
                    cond.notify_all()

        threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if st["err"] is not None:
            raise st["err"]
        return st["rec"], st["so"], st["qo"], st["io"]
