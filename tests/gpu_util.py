"""Helpers shared by the GPU parity tests: run a stream through the C ABI and through the oracle
and compare everything the path produces."""
from __future__ import annotations

import numpy as np

import blazeseq_b200 as B
from blazeseq_b200 import _capi as capi

NAMES5 = ("header_start", "seq_start", "sep_start", "qual_start", "record_end")


def make_gpu(check_ascii=False, check_quality=False, schema="generic", batch_size=4096, **kw):
    return B.GpuParser(check_ascii, check_quality, B.parse_schema(schema), batch_size, **kw)


def gpu_offsets(gpu: B.GpuParser, res):
    """All windows' offsets as absolute int64 columns (stream_offset 0)."""
    cols = {k: [] for k in NAMES5 + ("id_start", "id_len")}
    for w in range(res.n_windows):
        v, le, sp = gpu.offsets_to_host(w)
        n = int(v.n_records)
        if not n:
            continue
        base = int(v.stream_base)
        for k in range(4):  # u32 arithmetic: the leading sentinel is begin-1 and may wrap
            cols[NAMES5[k]].append((le[k:4 * n:4] + np.uint32(1)).astype(np.int64) + base)
        cols["record_end"].append(le[4:4 * n + 1:4].astype(np.int64) + base)
        cols["id_start"].append(sp[0::2].astype(np.int64) + base)
        cols["id_len"].append(sp[1::2].astype(np.int64))
    return {k: (np.concatenate(v) if v else np.zeros(0, np.int64)) for k, v in cols.items()}


def check_stream(oracle, data: bytes | np.ndarray, *, check_ascii=False, check_quality=False,
                 schema="generic", batch_size=4096, growth=False, gpu=None, via="host", torch_dev=None,
                 force_id_slow=False, want=capi.WANT_OFFSETS | capi.WANT_BATCHES, buffer_capacity=None,
                 buffer_max_capacity=None, compat_q5_width=0):
    """Parses `data` on the GPU through the C ABI and on the CPU with the oracle; asserts that the
    records, offsets, id spans, SoA batches, totals and the stop reason (code, context, text) agree.
    Returns the PassResult."""
    arr = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
    own = gpu is None
    if own:
        kw = {}
        if buffer_capacity is not None:
            kw["buffer_capacity"] = buffer_capacity
        if buffer_max_capacity is not None:
            kw["buffer_max_capacity"] = buffer_max_capacity
        gpu = B.GpuParser(check_ascii, check_quality, B.parse_schema(schema), batch_size,
                          buffer_growth_enabled=growth, force_id_slow_path=force_id_slow,
                          compat_q5_width=compat_q5_width, **kw)
    cfg = oracle.config(check_ascii, check_quality, schema, buffer_growth_enabled=growth,
                        buffer_capacity=buffer_capacity, buffer_max_capacity=buffer_max_capacity,
                        compat_simd_width=compat_q5_width)
    views, bases, err = oracle.parse_all(arr, cfg)
    if via == "host":
        res = gpu.parse_host(np.ascontiguousarray(arr), 0, 0, True, want)
    else:
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr).copy()).to(torch_dev or "cuda:0")
        res = gpu.parse_device(t.data_ptr(), t.numel(), 0, 0, True, want)
    n = len(views)
    assert res.n_records == n, (res.n_records, n, res.stop.text, err.text)
    assert res.stop.code == err.code, (res.stop.code, err.code, res.stop.text, err.text)
    assert res.stop.message == err.message, (res.stop.message, err.message)
    assert (res.stop.record_number, res.stop.line_number, res.stop.file_position) == \
        (err.record_number, err.line_number, err.file_position)
    assert res.n_newlines == int((arr == 10).sum())
    if want & capi.WANT_OFFSETS:
        g = gpu_offsets(gpu, res)
        for k in NAMES5 + ("id_start", "id_len"):
            assert g[k].shape[0] == n, (k, g[k].shape, n)
            # the oracle leaves id_start unspecified for empty ids only by its strip loop; both sides
            # run the same strip, so positions must agree exactly
            assert np.array_equal(g[k], views[k]), k
    if want & capi.WANT_BATCHES:
        m = batch_size
        assert res.n_batches == (n + m - 1) // m
        for b in range(int(res.n_batches)):
            seq, qual, idb, ends, id_ends = gpu.batch_to_host(b)
            oi, os_, oq, oie, oe = oracle.build_batch(arr, views[b * m:(b + 1) * m])
            assert np.array_equal(ends, oe), b
            assert np.array_equal(id_ends, oie), b
            assert np.array_equal(seq, os_), b
            assert np.array_equal(qual, oq), b
            assert np.array_equal(idb, oi), b
        if err.code == oracle.EOF or n > 0:
            assert res.n_bases == bases
    if own:
        gpu.close()
    return res
