import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'oracle'); sys.path.insert(0,'tests')
import blazeseq_b200 as B, oracle_py as O
from test_gpu_parity import _rand_stream
rng = np.random.default_rng(700 + len("notail"))
data = _rand_stream(rng, 400, "notail", maxlen=120)
views, bases, err = O.parse_all(data, O.config(True, True, "generic", buffer_growth_enabled=True))
print(len(views), err.text, len(data), data[-40:])
p = B.FastqParser(B.MemoryReader(data), config=B.ParserConfig(check_ascii=True, check_quality=True, buffer_growth_enabled=True), region_bytes=700)
n=0
try:
    while True:
        v=p.next_view(); n+=1
except Exception as e:
    print("stopped after", n, type(e).__name__, repr(str(e)), getattr(e,'code',None))
g = B.GpuParser(True, True, B.parse_schema("generic"), 4096, buffer_growth_enabled=True)
tail = np.frombuffer(data, np.uint8)[int(views[398]["record_end"])+1:]
print("tail bytes", bytes(tail))
r = g.parse_host(np.ascontiguousarray(tail), int(views[398]["record_end"])+1, 399, True, 1)
print(r.n_records, r.stop.code, repr(r.stop.text), r.bytes_consumed, r.n_newlines)
