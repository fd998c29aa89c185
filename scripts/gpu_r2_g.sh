# round 2: device BGZF inflate -- bit-exactness, the stream tests, the gzip bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 -k "inflate or bgzf or stream or writer or binding or file_and_gzip or native_and_python or device_batches" > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --gzip --gib 2 > gpurun_out/g_gzip2.json 2> gpurun_out/g.err; python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/g_gzip2.json"))["gzip"]
    for k, v in d.items():
        if isinstance(v, dict): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
        else: print(k, v)
except Exception as e:
    print("failed", e)
P
tail -5 gpurun_out/g.err
