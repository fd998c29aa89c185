mkdir -p gpurun_out
timeout 900 python bench.py --gzip --gib 2 > gpurun_out/g_gzip2.json 2> gpurun_out/g.err; python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/g_gzip2.json"))["gzip"]
    for k, v in d.items():
        if isinstance(v, dict): print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
        else: print(k, v)
except Exception as e:
    print("failed", e)
P
tail -5 gpurun_out/g.err
