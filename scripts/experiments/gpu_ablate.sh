set -x
mkdir -p gpurun_out
for d in 0 1 3; do
BSQ_DEBUG_SKIP=$d timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --gib 4 > gpurun_out/abl_$d.json 2>gpurun_out/abl_$d.err; python -c "
import json;d=json.load(open('gpurun_out/abl_$d.json'));r=d['roofline']
print('skip=$d resolve %.3f ms/launch  summarize %.2f ms/step'%(r['avg_launch_ms'],r['summarize_ms_per_step']))"
done
