# single-pass (fused look-back) kernel: parity, two-pass cross-check, benches, profile
set -x
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
BSQ_TWO_PASS=1 timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "corpus or random or synthetic or degenerate" > gpurun_out/pytest_gpu_2p.log 2>&1; tail -3 gpurun_out/pytest_gpu_2p.log
timeout 600 $B > gpurun_out/ab_fused.json 2> gpurun_out/ab.err; show gpurun_out/ab_fused.json fused
BSQ_TWO_PASS=1 timeout 600 $B > gpurun_out/ab_2p.json 2>> gpurun_out/ab.err; show gpurun_out/ab_2p.json two_pass
timeout 600 $B --mode views > gpurun_out/ab_fused_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_fused_views.json fused_views
timeout 600 $B --validate > gpurun_out/ab_fused_val.json 2>> gpurun_out/ab.err; show gpurun_out/ab_fused_val.json fused_val
timeout 600 $B --mixed > gpurun_out/ab_fused_mixed.json 2>> gpurun_out/ab.err; show gpurun_out/ab_fused_mixed.json fused_mixed
tail -5 gpurun_out/ab.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 12 -c 1 -o gpurun_out/r2d_prof_resolve -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_resolve.log 2>&1
tail -2 gpurun_out/ncu_resolve.log
