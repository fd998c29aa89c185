# two-level look-back: parity of the single-pass kernel + benches against the two-pass default
set -x
mkdir -p gpurun_out
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  tail %.2f step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['tail_rebase_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "single_pass" > gpurun_out/pytest_sp.log 2>&1; tail -5 gpurun_out/pytest_sp.log
BSQ_SINGLE_PASS=1 timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 -k "not stream and not bgzf" > gpurun_out/pytest_sp_all.log 2>&1; tail -5 gpurun_out/pytest_sp_all.log
BSQ_SINGLE_PASS=1 timeout 600 $B > gpurun_out/ab_sp.json 2> gpurun_out/ab.err; show gpurun_out/ab_sp.json single_pass
BSQ_SINGLE_PASS=1 timeout 600 $B --mode views > gpurun_out/ab_sp_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_sp_views.json single_pass_views
BSQ_SINGLE_PASS=1 timeout 600 $B --validate > gpurun_out/ab_sp_val.json 2>> gpurun_out/ab.err; show gpurun_out/ab_sp_val.json single_pass_val
BSQ_SINGLE_PASS=1 timeout 600 $B --mixed > gpurun_out/ab_sp_mixed.json 2>> gpurun_out/ab.err; show gpurun_out/ab_sp_mixed.json single_pass_mixed
timeout 600 $B > gpurun_out/ab_def.json 2>> gpurun_out/ab.err; show gpurun_out/ab_def.json two_pass
tail -5 gpurun_out/ab.err
BSQ_SINGLE_PASS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 6 -c 1 -o gpurun_out/r2h_prof_sp -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_sp.log 2>&1
tail -2 gpurun_out/ncu_sp.log
