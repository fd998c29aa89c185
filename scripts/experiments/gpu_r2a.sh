# parity suite + A/B of the staged (TMA store) SoA copy against the direct copy, and the 1-stage variant
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 600 $B > gpurun_out/ab_staged.json 2> gpurun_out/ab.err; show gpurun_out/ab_staged.json staged_s2c3
BSQ_DEBUG_SKIP=4 timeout 600 $B > gpurun_out/ab_direct.json 2>> gpurun_out/ab.err; show gpurun_out/ab_direct.json direct_s2c3
BSQ_DEBUG_SKIP=1 timeout 600 $B > gpurun_out/ab_nocopy.json 2>> gpurun_out/ab.err; show gpurun_out/ab_nocopy.json nocopy_s2c3
BSQ_LIB=$PWD/blazeseq_b200/lib/variants/lib_s1c4.so timeout 600 $B > gpurun_out/ab_s1c4.json 2>> gpurun_out/ab.err; show gpurun_out/ab_s1c4.json staged_s1c4
BSQ_LIB=$PWD/blazeseq_b200/lib/variants/lib_s1c4.so timeout 600 $B --validate > gpurun_out/ab_s1c4_val.json 2>> gpurun_out/ab.err; show gpurun_out/ab_s1c4_val.json staged_s1c4_validate
timeout 600 $B --validate > gpurun_out/ab_val.json 2>> gpurun_out/ab.err; show gpurun_out/ab_val.json staged_s2c3_validate
timeout 600 $B --mode views > gpurun_out/ab_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_views.json views_s2c3
tail -5 gpurun_out/ab.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 12 -c 1 -o gpurun_out/r2a_prof_resolve -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_resolve.log 2>&1
tail -3 gpurun_out/ncu_resolve.log
ls -la gpurun_out | tail -20
