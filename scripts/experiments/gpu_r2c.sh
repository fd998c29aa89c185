# parity + bench of the trimmed resolve loop; tile-shape variants
set -x
mkdir -p gpurun_out
V=$PWD/blazeseq_b200/lib/variants
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 $B > gpurun_out/ab_def.json 2> gpurun_out/ab.err; show gpurun_out/ab_def.json default
timeout 600 $B --mode views > gpurun_out/ab_def_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_def_views.json default_views
timeout 600 $B --validate > gpurun_out/ab_def_val.json 2>> gpurun_out/ab.err; show gpurun_out/ab_def_val.json default_val
timeout 600 $B --mixed > gpurun_out/ab_def_mixed.json 2>> gpurun_out/ab.err; show gpurun_out/ab_def_mixed.json default_mixed
for t in t64k8c7 t128k8c6; do
BSQ_LIB=$V/lib_$t.so timeout 600 python -m pytest tests -m gpu -q -x --timeout=900 -k "corpus or random or synthetic" > gpurun_out/pytest_$t.log 2>&1; tail -2 gpurun_out/pytest_$t.log
BSQ_LIB=$V/lib_$t.so timeout 600 $B > gpurun_out/ab_$t.json 2>> gpurun_out/ab.err; show gpurun_out/ab_$t.json $t
BSQ_LIB=$V/lib_$t.so timeout 600 $B --mode views > gpurun_out/ab_${t}_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_${t}_views.json ${t}_views
done
tail -5 gpurun_out/ab.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 12 -c 1 -o gpurun_out/r2c_prof_resolve -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_resolve.log 2>&1
tail -2 gpurun_out/ncu_resolve.log
