# smem diet (validation bitmaps inside the stage buffer, 512-entry newline list): 4 vs 5 CTAs/SM
set -x
mkdir -p gpurun_out
V=$PWD/blazeseq_b200/lib/variants
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
BSQ_LIB=$V/lib_c5.so timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu_c5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c5.log
tail -5 gpurun_out/pytest_gpu_c5.log
for t in def c5; do
L=""; [ $t = c5 ] && L=$V/lib_c5.so
BSQ_LIB=$L timeout 600 $B > gpurun_out/ab_$t.json 2>> gpurun_out/ab.err; show gpurun_out/ab_$t.json $t
BSQ_LIB=$L timeout 600 $B --mode views > gpurun_out/ab_${t}_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_${t}_views.json ${t}_views
BSQ_LIB=$L timeout 600 $B --validate > gpurun_out/ab_${t}_val.json 2>> gpurun_out/ab.err; show gpurun_out/ab_${t}_val.json ${t}_val
BSQ_LIB=$L timeout 600 $B --mixed > gpurun_out/ab_${t}_mixed.json 2>> gpurun_out/ab.err; show gpurun_out/ab_${t}_mixed.json ${t}_mixed
done
tail -5 gpurun_out/ab.err
