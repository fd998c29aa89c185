# A/B of CTA shapes (threads x stages x CTAs/SM) on the staged resolve kernel; parity on the two candidates
set -x
mkdir -p gpurun_out
V=$PWD/blazeseq_b200/lib/variants
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 600 $B > gpurun_out/ab_t128s1c4.json 2> gpurun_out/ab.err; show gpurun_out/ab_t128s1c4.json t128s1c4
for t in t256s2c3 t256s1c4; do
BSQ_LIB=$V/lib_$t.so timeout 600 $B > gpurun_out/ab_$t.json 2>> gpurun_out/ab.err; show gpurun_out/ab_$t.json $t
BSQ_LIB=$V/lib_$t.so timeout 600 $B --mode views > gpurun_out/ab_${t}_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_${t}_views.json ${t}_views
BSQ_LIB=$V/lib_$t.so timeout 600 $B --validate > gpurun_out/ab_${t}_val.json 2>> gpurun_out/ab.err; show gpurun_out/ab_${t}_val.json ${t}_val
done
timeout 600 $B --mode views > gpurun_out/ab_t128s1c4_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_t128s1c4_views.json t128s1c4_views
tail -5 gpurun_out/ab.err
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
BSQ_LIB=$V/lib_t256s1c4.so timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu_t256.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_t256.log
tail -5 gpurun_out/pytest_gpu_t256.log
BSQ_LIB=$V/lib_t256s1c4.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 12 -c 1 -o gpurun_out/r2b_prof_resolve -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_resolve.log 2>&1
tail -3 gpurun_out/ncu_resolve.log
