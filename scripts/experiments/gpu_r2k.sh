# micro-experiments on k_resolve (same box, back to back): default vs e1 (noinline checked fill), e2 (8-byte stage stores),
# e4 (full-tile fast path in the bitmap builder), all three
set -x
mkdir -p gpurun_out
V=$PWD/blazeseq_b200/lib/variants
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e"
timeout 600 python -m pytest tests -m gpu -q -x --timeout=600 -k "quality_sums" > gpurun_out/pytest_new.log 2>&1; tail -4 gpurun_out/pytest_new.log
for t in def e1 e2 e4 e124 def; do
L=""; [ $t != def ] && L=$V/lib_$t.so
BSQ_LIB=$L timeout 600 $B > gpurun_out/ab_$t.json 2>> gpurun_out/ab.err; show gpurun_out/ab_$t.json $t
done
BSQ_LIB=$V/lib_e124.so timeout 600 $B --mode views > gpurun_out/ab_e124_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_e124_views.json e124_views
timeout 600 $B --mode views > gpurun_out/ab_def_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_def_views.json def_views
BSQ_LIB=$V/lib_e124.so timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 -k "not stream and not bgzf and not single_pass" > gpurun_out/pytest_e124.log 2>&1; tail -3 gpurun_out/pytest_e124.log
tail -3 gpurun_out/ab.err
