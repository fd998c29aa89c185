# CTA / tile shapes with 128 bytes per thread kept: 96 x 12 KiB x 6, 128 x 16 KiB x 5 (default), 160 x 20 KiB x 4, 192 x 24 KiB x 3
set -x
mkdir -p gpurun_out
V=$PWD/blazeseq_b200/lib/variants
show() { python -c "
import json,sys;d=json.load(open('$1'));r=d['roofline']
print('$2 value %.3g reads/s  ms/step %.2f  resolve %.3f ms/launch frac %.3f  summarize %.2f ms/step  step_frac %.3f'%(d['value'],d['ms_per_step'],r['avg_launch_ms'],r['frac'],r['summarize_ms_per_step'],r['step_frac']))"; }
B="python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e"
for t in def t96 t160 t192; do
L=""; [ $t != def ] && L=$V/lib_$t.so
BSQ_LIB=$L timeout 600 $B > gpurun_out/ab_$t.json 2>> gpurun_out/ab.err; show gpurun_out/ab_$t.json $t
BSQ_LIB=$L timeout 600 $B --mode views > gpurun_out/ab_${t}_views.json 2>> gpurun_out/ab.err; show gpurun_out/ab_${t}_views.json ${t}_views
done
for t in t96 t160 t192; do
BSQ_LIB=$V/lib_$t.so timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -k "corpus or random or degenerate or all_newlines or long_reads or scale" > gpurun_out/pytest_$t.log 2>&1; tail -2 gpurun_out/pytest_$t.log
done
tail -3 gpurun_out/ab.err
