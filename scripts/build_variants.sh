# tuning variants of the library (same sources, different -D): blazeseq_b200/lib/variants/lib_<tag>.so
# usage: bash scripts/build_variants.sh s1c4:"-DBSQ_STAGES=1 -DBSQ_RESOLVE_CTAS=4" ...
set -e
cd "$(dirname "$0")/../blazeseq_b200/csrc"
mkdir -p ../lib/variants
for spec in "$@"; do
  tag=${spec%%:*}; defs=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared --cudart static $defs \
       -o ../lib/variants/lib_$tag.so bsq_capi.cu -lz -lpthread
  echo "built lib_$tag.so ($defs)"
done
