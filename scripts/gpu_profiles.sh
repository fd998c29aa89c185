# the committed evidence of profiles/: ncu launch list of the bench command + one `ncu --set full` capture per hot kernel,
# summarised on the box (the reports are ~10 MB each; gpurun brings back at most 64 MiB) -- gpurun -- 'bash scripts/gpu_profiles.sh'
mkdir -p gpurun_out
R=${ROUND_TAG:-r02}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${R}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|NUMA" >> gpurun_out/${R}_gpu.txt
nvidia-smi topo -m >> gpurun_out/${R}_gpu.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-sub > gpurun_out/${R}_ncu_launch.log 2>&1
# one window of 1.9 GiB per pass: launches per pass = 1 k_summarize + 1 k_resolve; 3 warm-up passes, the 4th is captured
N="ncu --set full --clock-control none --import-source on -s 3 -c 1 -f"
P="python bench.py --gib 1.9 --steps 1 --warmup 3 --no-cpu --no-e2e --no-sub"
cap() {  # name kernel-regex args...
  name=$1; kern=$2; shift 2
  timeout 300 $N -k regex:$kern -o gpurun_out/${R}_prof_$name $P "$@" > gpurun_out/ncu_$name.log 2>&1
  timeout 120 python scripts/profile_summary.py gpurun_out/${R}_prof_$name.ncu-rep $kern > gpurun_out/${R}_$name.txt 2>&1
}
cap k_resolve k_resolve
cap k_summarize k_summarize
timeout 120 python scripts/traffic_from_ncu.py gpurun_out/${R}_prof_k_resolve.ncu-rep gpurun_out/${R}_prof_k_summarize.ncu-rep 6394930 > gpurun_out/traffic.json 2>> gpurun_out/ncu_k_resolve.log
cap k_resolve_stride320 k_resolve --id-digits 9
cap k_resolve_validate k_resolve --validate
cap k_summarize_validate k_summarize --validate
cap k_resolve_mixed k_resolve --mixed
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inflate_members -s 3 -c 1 -f -o gpurun_out/${R}_prof_k_inflate python bench.py --gzip --gib 1 --region-mib 256 > gpurun_out/ncu_k_inflate.log 2>&1
timeout 120 python scripts/profile_summary.py gpurun_out/${R}_prof_k_inflate.ncu-rep k_inflate > gpurun_out/${R}_k_inflate_members.txt 2>&1
# keep two reports for reading source pages later; the rest stays on the box
find gpurun_out -name "*.ncu-rep" ! -name "${R}_prof_k_resolve.ncu-rep" ! -name "${R}_prof_k_inflate.ncu-rep" -delete
ls -la gpurun_out | tail -20; du -sh gpurun_out
