mkdir -p gpurun_out
rm -rf gpurun_out/*
timeout 900 python -m pytest tests/test_gpu_fasta.py -m gpu -q -x --timeout=600 > gpurun_out/pytest_fasta.log 2>&1; tail -12 gpurun_out/pytest_fasta.log
bash scripts/gpu_profiles.sh
