cd /root/repo
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t38.log 2>&1; tail -3 gpurun_out/t38.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof_b256_29.json > gpurun_out/bench_g22.log 2>&1; grep -o "\"value\": [0-9.]*, \"unit\": \"img/s\", \"n_gpus\"" gpurun_out/bench_g22.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 32 > gpurun_out/bench_g22_b32.log 2>&1; grep -o "\"value\": [0-9.]*, \"unit\": \"img/s\", \"n_gpus\"" gpurun_out/bench_g22_b32.log
