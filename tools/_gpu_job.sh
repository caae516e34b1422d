cd /root/repo
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/t37.log 2>&1; tail -8 gpurun_out/t37.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof_b256_28.json > gpurun_out/bench_g21.log 2>&1; grep -o "\"value\": [0-9.]*, \"unit\": \"img/s\", \"n_gpus\"" gpurun_out/bench_g21.log
