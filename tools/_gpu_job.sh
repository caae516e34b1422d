cd /root/repo
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/t35.log 2>&1; tail -3 gpurun_out/t35.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof_b256_26.json > gpurun_out/bench_g19.log 2>&1; grep -o "\"value\": [0-9.]*, \"unit\": \"img/s\", \"n_gpus\"" gpurun_out/bench_g19.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 32 --profile-json gpurun_out/prof_b32_6.json > gpurun_out/bench_g19_b32.log 2>&1; grep -o "\"value\": [0-9.]*, \"unit\": \"img/s\", \"n_gpus\"" gpurun_out/bench_g19_b32.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch 64 > gpurun_out/bench_g19_b64.log 2>&1; grep -o "\"value\": [0-9.]*, \"unit\": \"img/s\", \"n_gpus\"" gpurun_out/bench_g19_b64.log
