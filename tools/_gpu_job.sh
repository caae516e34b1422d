cd /root/repo
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/t36.log 2>&1; tail -6 gpurun_out/t36.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof_b256_27.json > gpurun_out/bench_g20.log 2>&1; grep -o "\"value\": [0-9.]*, \"unit\": \"img/s\", \"n_gpus\"" gpurun_out/bench_g20.log
BG_TC_FOLD=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g20_nofold.log 2>&1; grep -o "\"value\": [0-9.]*, \"unit\": \"img/s\", \"n_gpus\"" gpurun_out/bench_g20_nofold.log
