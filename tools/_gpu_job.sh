cd /root/repo
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/t31.log 2>&1; tail -5 gpurun_out/t31.log
timeout 120 python tools/sn_bench.py > gpurun_out/sn_bench5.log 2>&1; tail -4 gpurun_out/sn_bench5.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof_b256_23.json > gpurun_out/bench_g16.log 2>&1; grep -o "\"value\": [0-9.]*, \"unit\": \"img/s\", \"n_gpus\"" gpurun_out/bench_g16.log
