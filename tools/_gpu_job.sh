cd /root/repo
timeout 60 tools/micro/mma_bench > gpurun_out/mma_bench2.log 2>&1; grep "elect" gpurun_out/mma_bench2.log | head -12
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/t24.log 2>&1; tail -6 gpurun_out/t24.log
lb() { timeout 60 python tools/layer_bench.py "$@" 2>&1 | tail -1; }
{
lb --op fwd --Ci 96 --Co 16
lb --op fwd --Ci 96 --Co 16 --pad zero
BG_TC_HALO=0 lb --op fwd --Ci 96 --Co 16 --pad zero
lb --op fwd --Ci 96 --Co 96 --deconv
BG_TC_HALO=0 lb --op fwd --Ci 96 --Co 96 --deconv
lb --op wgrad --Ci 96 --Co 96 --deconv
lb --op fwd --H 8 --Ci 1536 --Co 1536 --deconv
lb --op wgrad --H 8 --Ci 1536 --Co 1536 --deconv
lb --op fwd --H 64 --Ci 192 --Co 96 --k 4 --stride 2 --deconv
lb --op dgrad --H 64 --Ci 96 --Co 16 --k 1 --B 512 --residual
} > gpurun_out/layer_bench5.log 2>&1
cat gpurun_out/layer_bench5.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof_b256_18.json > gpurun_out/bench_g11.log 2>&1; grep -o "\"value\": [0-9.]*, \"unit\": \"img/s\", \"n_gpus\"" gpurun_out/bench_g11.log
