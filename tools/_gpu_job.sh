cd /root/repo
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/t25.log 2>&1; tail -6 gpurun_out/t25.log
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
lb --op fwd --H 32 --Ci 384 --Co 192 --k 4 --stride 2 --deconv
lb --op dgrad --H 64 --Ci 96 --Co 16 --k 1 --B 512 --residual
lb --op fwd --Ci 16 --Co 96 --stride 2 --B 512
} > gpurun_out/layer_bench7.log 2>&1
cat gpurun_out/layer_bench7.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof_b256_19.json > gpurun_out/bench_g12.log 2>&1; grep -o "\"value\": [0-9.]*, \"unit\": \"img/s\", \"n_gpus\"" gpurun_out/bench_g12.log
