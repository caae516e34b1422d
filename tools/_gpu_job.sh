cd /root/repo
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t33.log 2>&1; tail -3 gpurun_out/t33.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; tail -2 gpurun_out/smoke2.log
timeout 400 python bench.py > gpurun_out/bench_default2.log 2>&1; tail -1 gpurun_out/bench_default2.log | cut -c1-600
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref2.log 2>&1; tail -1 gpurun_out/bench_ref2.log | cut -c1-300
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1000 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 3 --quick --no-graph --no-cpu-baseline > gpurun_out/ncu_list3.log 2>&1
tail -1 gpurun_out/ncu_list3.log
