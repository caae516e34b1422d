cd /root/repo
timeout 400 python bench.py > gpurun_out/bench_default3.log 2>&1; tail -1 gpurun_out/bench_default3.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1; tail -2 gpurun_out/smoke3.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 900 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 3 --quick --no-graph --no-cpu-baseline > gpurun_out/ncu_list4.log 2>&1
tail -1 gpurun_out/ncu_list4.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tc_fwd_kernel|tc_wgrad_kernel|tc_fwd_halo_kernel|attn_fwd_kernel|attn_bwd_dkdv" -s 60 -c 14 -o gpurun_out/prof_r1_final -f python bench.py --steps 1 --warmup 3 --quick --no-graph --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -1 gpurun_out/ncu_full2.log; ls -la gpurun_out/prof_r1_final.ncu-rep
