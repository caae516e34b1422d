cd /root/repo
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t29.log 2>&1; tail -12 gpurun_out/t29.log
