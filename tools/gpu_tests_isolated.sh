#!/bin/bash
# Runs each GPU test function in its own process (a sticky CUDA error cannot poison the others).
# usage: tools/gpu_tests_isolated.sh [logfile]
LOG=${1:-gpurun_out/isolated.log}
mkdir -p "$(dirname "$LOG")"; : > "$LOG"
for k in test_spectral_norm test_conv test_deconv test_fully_connected test_condition_batch_norm test_batch_norm \
         test_self_attention test_resblock_up test_resblock_down "test_resblock and not up and not down" test_hinge test_logistic test_adam \
         test_cuda test_train_step test_parameters_after; do
  echo "=== $k" >> "$LOG"
  timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "$k" >> "$LOG" 2>&1
done
grep -E "^(=== |FAILED|[0-9]+ (passed|failed))|passed|failed" "$LOG" | grep -v "^E " | tail -60
