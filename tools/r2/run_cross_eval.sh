#!/bin/bash
# DESIGN.md section 7.1, experiments (1)-(2): one gpurun call (about 6 GPU-minutes; the oracle legs run on the box's CPU).
#   gpurun --timeout 900 -- 'bash tools/r2/run_cross_eval.sh'
mkdir -p gpurun_out
T="timeout -s KILL"
$T 200 python -m pytest tests/test_gpu_training.py -m gpu -q --timeout 150 -k "every_held_out_view or trained_model_matches_oracle" > gpurun_out/r3_pytest_every_view.log 2>&1
echo "every-view parity rc=$?"; tail -3 gpurun_out/r3_pytest_every_view.log | cut -c1-300
for seed in 0 1; do
  rm -rf /tmp/run_s$seed
  $T 120 python -m fruitnerf_b200.scripts.train --steps 3000 --seed $seed --output-dir /tmp/run_s$seed --json gpurun_out/r3_train_seed$seed.json > gpurun_out/r3_train_seed$seed.log 2>&1
  CK=$(ls /tmp/run_s$seed/nerfstudio_models/step-*.ckpt | tail -1)
  $T 60 python tools/cross_eval.py --ckpt $CK --seed $seed --through kernels > gpurun_out/r3_cross_kernels_seed$seed.jsonl 2>&1
  $T 300 python tools/cross_eval.py --ckpt $CK --seed $seed --through oracle --threads 32 > gpurun_out/r3_cross_oracle_seed$seed.jsonl 2>&1
  echo "seed $seed: kernel-trained weights through the kernels / through the oracle"; tail -1 gpurun_out/r3_cross_kernels_seed$seed.jsonl; tail -1 gpurun_out/r3_cross_oracle_seed$seed.jsonl
done
