#!/bin/bash
# round-2 GPU call N: re-check the dw-mode math-warp changes per layer (2 repeats) + the tests that cover them
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_16bit.py -q -x > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2n_tests.log
for i in 1 2; do
  timeout 200 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2n_steps_$i.txt > /dev/null 2>&1
  echo "rep $i: $(tail -1 gpurun_out/r2n_steps_$i.txt | cut -c1-60)"
done
grep -E "^deconv0_0|^deconv1_0|^cnn0_last|^deconv0_1|^deconv1_1|^deconv2_0|^cnn2_last|^deconv2_1|^cnn1 |^final|^residual0_d" gpurun_out/r2n_steps_1.txt
