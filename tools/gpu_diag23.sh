#!/bin/bash
# diagnostics: trainer test, per-kernel table of a normal eager step, ncu --set full of the memory-bound kernels
timeout 300 python -m pytest tests/test_trainer_gpu.py -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/t23.log
timeout 300 python tools/torch_profile_step.py normal > gpurun_out/step_kernels23.txt 2> gpurun_out/step_kernels23.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'upfirdn2d|torgb_fwd|bias_act' -c 8 -o gpurun_out/mem23 -f python tools/bench_ops.py memory > gpurun_out/ncu23.log 2>&1
tail -4 gpurun_out/t23.log; head -30 gpurun_out/step_kernels23.txt
