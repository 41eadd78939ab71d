#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -6 > gpurun_out/t35.log
timeout 200 python tools/bench_ops.py memory > gpurun_out/mem35.jsonl 2> gpurun_out/mem35.err
timeout 300 python tools/torch_profile_step.py normal ppl > gpurun_out/step_kernels35.txt 2> gpurun_out/step_kernels35.err
tail -3 gpurun_out/t35.log; grep torgb gpurun_out/mem35.jsonl | cut -c1-200; head -34 gpurun_out/step_kernels35.txt | cut -c1-150
