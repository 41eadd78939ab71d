#!/bin/bash
timeout 500 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py tests/test_conv_tc_gpu.py -m gpu -q --tb=short 2>&1 | tail -12 > gpurun_out/t27.log
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e --no-ppl > gpurun_out/bench27.json 2> gpurun_out/bench27.err
timeout 300 python tools/torch_profile_step.py normal > gpurun_out/step_kernels27.txt 2> gpurun_out/step_kernels27.err
tail -3 gpurun_out/t27.log; python -c "
import json;d=json.load(open('gpurun_out/bench27.json'));print(d['value'],d['ms_per_step'])"; head -12 gpurun_out/step_kernels27.txt | cut -c1-150
