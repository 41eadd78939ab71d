#!/bin/bash
timeout 500 python -m pytest tests/test_conv_tc_gpu.py tests/test_models_gpu.py tests/test_ops_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -6 > gpurun_out/t30.log
timeout 200 python tools/conv_probe.py > gpurun_out/probe30.txt 2>&1
GIFB200_SHAPE_PROFILE=gpurun_out/shapes30.txt timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e --no-ppl > gpurun_out/bench30.json 2> gpurun_out/bench30.err
tail -3 gpurun_out/t30.log; cat gpurun_out/probe30.txt; python -c "
import json;d=json.load(open('gpurun_out/bench30.json'));print(d['value'],d['ms_per_step'],d['roofline']['achieved'],d['roofline']['all_tensor_core_conv_launches'])"
