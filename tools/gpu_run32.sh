#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -8 > gpurun_out/t32.log
timeout 200 python tools/conv_probe.py > gpurun_out/probe32.txt 2>&1
GIFB200_SHAPE_PROFILE=gpurun_out/shapes32.txt timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench32.json 2> gpurun_out/bench32.err
tail -4 gpurun_out/t32.log; cat gpurun_out/probe32.txt; python -c "
import json;d=json.load(open('gpurun_out/bench32.json'));print(d['value'],d['ms_per_step'],d['same_step_without_path_length_reg'],d['roofline']['achieved'],d['roofline']['all_tensor_core_conv_launches'])"
