#!/bin/bash
timeout 300 python -m pytest tests/test_conv_tc_gpu.py tests/test_models_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -5 > gpurun_out/t34.log
timeout 200 python tools/wgrad_probe.py > gpurun_out/wprobe34.txt 2>&1
GIFB200_SHAPE_PROFILE=gpurun_out/shapes34.txt timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench34.json 2> gpurun_out/bench34.err
tail -3 gpurun_out/t34.log; cat gpurun_out/wprobe34.txt; python -c "
import json;d=json.load(open('gpurun_out/bench34.json'));print(d['value'],d['ms_per_step'],d['same_step_without_path_length_reg'])"; grep wgrad gpurun_out/shapes34.txt | head -8
