#!/bin/bash
timeout 500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12 > gpurun_out/t25.log
timeout 200 python tools/bench_ops.py memory > gpurun_out/mem25.jsonl 2> gpurun_out/mem25.err
GIFB200_SHAPE_PROFILE=gpurun_out/shapes25.txt timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench25.json 2> gpurun_out/bench25.err
tail -3 gpurun_out/t25.log; cut -c1-200 gpurun_out/mem25.jsonl | head -3; python -c "
import json;d=json.load(open('gpurun_out/bench25.json'));print(d['value'],d['ms_per_step'],d['same_step_without_path_length_reg'])"
