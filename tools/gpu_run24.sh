#!/bin/bash
timeout 500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -12 > gpurun_out/t24.log
timeout 200 python tools/bench_ops.py memory > gpurun_out/mem24.jsonl 2> gpurun_out/mem24.err
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-ppl --no-e2e > gpurun_out/bench24_noppl.json 2> gpurun_out/bench24.err
tail -3 gpurun_out/t24.log; cut -c1-200 gpurun_out/mem24.jsonl; python -c "
import json;d=json.load(open('gpurun_out/bench24_noppl.json'));print(d['value'],d['ms_per_step'])"
