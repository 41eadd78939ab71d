#!/bin/bash
# One gpurun call: GPU parity suite + a short step bench with the per-shape conv / wgrad table.
# usage: gpurun --timeout 900 -- 'bash tools/gpu_checks_step.sh TAG'
T=${1:-run}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 600 > gpurun_out/${T}_tests.log 2>&1
GIFB200_SHAPE_PROFILE=gpurun_out/${T}_shapes.txt python bench.py --no-extras --no-gpu-reference --no-cpu-baseline --no-e2e --no-other-precision > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
grep -n "^E  \|passed\|failed" gpurun_out/${T}_tests.log | cut -c1-300 | head -20
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['same_step_without_path_length_reg'],d['gpu_launches'],d['clocks'])"
grep "wgrad" gpurun_out/${T}_shapes.txt | grep ", 4, 4,\|, 8, 8,\|, 16, 16,\|, 9, 9,\|, 17, 17,\|, 33, 33,"
