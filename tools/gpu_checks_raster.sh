#!/bin/bash
# One gpurun call: GPU parity suite, the rasteriser microbenchmark, an `ncu --set full` capture of raster_tile_kernel, and a
# short step bench.  usage: gpurun --timeout 900 -- 'bash tools/gpu_checks_raster.sh TAG'
T=${1:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short --timeout 600 > gpurun_out/${T}_tests.log 2>&1
python tools/bench_ops.py raster > gpurun_out/${T}_raster.jsonl 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:raster_tile_kernel -c 3 -o gpurun_out/${T}_raster_tile python tools/bench_ops.py raster > gpurun_out/${T}_ncu.log 2>&1
python bench.py --no-extras --no-gpu-reference --no-cpu-baseline --no-e2e --no-other-precision > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
grep -n "^E  \|passed\|failed" gpurun_out/${T}_tests.log | cut -c1-300 | head -20
head -c 700 gpurun_out/${T}_raster.jsonl
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['same_step_without_path_length_reg'],d['gpu_launches'],d['clocks'])"
