#!/bin/bash
timeout 500 python -m pytest tests/test_texture_gpu.py tests/test_flame_gpu.py -m gpu -q --tb=short 2>&1 | tail -25 > gpurun_out/t31.log
timeout 200 python tools/bench_ops.py raster > gpurun_out/ops31.jsonl 2> gpurun_out/ops31.err
tail -25 gpurun_out/t31.log; cut -c1-330 gpurun_out/ops31.jsonl; tail -3 gpurun_out/ops31.err
