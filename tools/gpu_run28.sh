#!/bin/bash
timeout 500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12 > gpurun_out/t28.log
timeout 200 python tools/bench_ops.py raster memory > gpurun_out/ops28.jsonl 2> gpurun_out/ops28.err
tail -5 gpurun_out/t28.log; cut -c1-400 gpurun_out/ops28.jsonl; tail -3 gpurun_out/ops28.err
