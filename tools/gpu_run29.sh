#!/bin/bash
timeout 300 python -m pytest tests/test_flame_gpu.py -m gpu -q --tb=short 2>&1 | tail -5 > gpurun_out/t29.log
timeout 200 python tools/conv_probe.py > gpurun_out/probe29.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/t2_29 -f python tools/conv_probe.py t2 > gpurun_out/ncu29.log 2>&1
tail -3 gpurun_out/t29.log; cat gpurun_out/probe29.txt
