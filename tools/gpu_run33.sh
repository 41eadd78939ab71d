#!/bin/bash
timeout 200 python tools/wgrad_probe.py > gpurun_out/wprobe33.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/wg_ns33 -f python tools/wgrad_probe.py ns > gpurun_out/ncu33.log 2>&1
cat gpurun_out/wprobe33.txt
