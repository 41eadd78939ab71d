#!/bin/bash
# One gpurun call at the end of a round: GPU parity tests, the default bench (+ the flagship-run variant), and an ncu launch
# list of exactly one timed step (bench.py brackets the timed region with cudaProfilerStart/Stop when GIFB200_CUPROFILE=1).
timeout 500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12 > gpurun_out/t_final.log
timeout 400 python bench.py --steps 16 --warmup 3 --texture-loss > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
GIFB200_CUPROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ppl-extra > gpurun_out/ncu_final.log 2>&1
tail -3 gpurun_out/t_final.log; python -c "
import json;d=json.load(open('gpurun_out/bench_final.json'));print(d['value'],d['ms_per_step'],d['e2e'],d['same_step_without_path_length_reg'],d['same_step_with_texture_interpolation_loss_instead_of_ppl'],d['cpu_baseline'],d['roofline']['achieved'],d['roofline']['share_of_step'],d['gpu_launches'],d['clocks'])"
wc -l gpurun_out/launches_final.csv
