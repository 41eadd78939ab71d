#!/bin/bash
# One gpurun call: GPU parity tests, the default bench, and an ncu launch list of exactly one timed step.
timeout 500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12 > gpurun_out/t22.log
timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-ppl > gpurun_out/bench22_noppl.json 2> gpurun_out/bench22.err
GIFB200_CUPROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches22.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ppl-extra > gpurun_out/ncu22.log 2>&1
tail -3 gpurun_out/t22.log; python -c "
import json;d=json.load(open('gpurun_out/bench22_noppl.json'));print(d['value'],d['ms_per_step'],d['e2e'],d['roofline'])"
wc -l gpurun_out/launches22.csv
