#!/bin/bash
# GPU-box helper: parity suite, crowded / lone launches and the bench line without the CPU leg.
tag=${1:-quick}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/one_launch.py --B 8192 --reps 3 | tail -1
python tools/one_launch.py --B 1 --reps 3 | tail -1
python bench.py --no-cpu-baseline --steps 40 > gpurun_out/${tag}_bench.json
python -c "
import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print('value',round(d['value']),'ms/step', round(d['ms_per_step'],2),'e2e', round(d['e2e']['value']),'p50 batch', round(d['e2e']['p50_latency_ms_one_batch'],2),'p50 B=1', round(d['e2e']['p50_latency_ms_batch1'],3),'serial', round(d['one_batch_at_a_time']['value']),'conv', d['config']['converged_frac'])"
