#!/bin/bash
# A/B on the GPU box: C3 (interpreted DSP list) with variants/libicw_b200_prev.so against the build in the tree
cp in_cwave_b200/libicw_b200.so /tmp/base.so
for v in interp3 base interp3 base; do
  if [ $v = base ]; then cp /tmp/base.so in_cwave_b200/libicw_b200.so; else cp variants/libicw_b200_$v.so in_cwave_b200/libicw_b200.so; fi
  python bench.py --workload c3 --no-workloads --no-cpu --no-e2e --no-parity --steps 20 --warmup 5 > gpurun_out/c3_$v.json 2> gpurun_out/c3_$v.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/c3_$v.json').read().splitlines() if l.startswith('{')][-1]); print('$v', d['ms_per_step'], d['value'])"
done
cp /tmp/base.so in_cwave_b200/libicw_b200.so
