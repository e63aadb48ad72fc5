#!/bin/bash
# A/B on the GPU box: role order of the one-kernel scan path's warps (variants/libicw_b200_ordN.so, tools/build_variant.sh)
./tools/warpid_probe > gpurun_out/warpid_probe.txt 2>&1
cp in_cwave_b200/libicw_b200.so /tmp/base.so
for v in base ord1 ord2 ord3 base; do
  if [ $v = base ]; then cp /tmp/base.so in_cwave_b200/libicw_b200.so; else cp variants/libicw_b200_$v.so in_cwave_b200/libicw_b200.so; fi
  python bench.py --frames 276480000 --no-workloads --no-cpu --no-e2e --no-parity --steps 10 --warmup 3 > gpurun_out/ord_$v.json 2> gpurun_out/ord_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/ord_$v.json')); print('$v', d['ms_per_step'], d['value'])"
done
cp /tmp/base.so in_cwave_b200/libicw_b200.so
