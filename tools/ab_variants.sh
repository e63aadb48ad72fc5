cp in_cwave_b200/libicw_b200.so /tmp/base.so
for v in base old; do
  if [ $v = base ]; then cp /tmp/base.so in_cwave_b200/libicw_b200.so; else cp variants/libicw_b200_$v.so in_cwave_b200/libicw_b200.so; fi
  python bench.py --workload c4 --no-workloads --no-cpu --no-e2e --no-parity --steps 5 --warmup 3 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/ab_$v.json')); print('$v', d['ms_per_step'])"
done
