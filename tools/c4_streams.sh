# C4 (exact mode) run time against the stream count: streams per CTA follow the batch (ICW_SPLIT_NS forces 28)
for k in 4096 2048 1024 512 64; do
  python bench.py --workload c4 --streams $k --no-workloads --no-cpu --no-e2e --steps 3 > gpurun_out/c4s_$k.json 2>/dev/null
  ICW_SPLIT_NS=28 python bench.py --workload c4 --streams $k --no-workloads --no-cpu --no-e2e --no-parity --steps 3 > gpurun_out/c4s_${k}_ns28.json 2>/dev/null
  python -c "
import json
a=json.load(open('gpurun_out/c4s_$k.json')); b=json.load(open('gpurun_out/c4s_${k}_ns28.json'))
print('streams $k: %.1f ms (%.0f Mframes/s, parity mismatches %s) ; forced 28 per CTA: %.1f ms' % (a['ms_per_step'], a['value'], a['parity_check']['mismatches'], b['ms_per_step']))"
done
