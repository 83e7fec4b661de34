# e2e throughput against the sub-batch size of the host-buffer calls (PHQ_SUB_BATCH_READS)
mkdir -p gpurun_out
for s in 20 21 22 23 24; do
  PHQ_SUB_BATCH_READS=$((1<<s)) python bench.py --steps 3 --warmup 3 --reads $((1<<26)) --no-cpu-baseline > gpurun_out/e2e_sweep_$s.json 2> gpurun_out/e2e_sweep_$s.err
  python - <<PY
import json
l=json.load(open('gpurun_out/e2e_sweep_$s.json'))
print($s, 'e2e %.3e  full %.3e  raw %.3e'%(l['e2e']['value'], l['e2e_full']['value'], l.get('e2e_raw',{}).get('value',0)))
PY
done
