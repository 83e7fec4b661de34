#!/bin/bash
# round 2, call P: thread-per-read tie kernel — parity tests (short set), quick bench, launch lists, per-line captures of the
# tie kernel (c1, c3) and of the small-codec prefilter launch of c3 (4 barcodes: the per-read overhead on its own)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_whitelist_depth.py ) > gpurun_out/r2p_tests.log 2>&1
tail -5 gpurun_out/r2p_tests.log
( time timeout 900 python bench.py --no-cpu-baseline --no-e2e ) > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
tail -c 600 gpurun_out/r2p_bench.err
python scripts/bench_table.py gpurun_out/r2p_bench.json | cut -c1-200
for w in c1 c3 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pamld|mdd|count_kernel' -c 40 --csv --log-file gpurun_out/r2p_launches_$w.csv \
      python bench.py --workload $w --reads 16777216 --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2p_list_$w.log 2>&1
  python - $w <<'PY'
import csv, sys
rows = [r for r in csv.reader(open("gpurun_out/r2p_launches_%s.csv" % sys.argv[1])) if len(r) > 5]
head = next(r for r in rows if "Kernel Name" in r)
k, v = head.index("Kernel Name"), head.index("Metric Value")
seen = {}
for r in rows[rows.index(head) + 1:]:
    seen.setdefault(r[k][:70], []).append(float(r[v].replace(",", "")) / 1000)
for name, times in seen.items():
    print("   %-70s n=%2d median %9.1f us  all %s" % (name, len(times), sorted(times)[len(times) // 2], " ".join("%.0f" % t for t in times[:8])))
PY
done
cap() { # tag workload kernel-regex skip count reads lines
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 -o /tmp/r2p_full_$1 -f \
      python bench.py --workload $2 --reads $6 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2p_full_$1.log 2>&1
  ncu -i /tmp/r2p_full_$1.ncu-rep --page raw --csv > gpurun_out/r2p_full_$1_raw.csv 2>/dev/null
  for k in $(seq 1 $5); do python scripts/ncu_lines.py /tmp/r2p_full_$1.ncu-rep $k $7 > gpurun_out/r2p_full_$1_lines_$k.txt 2>&1; done
}
# c3 launches per step: (fast, exact, tie) x 4 decoders = 12; the 4th decoder (4 barcodes) of the 4th step is launches 46-48
cap c3small c3 'pamld_fast_kernel|pamld_kernel|pamld_tie' 45 3 16777216 160
cap c3tie c3 'pamld_tie' 12 1 16777216 120
cap c1tie c1 'pamld_tie' 3 1 16777216 120
du -sh gpurun_out; ls gpurun_out | grep r2p | head -40
