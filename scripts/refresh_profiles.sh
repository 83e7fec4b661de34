#!/bin/bash
# gpurun_out/<tag>_* (what scripts/gpu_profiles.sh brings back from the GPU box) -> profiles/r02_*: usage: refresh_profiles.sh [tag]
tag=${1:-profiles}
cd "$(dirname "$0")/.."
for w in c1 c2 c3 c4 c5; do cp gpurun_out/${tag}_launches_$w.csv profiles/r02_launches_$w.csv; done
python scripts/ncu_to_profiles.py gpurun_out/${tag}_full_c1_raw.csv c1 16777216 pamld_fast_grid_kernel r02_ncu_fast_grid_c1 > /dev/null
python scripts/ncu_to_profiles.py gpurun_out/${tag}_full_c1_raw.csv - 16777216 "pamld_grid_kernel" r02_ncu_exact_grid_c1 > /dev/null
python scripts/ncu_to_profiles.py gpurun_out/${tag}_full_c1_raw.csv - 16777216 pamld_tie_kernel r02_ncu_tie_c1 > /dev/null
python scripts/ncu_to_profiles.py gpurun_out/${tag}_full_c2_raw.csv c2 16777216 mdd_table_kernel r02_ncu_mdd_table_c2 > /dev/null
python scripts/ncu_to_profiles.py gpurun_out/${tag}_full_c3_raw.csv c3 16777216 pamld_fast_kernel r02_ncu_fast_c3 > /dev/null
python scripts/ncu_to_profiles.py gpurun_out/${tag}_full_c3_raw.csv - 16777216 "pamld_kernel" r02_ncu_exact_c3 > /dev/null
python scripts/ncu_to_profiles.py gpurun_out/${tag}_full_c3_raw.csv - 16777216 pamld_tie_kernel r02_ncu_tie_c3 > /dev/null
python scripts/ncu_to_profiles.py gpurun_out/${tag}_full_c4_raw.csv c4 16777216 "pamld_fast_kernel<5" r02_ncu_fast_c4 > /dev/null
python scripts/ncu_to_profiles.py gpurun_out/${tag}_full_c5_raw.csv c5 1136640 pamld_whitelist_kernel r02_ncu_whitelist_c5 > /dev/null
cp gpurun_out/${tag}_full_c1_lines_1.txt profiles/r02_lines_fast_grid_c1.txt
cp gpurun_out/${tag}_full_c1_lines_3.txt profiles/r02_lines_tie_c1.txt
cp gpurun_out/${tag}_full_c3_lines_1.txt profiles/r02_lines_fast_c3.txt
cp gpurun_out/${tag}_full_c3_lines_3.txt profiles/r02_lines_tie_c3.txt
python scripts/sass_evidence.py > profiles/r02_sass_evidence.txt 2>/dev/null
ls -la profiles | grep r02
