#!/bin/bash
# final build: one full ncu capture of the IN-LOOP cell kernel (merged CG, fused dot product) at the bench size, and the
# launch list of the bench command.  Every command runs once without ncu first (&&).
cd "$(dirname "$0")/.."
O=gpurun_out
NCU_TARGET_CG=1 python scripts/ncu_target.py 6 gll 88 6 > $O/ncu_plain_cg.log 2>&1 &&
NCU_TARGET_CG=1 ncu --set full --clock-control none --import-source on -k regex:bp5_apply -s 3 -c 1 -o $O/r2_final_cg_p6_gll_88 -f \
    python scripts/ncu_target.py 6 gll 88 6 > $O/ncu_full_cg.log 2>&1
echo "cg capture rc=$?"
ncu -i $O/r2_final_cg_p6_gll_88.ncu-rep --page raw --csv > $O/r2_final_cg_p6_gll_88_raw.csv 2>/dev/null
rm -f $O/r2_final_cg_p6_gll_88.ncu-rep
python bench.py --steps 1 --warmup 1 --no-variants --no-cpu-baseline > $O/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file $O/r2_final_launches_bench_p6_gll.csv \
   python bench.py --steps 1 --warmup 1 --no-variants --no-cpu-baseline > $O/ncu_launches_bench.log 2>&1
echo "launch list rc=$?"
ls -la $O/ | tail -8
