#!/bin/bash
# round 2: ncu evidence -- one full capture of the cell kernel per degree (and Gauss, and geometry on the fly),
# and the launch list of the bench command.  Every command runs once without ncu first (&&).
cd "$(dirname "$0")/.."
O=gpurun_out
cap() {   # name, args...
  name=$1; shift
  python scripts/ncu_target.py "$@" > $O/ncu_plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:bp5_apply -s 2 -c 1 -o $O/r2_$name -f \
      python scripts/ncu_target.py "$@" > $O/ncu_full_$name.log 2>&1
  echo "$name rc=$?"
  # gpurun brings back at most 64 MiB: keep the raw page as CSV (and the source page of the headline degree)
  ncu -i $O/r2_$name.ncu-rep --page raw --csv > $O/r2_${name}_raw.csv 2>/dev/null
  if [ "$name" = "apply_p6_gll" ]; then ncu -i $O/r2_$name.ncu-rep --page source --csv > $O/r2_${name}_source.csv 2>/dev/null; fi
  rm -f $O/r2_$name.ncu-rep
}
cap apply_p4_gll 4 gll 96 4
cap apply_p5_gll 5 gll 77 4
cap apply_p6_gll 6 gll 64 4
cap apply_p7_gll 7 gll 55 4
cap apply_p8_gll 8 gll 48 4
cap apply_p6_gauss 6 gauss 64 4
cap apply_otf_p5_gll 5 gll 60 4 otf 0.1
python bench.py --steps 1 --warmup 1 --no-variants --no-cpu-baseline > $O/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file $O/r2_launches_bench_p6_gll.csv \
   python bench.py --steps 1 --warmup 1 --no-variants --no-cpu-baseline > $O/ncu_launches_bench.log 2>&1
echo "launch list rc=$?"
ls -la $O/
