set -x
python bench.py --steps 1 --warmup 1 --no-variants --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
ncu --set full --clock-control none --import-source on -k regex:bp5_apply_kernel -s 20 -c 1 -o gpurun_out/r1_apply_p6_gll_v5 -f \
   python bench.py --steps 1 --warmup 1 --no-variants --no-cpu-baseline > gpurun_out/ncu_full_v5.log 2>&1
ls -la gpurun_out/r1_apply_p6_gll_v5.ncu-rep
