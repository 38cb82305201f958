# 1-GPU job: tests, bench (headline), launch list of the bench command, full ncu capture of the in-CG cell kernel
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_n1_v4.json 2> gpurun_out/bench_r1_n1_v4.err || tail -20 gpurun_out/bench_r1_n1_v4.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err || tail -20 gpurun_out/bench_r1_ref.err
# launch list of the same command (share of the step); bench already exited 0 without ncu above
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/r1_launches_v4.csv \
   python bench.py --steps 1 --warmup 1 --no-variants --no-cpu-baseline > gpurun_out/ncu_bench_v4.log 2>&1
# full capture of the cell kernel as the CG runs it (fused dot product) at the headline size
ncu --set full --clock-control none --import-source on -k regex:bp5_apply_kernel -s 20 -c 2 -o gpurun_out/r1_apply_p6_gll_v4 -f \
   python bench.py --steps 1 --warmup 1 --no-variants --no-cpu-baseline > gpurun_out/ncu_full_v4.log 2>&1
ls -la gpurun_out/*.ncu-rep
