# final 1-GPU measurement set of the round
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref_final.json 2> gpurun_out/bench_r1_ref_final.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_n1_final.json 2> gpurun_out/bench_r1_n1_final.err || tail -20 gpurun_out/bench_r1_n1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/r1_launches_final.csv \
   python bench.py --steps 1 --warmup 1 --no-variants --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1
python scripts/configs_2_5.py > gpurun_out/configs_2_5_final.jsonl 2> gpurun_out/configs_2_5_final.err
python scripts/sweep.py gpurun_out/sweep_final.jsonl > gpurun_out/sweep_final.log 2>&1
tail -2 gpurun_out/sweep_final.log
