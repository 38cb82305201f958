#!/bin/bash
# round 2: first run of the fused per-iteration kernel -- parity, then A/B against the separate kernels
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15 > $O/r2_fused1_tests.log
tail -3 $O/r2_fused1_tests.log
for cfg in "BP5_NO_FUSE=1" "BP5_FUSE_S=1" "BP5_FUSE_S=2" "BP5_FUSE_S=4" "BP5_FUSE_S=8"; do
  echo "== $cfg" >> $O/r2_fused1_probe.log
  env $cfg PROBE_REPS=10 timeout 300 python scripts/gpu_perf_probe.py 148e6 6 1 >> $O/r2_fused1_probe.log 2>&1
done
for cfg in "BP5_NO_FUSE=1" "BP5_FUSE_S=1" "BP5_FUSE_S=4"; do
  echo "== $cfg" >> $O/r2_fused1_probe.log
  env $cfg PROBE_REPS=10 timeout 300 python scripts/gpu_perf_probe.py 57e6 4,5,7,8 1 >> $O/r2_fused1_probe.log 2>&1
  env $cfg PROBE_REPS=10 timeout 300 python scripts/gpu_perf_probe.py 57e6 6 0 >> $O/r2_fused1_probe.log 2>&1
done
cat $O/r2_fused1_probe.log | cut -c1-400
