#!/bin/bash
# round 2: fused kernel A/B (quick): parity, then p=6 at the headline size for a few schedules
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > $O/r2_fused2_tests.log
tail -2 $O/r2_fused2_tests.log
rm -f $O/r2_fused2_probe.log
for cfg in ${CFGS:-"BP5_NO_FUSE=1" "BP5_FUSE_S=1" "BP5_FUSE_S=2" "BP5_FUSE_S=4" "BP5_FUSE_S=2,BP5_FUSE_UA=3,BP5_FUSE_DL=3"}; do
  echo "== $cfg" >> $O/r2_fused2_probe.log
  env $(echo $cfg | tr ',' ' ') PROBE_REPS=10 timeout 300 python scripts/gpu_perf_probe.py ${SIZE:-148e6} ${DEG:-6} ${QUAD:-1} >> $O/r2_fused2_probe.log 2>&1
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2_fused2_probe.log'):
    ln=ln.strip()
    if ln.startswith('{'):
        d=json.loads(ln); print(d['p'],d['quad'],d['dofs'],'vmult',d['vmult_ms'],d['vmult_frac'],'cell',d['cellloop_ms'],'cg',d['cg_ms_per_it'],d['cg_gdofs'],d['cg_frac'])
    else: print(ln[:300])
PY
