#!/bin/bash
# round 2: slab-pipelined CG iteration -- parity with the pipeline forced on tiny meshes, then A/B at the headline size
cd "$(dirname "$0")/.."
O=gpurun_out
BP5_SLAB_MIN_DOFS=0 BP5_SLAB_ROUNDS=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_manufactured.py -x -q 2>&1 | tail -5 > $O/r2_slab_tests.log
tail -3 $O/r2_slab_tests.log
rm -f $O/r2_slab_probe.log
for cfg in ${CFGS:-"BP5_NO_SLAB=1" "BP5_SLAB_ROUNDS=4" "BP5_SLAB_ROUNDS=2" "BP5_SLAB_ROUNDS=8" "BP5_SLAB_ROUNDS=4,BP5_SLAB_AHEAD=1" "BP5_SLAB_ROUNDS=4,BP5_SLAB_AHEAD=3" "BP5_SLAB_ROUNDS=16,BP5_SLAB_AHEAD=1"}; do
  echo "== $cfg" >> $O/r2_slab_probe.log
  env $(echo $cfg | tr ',' ' ') PROBE_REPS=5 timeout 300 python scripts/gpu_perf_probe.py ${SIZE:-148e6} ${DEG:-6} ${QUAD:-1} >> $O/r2_slab_probe.log 2>&1
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2_slab_probe.log'):
    ln=ln.strip()
    if ln.startswith('{'):
        d=json.loads(ln); print(d['p'],d['quad'],d['dofs'],'vmult',d['vmult_ms'],d['vmult_frac'],'cg ms/it',d['cg_ms_per_it'],'GDoF/s',d['cg_gdofs'],'cg_frac72',d['cg_frac'])
    else: print(ln[:300])
PY
