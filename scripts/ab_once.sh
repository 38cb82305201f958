#!/bin/bash
# A/B (one pass each, shipped first) of a tuning build on the perf probe: scripts/ab_once.sh <variant> <target dofs> <degrees> <quads>
name=$1; target=$2; degrees=$3; quads=$4
for lib in shipped $name; do
  if [ $lib = shipped ]; then unset BP5_LIB; else export BP5_LIB=$PWD/deal-and-ceed-on-gpu_b200/libbp5b200_$lib.so; fi
  PROBE_REPS=${PROBE_REPS:-10} timeout 200 python scripts/gpu_perf_probe.py $target $degrees $quads 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception:
        print(l.rstrip()); continue
    print('$lib', {k: d[k] for k in ('p', 'quad', 'dofs', 'vmult_ms', 'vmult_frac', 'cellloop_ms', 'cg_ms_per_it', 'cg_gdofs', 'cg_frac')})"
done
