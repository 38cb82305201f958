for ng in 0 1; do
  if [ $ng = 1 ]; then export BP5_NO_GRAPH=1; else unset BP5_NO_GRAPH; fi
  for size in 1e6 4e6; do
    echo "=== BP5_NO_GRAPH=$ng size=$size"
    PROBE_REPS=20 python scripts/gpu_perf_probe.py $size 2,4,6,8 1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['p'], d['dofs'], 'cg_ms_per_it', d['cg_ms_per_it'], 'cg', d['cg_gdofs'], 'cg_frac', d['cg_frac'], 'vmult_frac', d['vmult_frac'])
"
  done
done
