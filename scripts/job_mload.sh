set -x
for m in 0 1 2; do
  echo "=== BP5_MLOAD=$m"
  BP5_MLOAD=$m PROBE_REPS=60 python scripts/gpu_perf_probe.py 57e6 4,5,6,7,8 1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['p'], d['kernel'], 'cellloop_ms', d['cellloop_ms'], 'frac', round(d['cellloop_gbs']/6548.2,3), 'vmult_frac', d['vmult_frac'], 'cg', d['cg_gdofs'])
"
done
