PROBE_GEOM=1 PROBE_EPS=0.1 PROBE_REPS=30 python scripts/gpu_perf_probe.py 30e6 2,3,4,5,6,7,8 1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['p'], d['dofs'], 'vmult_ms', d['vmult_ms'], 'vmult_gdofs', d['vmult_gdofs'], 'vmult_gbs', d['vmult_gbs'], 'cg_gdofs', d['cg_gdofs'])
"
