python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
PROBE_REPS=40 python scripts/gpu_perf_probe.py 57e6 2,3,4,5,6,7,8 0,1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['p'], d['quad'], 'cellloop', round(d['cellloop_gbs']/6548.2,3), 'vmult_frac', d['vmult_frac'], 'cg', d['cg_gdofs'], 'cg_frac', d['cg_frac'])
"
