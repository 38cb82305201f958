for v in ${VARIANTS:-base A B C}; do
  if [ $v = base ]; then unset BP5_LIB; else export BP5_LIB=$PWD/build/tune/$v/libbp5b200.so; fi
  echo "=== variant $v"
  PROBE_REPS=40 python scripts/gpu_perf_probe.py 57e6 ${DEGREES:-4,5,6,8} ${QUADS:-1} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['p'], d['kernel'], 'cellloop_ms', d['cellloop_ms'], 'frac', round(d['cellloop_gbs']/6548.2,3), 'vmult_frac', d['vmult_frac'], 'cg', d['cg_gdofs'], 'cg_frac', d['cg_frac'])
"
done
