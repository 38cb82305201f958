python scripts/configs_2_5.py > gpurun_out/configs_2_5_v4.jsonl 2> gpurun_out/configs_2_5_v4.err || tail -5 gpurun_out/configs_2_5_v4.err
python - <<PY
import json
for l in open("gpurun_out/configs_2_5_v4.jsonl"):
    d=json.loads(l); print(d["config"], d["dofs"], "merged", d["merged"]["its"], round(d["merged"]["gdofs"],2), "std", d["standard"]["its"], round(d["standard"]["gdofs"],2), "vmult gdofs", round(d["vmult"]["gdofs"],2), "frac", round(d["vmult"]["frac"],3), "sym", d["symmetry_rel"])
PY
for g in 0 1; do
  echo "=== PROBE_GEOM=$g (deformed eps 0.1)"
  PROBE_GEOM=$g PROBE_EPS=0.1 PROBE_REPS=30 python scripts/gpu_perf_probe.py 30e6 2,3,4,5,6,7,8 1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['p'], d['dofs'], 'vmult_ms', d['vmult_ms'], 'vmult_gdofs', d['vmult_gdofs'], 'vmult_gbs', d['vmult_gbs'], 'cg_gdofs', d['cg_gdofs'])
"
done
