N=$1
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 2 "$@"; }
run --transport peer --scaling weak   > gpurun_out/multi_${N}_peer_weak.json   2> gpurun_out/multi_${N}_peer_weak.err
run --transport peer --scaling strong > gpurun_out/multi_${N}_peer_strong.json 2> gpurun_out/multi_${N}_peer_strong.err
run --transport peer --scaling strong --cells 44 > gpurun_out/multi_${N}_peer_strong44.json 2> gpurun_out/multi_${N}_peer_strong44.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/multi_${N}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 1), "roof", round(d["roofline"]["frac"], 3), d["config"]["dofs_global"])
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
