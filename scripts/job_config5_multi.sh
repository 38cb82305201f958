# BASELINE config 5 on N GPUs: BP5 p=5, smoothly deformed mesh, 60^3 cells per GPU, stored metric vs on-the-fly geometry
N=$1
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 2 --warmup 1 --degree 5 --cells 60 --deformation 0.1 "$@"; }
run --geometry stored > gpurun_out/config5_${N}_stored.json 2> gpurun_out/config5_${N}_stored.err
run --geometry otf    > gpurun_out/config5_${N}_otf.json    2> gpurun_out/config5_${N}_otf.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/config5_${N}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms/step", round(d["ms_per_step"], 1), "roof", round(d["roofline"]["frac"], 3), d["config"]["dofs_global"], d["config"]["kernel"], "x_l2", d["check"]["x_l2"])
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
