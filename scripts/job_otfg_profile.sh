#!/bin/bash
# 1-GPU job: parity of the on-the-fly kernels (shipped library), the config 2 / 5 probe, and one full ncu capture of the
# general on-the-fly kernel on BASELINE config 5 (p = 5, QGauss(6), 60^3 deformed cells)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "on_the_fly" > gpurun_out/otfg_tests3.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/otfg_tests3.log
timeout 150 python scripts/otf_general_probe.py > gpurun_out/otfg_probe2.jsonl 2> gpurun_out/otfg_probe2.err; echo "probe rc=$?"
python - <<PY
import json
for l in open("gpurun_out/otfg_probe2.jsonl"):
    d = json.loads(l)
    print(d["config"], "stored", round(d["stored"]["cg_gdofs"], 2), round(d["stored"]["vmult_ms"], 3), "otf", round(d["on_the_fly"]["cg_gdofs"], 2),
          round(d["on_the_fly"]["vmult_ms"], 3), "ratio", round(d["otf_over_stored_cg"], 3), "its", d["stored"]["its"], d["on_the_fly"]["its"], "xdiff", d["x_rel_diff"])
PY
# the probe above exited without ncu; now one capture of the kernel
timeout 200 ncu --set full --clock-control none --import-source on -k regex:bp5_apply_otfg -s 1 -c 1 -o gpurun_out/r2_apply_otfg_p5_gauss -f \
   python scripts/ncu_target.py 5 gauss 60 3 otf 0.1 > gpurun_out/ncu_otfg.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_otfg.log
ls -la gpurun_out/*.ncu-rep
