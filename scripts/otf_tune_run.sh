#!/bin/bash
# one GPU call: parity of the on-the-fly kernels with the shipped library, then the tuning probe over the variant
# builds libbp5b200_otfg_{a,b,c,...}.so (scripts/build_variant.sh); a = shipped tile sizes.  usage: otf_tune_run.sh "a b c" 4,5,6
variants=${1:-"a b c"}; degrees=${2:-4,5,6}
[ -n "$TEST_LIB" ] && export BP5_LIB=$PWD/deal-and-ceed-on-gpu_b200/libbp5b200_otfg_$TEST_LIB.so   # parity of a variant build
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "on_the_fly" > gpurun_out/otfg_tests2.log 2>&1
echo tests rc=$?; tail -3 gpurun_out/otfg_tests2.log; unset BP5_LIB
rm -f gpurun_out/otf_tune.jsonl gpurun_out/otf_tune.err
for v in $variants; do
  l=$v; [ $v = a ] && l=default
  BP5_LIB=$PWD/deal-and-ceed-on-gpu_b200/libbp5b200_otfg_$v.so timeout 150 python scripts/otf_tune_probe.py $l $degrees >> gpurun_out/otf_tune.jsonl 2>> gpurun_out/otf_tune.err
done
python - <<PY
import json
for l in open("gpurun_out/otf_tune.jsonl"):
    d = json.loads(l)
    print(d["build"], d["p"], d["case"], d["kernel"].split("cells_per_tile=")[1].split(",")[0].rstrip(">"), d["vmult_ms"], d["vmult_gdofs"], d["cg_gdofs"])
PY
tail -3 gpurun_out/otf_tune.err
