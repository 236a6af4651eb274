#!/bin/bash
# One 8-GPU box: the multi-device parity test, bench.py at 1/2/4/8 ranks (strong scaling on C4), C4 / C5 through the array pipeline on 1 and 8 GPUs.
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "all_devices or parts_union" > gpurun_out/r2p_pytest_8gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest_8gpu.log
bash tools/scale_bench.sh
[ -n "$SKIP_CONFIGS" ] && exit 0
CONFIGS=C4,C5 DEVICE_SETS=1,8 SYNTH_WORKERS=32 timeout 600 python tools/run_configs.py > gpurun_out/r2p_configs_1_8gpu.jsonl 2> gpurun_out/r2p_configs_1_8gpu.err; echo "configs rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/r2p_configs_1_8gpu.jsonl"):
    d=json.loads(ln); print(d["config"], d["n_gpus"], d["seconds"], d["stages_s"], d["identical_to_first_device_set"])
PY
