#!/bin/bash
# One B200: the whole GPU test suite, the driver's bench command, the reference arm, smoke, and the ncu evidence of the same build.
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/r2z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2z_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench_c4.json 2> gpurun_out/r2z_bench_c4.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_bench_ref.err; echo "ref rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline"
$B > gpurun_out/r2z_plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2z_launches.csv $B > gpurun_out/r2z_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:join_kernel -s 30 -c 4 -o gpurun_out/r2z_join $B > gpurun_out/r2z_ncu2.log 2>&1; echo "ncu full rc=$?"
