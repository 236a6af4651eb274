for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2p_scale_g$n.json 2> gpurun_out/r2p_scale_g$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2p_scale_g$n.json 2> gpurun_out/r2p_scale_g$n.err
  fi
  echo "n=$n rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2p_scale_g$n.json").read().strip().splitlines()[-1])
    print($n, "ms", d["ms_per_step"], "value %.3e"%d["value"], "e2e ms", d["e2e"]["ms_per_step"], "edges", d["details"]["edges"], "rank0 edges", d["details"]["edges_rank0"], "frac", d["roofline"]["frac"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2p_scale_g$n.err").read()[-1500:])
PY
done
