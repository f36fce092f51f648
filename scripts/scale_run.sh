#!/bin/bash
# usage: scale_run.sh N   — runs bench.py on N GPUs under torchrun (N > 1) and prints a one-line summary
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
fi
echo "rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$N.json"))
print("N=%d it/s %.1f ms %.3f e2e %.1f stages %s allreduce/step %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 3) for k, v in d["stages_ms"].items()}, d["config"].get("allreduces_per_step")))
PY
tail -3 gpurun_out/bench_n$N.err
