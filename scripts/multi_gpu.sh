#!/bin/bash
# Multi-GPU session on ONE box: value-level check of the engine-owned NCCL path and bench lines at N ranks.
#   scripts/multi_gpu.sh N TAG [steps...]     steps: check c3 c3_1m c4 c5   (outputs under gpurun_out/TAG_*_nN.*)
N=$1; TAG=$2; shift 2
STEPS=${@:-"check c3"}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
for s in $STEPS; do
  case $s in
    check) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/nccl_check.py > gpurun_out/${TAG}_nccl_check_n$N.log 2>&1; echo "nccl_check rc=$?"; grep -E "PASS|FAIL|Error|error" gpurun_out/${TAG}_nccl_check_n$N.log | head -12;;
    c3|c3_fuse|c3_1m|c4|c5)
      unset GPB_FUSE_L0
      case $s in c3) ARGS="";; c3_fuse) ARGS=""; export GPB_FUSE_L0=1;; c3_1m) ARGS="--states 1000000";; c4) ARGS="--config C4";; c5) ARGS="--config C5";; esac
      timeout 900 $LAUNCH bench.py --gpus $N --steps 20 --warmup 3 --no-cpu $ARGS > gpurun_out/${TAG}_bench_${s}_n$N.json 2> gpurun_out/${TAG}_bench_${s}_n$N.err; echo "bench $s N=$N rc=$?"
      python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_${s}_n$N.json"))
    print("  %s N=%d: %.1f it/s, %.3f ms/iter, e2e %.1f it/s, all-reduces/step %s, launches %s, stages %s" % ("$s", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("allreduces_per_step"), d.get("gpu_launches"), {k: round(v, 3) for k, v in d["stages_ms"].items()}))
except Exception as e:
    print("  no bench line:", e)
PY
      tail -2 gpurun_out/${TAG}_bench_${s}_n$N.err;;
  esac
done
