mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu -k "not full_size" > gpurun_out/pytest10.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest10.log
for cfg in "32 8" "32 4" "32 3" "16 4" "24 4" "48 4" "64 4"; do
  set -- $cfg
  GPB_M0=$1 GPB_MUP=$2 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b_$1_$2.json 2>/dev/null
  python -c "
import json,sys; d=json.load(open('gpurun_out/b_$1_$2.json')); s=d['stages_ms']; print('M0=$1 MUP=$2  it/s %.1f  ms %.3f  solve %.3f fwd0 %.3f asm %.3f lin_other %.3f levels %d' % (d['value'], d['ms_per_step'], s['solve'], s['solve_fwd_level0'], s['assemble'], s['linearise_other'], d['config']['solver_levels']))"
done
