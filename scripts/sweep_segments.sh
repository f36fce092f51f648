mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "not full_size" > gpurun_out/pytest10.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest10.log
for cfg in ${SWEEP:-"32 8" "24 8" "40 8" "44 8" "48 8" "44 6"}; do
  set -- $cfg
  GPB_M0=$1 GPB_MUP=$2 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b_$1_$2.json 2>/dev/null
  python -c "
import json,sys; d=json.load(open('gpurun_out/b_$1_$2.json')); s=d['stages_ms']; print('M0=$1 MUP=$2  it/s %.1f  ms %.3f  e2e %.1f solve %.3f fwd0 %.3f asm %.3f lin_other %.3f levels %d clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], s['solve'], s['solve_fwd_level0'], s['assemble'], s['linearise_other'], d['config']['solver_levels'], d['clocks']))"
done
