mkdir -p gpurun_out
for cfg in ${SWEEP:-"32 8" "40 8" "44 8" "46 8" "48 8" "54 8" "62 8" "46 6" "46 12"}; do
  set -- $cfg
  GPB_M0=$1 GPB_MUP=$2 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b_$1_$2.json 2>/dev/null
  python -c "
import json,sys; d=json.load(open('gpurun_out/b_$1_$2.json')); s=d['stages_ms']; print('M0=$1 MUP=$2  it/s %.1f  ms %.3f  e2e %.1f solve %.3f fwd0 %.3f asm %.3f lin_other %.3f levels %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], s['solve'], s['solve_fwd_level0'], s['assemble'], s['linearise_other'], d['config']['solver_levels']))"
done
