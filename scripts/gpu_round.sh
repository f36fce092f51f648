#!/bin/bash
# one GPU session: parity tests, bench, A/B of the solver switches, compute-sanitizer.  usage: scripts/gpu_round.sh <tag> [steps...]
TAG=${1:-r2a}; shift
STEPS=${@:-"tests bench ab san"}
mkdir -p gpurun_out
for s in $STEPS; do
  case $s in
    tests) timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log;;
    tests_fast) timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log;;
    bench) timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json | head -c 6000; tail -3 gpurun_out/${TAG}_bench.err;;
    benchfast) timeout 600 python bench.py --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json | head -c 6000; tail -3 gpurun_out/${TAG}_bench.err;;
    ab) eval timeout 600 python scripts/ab_variants.py ${AB_SPECS:-'"" "GPB_DENSE_PANEL=1"'} > gpurun_out/${TAG}_ab.jsonl 2> gpurun_out/${TAG}_ab.err; echo "ab rc=$?"; cat gpurun_out/${TAG}_ab.jsonl; tail -3 gpurun_out/${TAG}_ab.err;;
    lat) python -c "from gpslam_b200 import capi; import json; print(json.dumps(capi.latencies()))" | tee gpurun_out/${TAG}_latency.json;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/one_iter.py ${NCU_CFG:-C3} ${NCU_STATES:-0} 3 > gpurun_out/${TAG}_launches.log 2>&1; echo "ncu launches rc=$?"; python scripts/ncu_summaries.py ${TAG} > /dev/null 2>&1; cat profiles/${TAG}_launch_shares.txt | head -30; cp profiles/${TAG}_launch_shares.txt gpurun_out/;;
    full) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-k_level_ws|k_panel0|k_spine|k_bwd2}" -c ${NCU_C:-12} -o gpurun_out/${TAG}_full -f python scripts/one_iter.py ${NCU_CFG:-C3} ${NCU_STATES:-0} 1 > gpurun_out/${TAG}_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/${TAG}_full.log; ls -la gpurun_out/${TAG}_full.ncu-rep;;
    c4) timeout 600 python scripts/run_config.py --config C4 --steps 10 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err; echo "c4 rc=$?"; cat gpurun_out/${TAG}_c4.json; tail -3 gpurun_out/${TAG}_c4.err;;
    c5) timeout 900 python scripts/run_config.py --config C5 --steps 10 > gpurun_out/${TAG}_c5.json 2> gpurun_out/${TAG}_c5.err; echo "c5 rc=$?"; cat gpurun_out/${TAG}_c5.json; tail -3 gpurun_out/${TAG}_c5.err;;
    c4tests) timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "c4" > gpurun_out/${TAG}_c4tests.log 2>&1; echo "c4tests rc=$?"; tail -5 gpurun_out/${TAG}_c4tests.log;;
    san) bash scripts/sanitize.sh ${TAG};;
    race) CS=/usr/local/cuda/bin/compute-sanitizer
          timeout 420 $CS --tool racecheck --racecheck-report all --error-exitcode 9 python scripts/sanitize_cases.py pose3_wide 1 > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck pose3_wide rc=$?"
          timeout 420 $CS --tool racecheck --racecheck-report all --error-exitcode 9 python scripts/sanitize_cases.py sharded 1 > gpurun_out/${TAG}_racecheck_sharded.log 2>&1; echo "racecheck sharded rc=$?"
          timeout 420 $CS --tool memcheck --leak-check no --error-exitcode 9 python scripts/sanitize_cases.py all 1 > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
          grep -E "SUMMARY|sanitize_cases ok" gpurun_out/${TAG}_racecheck.log gpurun_out/${TAG}_racecheck_sharded.log gpurun_out/${TAG}_memcheck.log;;
  esac
done
