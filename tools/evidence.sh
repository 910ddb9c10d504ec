#!/bin/bash
# dev helper (GPU box): refresh the round's evidence into gpurun_out/
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 48 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --quick --no-cpu > gpurun_out/launches.log 2>&1
( echo "## memcheck"; timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_a1_gpu.py -x -q -k "golden or parity or height" 2>&1 | tail -5; echo "MEMCHECK_EXIT=$?";
  echo "## racecheck"; timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_a1_gpu.py -x -q -k "golden" 2>&1 | tail -5 ) > gpurun_out/sanitizer.txt 2>&1
tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | cut -c1-600; cat gpurun_out/sanitizer.txt
