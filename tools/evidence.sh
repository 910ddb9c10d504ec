#!/bin/bash
# dev helper (GPU box): refresh the round's evidence into gpurun_out/ (copied to profiles/ by hand)
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_cpu.json 2> gpurun_out/r2_bench_ref.err
( echo "## memcheck: fused A1 path (pipelined + phased kernels, reset, eval_terms)"; timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_a1_gpu.py -x -q -k "golden or parity_random or height or instantiations or exact_division" 2>&1 | tail -4;
  echo "## memcheck: ABB post-physics / reset_idx, arm IK (N2), camera gather (N4), terrain generator (N3)"; timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_abb_gpu.py tests/test_camera_gpu.py tests/test_terrain_gpu.py -x -q 2>&1 | tail -4;
  echo "## racecheck: fused A1 path (shared-memory pipeline, 3-stage ring, split B group)"; timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_a1_gpu.py -x -q -k "golden" 2>&1 | tail -4;
  echo "## memcheck: HAS_EXTRA instantiation (legged_gym-style terms incl. feet_air_time)"; timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_dropin_gpu.py -x -q -k "stateful or edited" 2>&1 | tail -4;
  echo "## racecheck: HAS_EXTRA instantiation, asymmetric-grid fallback, curriculum off"; timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_dropin_gpu.py tests/test_a1_gpu.py -x -q -k "stateful or asymmetric or curriculum" 2>&1 | tail -4;
  echo "## racecheck: camera gather (shared-memory table)"; timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_camera_gpu.py -x -q -k "fixture" 2>&1 | tail -4 ) > gpurun_out/r2_compute_sanitizer.txt 2>&1
tail -3 gpurun_out/r2_bench_n1.err; cut -c1-400 gpurun_out/r2_bench_n1.json; cat gpurun_out/r2_compute_sanitizer.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pd_torque|a1_post|compact_ids|collect_stats|publish_extras|a1_reset|body_frame" -s 18 -c 24 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 3 --quick --no-cpu > gpurun_out/r2_launches.log 2>&1
timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:a1_post_physics_tma -s 4 -c 1 -o gpurun_out/prof_r2_final -f python tools/ncu_run.py > gpurun_out/ncu_r2_final.log 2>&1
tail -2 gpurun_out/ncu_r2_final.log
