#!/bin/bash
# one gpurun call: microbench, variant timings, phase timers, parity of candidate variants
mkdir -p gpurun_out
{
echo "== pipes"; tools/ubench/pipes
echo "== variants"; python tools/variants.py run 30
for f in tools/_prof/*.so; do echo "== phases $f"; SHIFU_B200_LIB=$f python tools/prof_phases.py; done
echo "== parity newpipe (in-tree)"; python -m pytest tests/test_a1_gpu.py -m gpu -x -q 2>&1 | tail -5
echo "== parity rz"; SHIFU_B200_LIB=tools/_variants/lib_rz.so python -m pytest tests/test_a1_gpu.py -m gpu -x -q 2>&1 | tail -5
} > gpurun_out/round_a.log 2>&1
tail -80 gpurun_out/round_a.log
