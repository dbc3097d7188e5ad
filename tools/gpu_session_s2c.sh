#!/bin/bash
# Single-GPU check of the matrix-Hamiltonian gradients and the leaner backward-pass launch sequence
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dist.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== examples"; timeout 300 python examples/test_gradients.py 2>&1 | tail -2
echo "== gradients"; timeout 900 python tools/bench_gradients.py > gpurun_out/gradients.jsonl 2> gpurun_out/gradients.err; cut -c1-300 gpurun_out/gradients.jsonl
