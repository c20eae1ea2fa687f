#!/bin/bash
# One gpurun call of the round-2 development loop: GPU tests, then the bench, logs under gpurun_out/.
# usage: scripts/gpu_round.sh TAG [pytest-args...]
TAG=${1:-x}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
