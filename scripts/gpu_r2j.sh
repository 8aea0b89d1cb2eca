#!/bin/bash
tag=${1:-r2j}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_general_action.py -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; tail -25 gpurun_out/pytest_$tag.log
