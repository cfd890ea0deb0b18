#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_reference.py -m gpu -x -q -k split 2>&1 | grep -v "^$" | tail -40
TEBGPU_SPLIT=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-single-request 2>&1 | tail -12
