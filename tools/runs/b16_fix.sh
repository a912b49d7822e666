#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_b16.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -25
timeout 900 compute-sanitizer --tool racecheck python tests/tools/sanitize_small.py b16 2>&1 | grep -E "RACECHECK SUMMARY|Race|b16|done" | head
