#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -15
