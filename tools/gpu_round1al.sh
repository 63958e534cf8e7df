#!/bin/bash
# round 1al: full-size property tests (1M CSTR QPs, 10M-state structured network)
set -x
timeout -k 10 600 python -m pytest tests/test_gpu_fullsize.py -x -q -s 2>&1 | tail -25
