#!/bin/bash
# round 1am: final check of the committed tree - every GPU test and the smoke entry point
set -x
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout -k 10 300 python __graft_entry__.py --smoke 2>&1 | tail -2
