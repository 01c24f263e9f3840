#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 400 python -m pytest tests/test_gpu_collectives.py -x -q -m gpu 2>&1 | tail -15
