#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python scripts/time_observer_call.py 2>&1 | tail -80
