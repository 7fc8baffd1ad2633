#!/bin/bash
# The driver's round-end sequence on the final tree: GPU test tier with the defaults, then smoke().
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_tests.log
tail -4 gpurun_out/final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
