#!/bin/bash
# Fifth call: GPU test tier, one timing line, then the evidence of the build (profiles/tools/evidence_call.sh).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/c5_tests.log 2>&1
rc=$?
echo "pytest rc=$rc" >> gpurun_out/c5_tests.log
tail -4 gpurun_out/c5_tests.log
[ $rc -eq 0 ] || exit 1
timeout 120 python profiles/tools/step_time.py 256 20 > gpurun_out/c5_ab.jsonl 2> gpurun_out/c5_ab.err
timeout 120 python profiles/tools/step_time.py 256 10 4 >> gpurun_out/c5_ab.jsonl 2>> gpurun_out/c5_ab.err
cat gpurun_out/c5_ab.jsonl
bash profiles/tools/evidence_call.sh
