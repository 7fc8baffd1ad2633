#!/bin/bash
# One GPU-box call of round 2's second session: the GPU test tier with the new kernel variants as defaults, then A/B timings
# of the variants (profiles/tools/step_time.py; options through the DCT3D_* environment defaults), then the parity file of the
# test tier once more on the plain paths.  Everything is written under gpurun_out/.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_tests.log
for v in "0 0 0" "1 0 0" "0 1 0" "0 1 1" "1 1 1"; do
  set -- $v
  DCT3D_ZERO_SKIP=$1 DCT3D_TMA_STORE=$2 DCT3D_COL_CLASSES=$3 timeout 120 python profiles/tools/step_time.py 256 20 >> gpurun_out/c1_ab.jsonl 2>> gpurun_out/c1_ab.err
done
if [ -f build/libdct3d_pad0.so ]; then
  DCT3D_LIB=$GRAFT_REPO_ROOT/build/libdct3d_pad0.so timeout 120 python profiles/tools/step_time.py 256 20 >> gpurun_out/c1_ab.jsonl 2>> gpurun_out/c1_ab.err
  DCT3D_LIB=$GRAFT_REPO_ROOT/build/libdct3d_pad0.so DCT3D_TMA_STORE=0 DCT3D_ZERO_SKIP=0 timeout 120 python profiles/tools/step_time.py 256 20 >> gpurun_out/c1_ab.jsonl 2>> gpurun_out/c1_ab.err
fi
timeout 120 python profiles/tools/step_time.py 64 10 8 noise >> gpurun_out/c1_ab.jsonl 2>> gpurun_out/c1_ab.err
DCT3D_ZERO_SKIP=0 DCT3D_TMA_STORE=0 timeout 120 python profiles/tools/step_time.py 64 10 8 noise >> gpurun_out/c1_ab.jsonl 2>> gpurun_out/c1_ab.err
( DCT3D_ZERO_SKIP=0 DCT3D_TMA_STORE=0 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/c1_tests_old.log 2>&1
tail -3 gpurun_out/c1_tests.log; cat gpurun_out/c1_ab.jsonl; tail -3 gpurun_out/c1_tests_old.log
