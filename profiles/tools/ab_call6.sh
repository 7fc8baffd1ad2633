#!/bin/bash
# Sixth call: the GPU test tier with the balanced packer forced through the environment default, then timings of the sorted
# (1) and balanced (2) packers on 8^3 and 4^3 cubes.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time DCT3D_PACK_SORT=2 timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/c6_tests.log 2>&1
rc=$?
echo "pytest rc=$rc" >> gpurun_out/c6_tests.log
tail -4 gpurun_out/c6_tests.log
OUT=gpurun_out/c6_ab.jsonl
run() { timeout 120 python profiles/tools/step_time.py "$@" >> $OUT 2>> gpurun_out/c6_ab.err; }
DCT3D_PACK_SORT=2 run 256 20
DCT3D_PACK_SORT=1 run 256 20
DCT3D_PACK_SORT=2 run 256 10 4
DCT3D_PACK_SORT=1 run 256 10 4
DCT3D_PACK_SORT=2 run 64 10 8 noise
cat $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['kind'], d['cube'], d['opts'], 'step %.4f enc %.4f dec %.4f | enc_k %.4f rec_k %.4f | rest_enc %.4f rest_dec %.4f' % (d['ms_per_step'], d['encode_ms'], d['decode_ms'], d['encode_kernel_ms'], d['reconstruct_kernel_ms'], d['encode_ms'] - d['encode_kernel_ms'], d['decode_ms'] - d['reconstruct_kernel_ms']), d['stream_sha'], d['frames_sha'])
"
