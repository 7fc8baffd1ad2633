#!/bin/bash
# Third A/B call: GPU test tier with the defaults (zero_skip, tma_store with the swizzled tile, pack_sort), then timings with
# the sorted packer on and off, on natural and noise content and for 4^3 cubes.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/c3_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c3_tests.log
OUT=gpurun_out/c3_ab.jsonl
run() { timeout 120 python profiles/tools/step_time.py "$@" >> $OUT 2>> gpurun_out/c3_ab.err; }
DCT3D_PACK_SORT=1 run 256 20
DCT3D_PACK_SORT=0 run 256 20
DCT3D_PACK_SORT=1 run 256 20
DCT3D_PACK_SORT=0 run 256 20
DCT3D_PACK_SORT=1 run 64 10 8 noise
DCT3D_PACK_SORT=0 run 64 10 8 noise
DCT3D_PACK_SORT=1 run 256 10 4
DCT3D_PACK_SORT=0 DCT3D_TMA_STORE=0 run 256 10 4
tail -4 gpurun_out/c3_tests.log
cat $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['kind'], d['cube'], d['opts'], 'step %.4f enc %.4f dec %.4f | enc_k %.4f rec_k %.4f | rest_enc %.4f rest_dec %.4f' % (d['ms_per_step'], d['encode_ms'], d['decode_ms'], d['encode_kernel_ms'], d['reconstruct_kernel_ms'], d['encode_ms'] - d['encode_kernel_ms'], d['decode_ms'] - d['reconstruct_kernel_ms']), d['stream_sha'], d['frames_sha'])
"
