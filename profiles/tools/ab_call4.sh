#!/bin/bash
# Fourth A/B call: GPU test tier of the build with the dense-cube paths and the deferred row-pointer clamp, timings on natural
# and noise content, and two small compile-time variants (prefix tile size, lists in flight in the emit kernel).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/c4_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c4_tests.log
OUT=gpurun_out/c4_ab.jsonl
run() { timeout 120 python profiles/tools/step_time.py "$@" >> $OUT 2>> gpurun_out/c4_ab.err; }
run 256 20
run 64 10 8 noise
for v in scanitems16 emit2; do
  if [ -f build/libdct3d_$v.so ]; then DCT3D_LIB=$GRAFT_REPO_ROOT/build/libdct3d_$v.so run 256 20; fi
done
run 256 20
tail -4 gpurun_out/c4_tests.log
cat $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['lib'], d['kind'], d['cube'], d['opts'], 'step %.4f enc %.4f dec %.4f | enc_k %.4f rec_k %.4f | rest_enc %.4f rest_dec %.4f' % (d['ms_per_step'], d['encode_ms'], d['decode_ms'], d['encode_kernel_ms'], d['reconstruct_kernel_ms'], d['encode_ms'] - d['encode_kernel_ms'], d['decode_ms'] - d['reconstruct_kernel_ms']), d['stream_sha'], d['frames_sha'])
"
