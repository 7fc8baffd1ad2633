#!/bin/bash
# Second A/B call: the two-class tail against the plain TMA tail, compile-time variants built under build/ (pack tile size,
# prefetch depth, CTAs of the encoder, run loop of the scan), then one `ncu --set full` capture of the two transform kernels.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
OUT=gpurun_out/c2_ab.jsonl
run() { timeout 120 python profiles/tools/step_time.py 256 20 >> $OUT 2>> gpurun_out/c2_ab.err; }
DCT3D_COL_CLASSES=1 run
DCT3D_COL_CLASSES=0 run
for v in cls0 pre2 pre4 pack128 pack512 enc4 runloop; do
  if [ -f build/libdct3d_$v.so ]; then DCT3D_LIB=$GRAFT_REPO_ROOT/build/libdct3d_$v.so run; fi
done
DCT3D_COL_CLASSES=1 run
DCT3D_COL_CLASSES=0 run
cat $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['lib'], d['opts'], 'step %.4f enc %.4f dec %.4f | enc_k %.4f rec_k %.4f | rest_enc %.4f rest_dec %.4f' % (d['ms_per_step'], d['encode_ms'], d['decode_ms'], d['encode_kernel_ms'], d['reconstruct_kernel_ms'], d['encode_ms'] - d['encode_kernel_ms'], d['decode_ms'] - d['reconstruct_kernel_ms']), d['stream_sha'], d['frames_sha'])
"
# ncu: the two transform kernels of the default build (classes on and off)
DCT3D_COL_CLASSES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'reconstruct_coo|encode_kernel' -s 4 -c 2 \
    -o gpurun_out/s2_full_tma python profiles/tools/prof_run.py 256 > gpurun_out/c2_ncu1.log 2>&1
DCT3D_COL_CLASSES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'reconstruct_coo' -s 2 -c 1 \
    -o gpurun_out/s2_full_cls python profiles/tools/prof_run.py 256 > gpurun_out/c2_ncu2.log 2>&1
tail -2 gpurun_out/c2_ncu1.log gpurun_out/c2_ncu2.log
ls -la gpurun_out/
