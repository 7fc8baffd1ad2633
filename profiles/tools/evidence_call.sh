#!/bin/bash
# The evidence call of a build: the default bench line, the ncu launch list of the same command in its short form
# (restricted to the library's kernels: the clip generator's torch kernels would fill the launch budget), one
# `ncu --set full` capture of the eight kernels of an encode+decode step, and the quick bench lines of configs 1, 4 and 5.
# Outputs under gpurun_out/ (r2_* names); profiles/summarize.py turns them into profiles/r2_summary.md.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:'encode_kernel|eg_pack|seg_|reconstruct|transform_kernel|stream_shift|codec_f64|zz_gather|coo_scatter|rgb_planes' -c 600 \
    --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'encode_kernel|eg_pack|seg_|reconstruct_coo' -s 16 -c 8 \
    -o gpurun_out/r2_full -f python profiles/tools/prof_run.py 256 > gpurun_out/ncu_r2.log 2>&1
echo "full capture rc=$?"
for c in c1 c4 c5; do
  timeout 300 python bench.py --config $c --quick > gpurun_out/${c}_n1.json 2> gpurun_out/${c}n1.err
  echo "$c rc=$?"
done
ls -la gpurun_out | tail -14
