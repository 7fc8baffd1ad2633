#!/bin/bash
# The evidence call of a build: the default bench line, the ncu launch list of the same command (short form), and one
# `ncu --set full` capture of the eight kernels of an encode+decode step.  Outputs under gpurun_out/ (r2_* names).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench.json
python bench.py --steps 2 --warmup 3 --quick > gpurun_out/plain_l.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
python profiles/tools/prof_run.py 256 > gpurun_out/plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'encode_kernel|eg_pack|seg_|reconstruct_coo' -s 16 -c 8 \
    -o gpurun_out/r2_full -f python profiles/tools/prof_run.py 256 > gpurun_out/ncu_r2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -12
