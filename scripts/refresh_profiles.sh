#!/bin/bash
# Round-end measurement refresh (run on the GPU box through gpurun): un-profiled bench lines, ncu captures, launch list, ops sweep.
TAG=${1:-r01h}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in cartpole quadrotor satellite sweep; do
  timeout 300 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
for w in cartpole quadrotor satellite; do
  timeout 200 $NCU -k regex:knot_kernel -s 5 -o gpurun_out/prof_${w}_$TAG python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_$w.log 2>&1
done
timeout 200 $NCU -k regex:knot_kernel -s 4 -o gpurun_out/prof_quaderr_$TAG python scripts/prof_extra.py err > gpurun_out/b_quaderr.log 2>&1
timeout 200 $NCU -k regex:implicit_midpoint_block -s 2 -o gpurun_out/prof_implicit_$TAG python scripts/prof_extra.py implicit > gpurun_out/b_implicit.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 python scripts/gpu_quick.py > gpurun_out/quick_$TAG.log 2>&1
timeout 120 python scripts/implicit_bench.py >> gpurun_out/quick_$TAG.log 2>&1
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cat gpurun_out/bench_cartpole.json
