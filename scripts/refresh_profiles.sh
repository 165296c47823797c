#!/bin/bash
# Round-end measurement refresh (run on the GPU box through gpurun): un-profiled bench lines, ncu captures, launch list, ops sweep.
TAG=${1:-r01h}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cartpole_s20.json 2> gpurun_out/bench_cartpole_s20.err
for w in cartpole quadrotor satellite sweep; do
  timeout 300 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
for w in cartpole quadrotor satellite; do
  timeout 200 $NCU -k regex:knot_kernel -s 5 -o gpurun_out/prof_${w}_$TAG python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/b_$w.log 2>&1
done
timeout 200 $NCU -k regex:knot_kernel -s 4 -o gpurun_out/prof_quaderr_$TAG python scripts/prof_extra.py err > gpurun_out/b_quaderr.log 2>&1
timeout 200 $NCU -k regex:knot_kernel -s 4 -o gpurun_out/prof_quadbody_$TAG python scripts/prof_extra.py body > gpurun_out/b_quadbody.log 2>&1
timeout 200 $NCU -k regex:knot_kernel -s 4 -o gpurun_out/prof_quadbody64_$TAG python scripts/prof_extra.py body64 > gpurun_out/b_quadbody64.log 2>&1
timeout 200 $NCU -k regex:implicit_midpoint_block -s 2 -o gpurun_out/prof_implicit_$TAG python scripts/prof_extra.py implicit > gpurun_out/b_implicit.log 2>&1
# the reports are large (~12 MB each with sources): keep their raw pages as csv (what scripts/ncu_summary.py reads) and only the C2 report itself
for r in gpurun_out/prof_*_$TAG.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}.csv 2>/dev/null
  case $r in *cartpole*) ;; *) rm -f $r ;; esac
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 python scripts/gpu_quick.py > gpurun_out/quick_$TAG.log 2>&1
timeout 120 python scripts/implicit_bench.py >> gpurun_out/quick_$TAG.log 2>&1
timeout 120 python scripts/soa_bench.py >> gpurun_out/quick_$TAG.log 2>&1
timeout 300 python scripts/variants_sweep.py > gpurun_out/variants_$TAG.md 2>&1
NCUM="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__warps_active.avg.per_cycle_active"
RDB_SWEEP_STEPS=1 RDB_SWEEP_WARM=1 timeout 400 ncu --metrics $NCUM --clock-control none -k regex:knot_kernel --csv --log-file gpurun_out/variants_ncu_$TAG.csv python scripts/variants_sweep.py > /dev/null 2>&1
timeout 300 python scripts/small_batch.py --out gpurun_out/small_batch_$TAG.md > /dev/null 2>&1
[ -n "$SKIP_TESTS" ] || timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cat gpurun_out/bench_cartpole_s20.json
