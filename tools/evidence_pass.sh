#!/bin/bash
# One GPU call that refreshes the round's evidence under gpurun_out/ (copied to profiles/ by hand): usage evidence_pass.sh TAG
T=${1:-rXX}
R=$(python -c "from parallel_dmd_for_biomolecules_b200.dmd import device_fill; print(device_fill(0)[0])")
mkdir -p gpurun_out
python bench.py > gpurun_out/${T}_bench_ours.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_raw.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dmd_event_loop_kernel -s 1 -c 1 -o gpurun_out/${T}_evl -f python tools/prof_run.py $R 20000 2000 > gpurun_out/${T}_ncu_full.log 2>&1
DMDB_DEBUG=1 DMDB_LIB=build/lib_prof.so python tools/prof_run.py $R 20000 20000 > gpurun_out/${T}_phase.log 2>&1
(time python -m pytest tests -x -q -m gpu) > gpurun_out/${T}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${T}_pytest_gpu.log; cut -c1-700 gpurun_out/${T}_bench_ours.json; tail -20 gpurun_out/${T}_phase.log
