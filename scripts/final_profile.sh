#!/bin/bash
# One GPU call: the evidence committed under profiles/ for the state of the tree (tests, bench line, launch list, full
# ncu capture of the stress-frame chain with the DMMA-pipe counters).  Usage: bash scripts/final_profile.sh r2b
tag=${1:-r2b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${tag}_tests.log
tail -2 gpurun_out/${tag}_tests.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -c 600 gpurun_out/${tag}_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-mc --no-cpu-baseline > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_summary.txt
DM=smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__pipe_tensor_subpipe_dmma_cycles_active.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum
ncu --set full --metrics $DM --clock-control none --import-source on --launch-skip 16 -c 8 -o gpurun_out/${tag}_full \
    python scripts/ncu_frame.py 4096 4 > gpurun_out/${tag}_ncu_full.log 2>&1
python scripts/ncu_summary.py gpurun_out/${tag}_full.ncu-rep > gpurun_out/${tag}_ncu_full_summary.csv
cat gpurun_out/${tag}_launch_summary.txt
cut -c1-200 gpurun_out/${tag}_ncu_full_summary.csv
