#!/bin/bash
# The round's ncu evidence for the bench step (1 GPU): launch list + one --set full capture of each
# of the three kernels of the step.  Everything lands in gpurun_out/<tag>_*; summaries are made
# here with tools/ncu_summary.py / tools/launch_shares.py and committed under profiles/.
tag=${1:-r02}
CMD="python bench.py --main-only --steps 2 --warmup 3"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
cat gpurun_out/${tag}_plain.log | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_launches_bench.csv $CMD > gpurun_out/${tag}_ncu0.log 2>&1
for k in fft_kernel k_mixdecim_stream k_tuner_scout; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/${tag}_$k $CMD > gpurun_out/${tag}_ncu_$k.log 2>&1
  tail -2 gpurun_out/${tag}_ncu_$k.log
done
ls -la gpurun_out/${tag}_*
