#!/bin/bash
# Sweep of the phase scout's shape (CTA size x chains per thread) on the pump step of
# BASELINE config 5: step time, per-kernel event times (scout, FFT, tuner + decimator).
# Output is committed as profiles/r02_scout_sweep.txt.
out=${1:-gpurun_out/scout_sweep.txt}
: > $out
for cpt in 1 2 3 4; do
  for t in 64 128 256; do
    sms=$(( (4096 + cpt*t - 1) / (cpt*t) ))
    echo -n "cpt=$cpt threads=$t SMs_held=$sms  " >> $out
    JSDR_SCOUT_CPT=$cpt JSDR_SCOUT_THREADS=$t python tools/kbench.py pump 2>&1 | tail -1 >> $out
  done
done
cat $out
