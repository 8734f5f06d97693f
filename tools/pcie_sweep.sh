#!/bin/bash
# tools/pcie_probe.py at 1/2/4/8 ranks on one box: the host<->device copy ceiling that bounds
# bench.py's e2e leg.  Output is committed as profiles/r02_pcie_probe.txt.
out=${1:-gpurun_out/pcie_probe.txt}
: > $out
nvidia-smi topo -m >> $out 2>&1
echo "host: $(nproc) cpus, $(grep MemTotal /proc/meminfo)" >> $out
lscpu | grep -E "NUMA|Model name|Socket" >> $out 2>&1
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
python tools/pcie_probe.py >> $out 2>&1
for n in 2 4 8; do $T --nproc-per-node $n tools/pcie_probe.py 2>&1 | grep '^{' >> $out; done
$T --nproc-per-node 8 tools/pcie_probe.py --affinity 2>&1 | grep '^{' >> $out
$T --nproc-per-node 8 tools/pcie_probe.py --mode up 2>&1 | grep '^{' >> $out
$T --nproc-per-node 8 tools/pcie_probe.py --mode down 2>&1 | grep '^{' >> $out
cat $out
