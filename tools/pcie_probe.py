#!/usr/bin/env python
"""Host <-> device copy ceiling of the box, for bench.py's e2e leg.

Every rank copies pinned host memory to its GPU and back at the same time on two streams —
plain cudaMemcpyAsync through torch, none of this repo's kernels — with the byte counts of one
bench.py e2e step per rank (537 MB up, 645 MB down by default).  Timed with CUDA events, max
over ranks; prints per-direction GB/s per rank and aggregate, and Msamples/s that ceiling allows
for the e2e leg (4 B/sample up, 4.8 B/sample down).

  python tools/pcie_probe.py                                   one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port 29511 tools/pcie_probe.py [--affinity]  N ranks at once

--affinity pins each rank to the CPUs of its GPU's NUMA node before allocating (first-touch puts
the pinned pages there).  --mode up|down|both selects the directions.
"""
import argparse
import json
import os
import subprocess

import torch


def numa_cpus(dev: int):
    """CPU list of the NUMA node the GPU hangs off (sysfs), or None."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(dev)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus            # 00000000:1B:00.0 -> 0000:1b:00.0
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read())
        cpus = open(base + "/local_cpulist").read().strip()
        out = []
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            out += list(range(int(a), int(b or a) + 1))
        return node, out
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--up-mb", type=float, default=536.9)
    ap.add_argument("--down-mb", type=float, default=644.6)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--mode", default="both", choices=["up", "down", "both"])
    ap.add_argument("--affinity", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    node, cpus = numa_cpus(local)
    if a.affinity and cpus:
        try:
            os.sched_setaffinity(0, cpus)
        except Exception:
            pass
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    up_n, down_n = int(a.up_mb * 1e6), int(a.down_mb * 1e6)
    h_up = torch.empty(up_n, dtype=torch.uint8, pin_memory=True)
    h_down = torch.empty(down_n, dtype=torch.uint8, pin_memory=True)
    h_up.fill_(1)
    d_up = torch.empty(up_n, dtype=torch.uint8, device="cuda")
    d_down = torch.zeros(down_n, dtype=torch.uint8, device="cuda")
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        if a.mode in ("up", "both"):
            with torch.cuda.stream(s_up):
                d_up.copy_(h_up, non_blocking=True)
        if a.mode in ("down", "both"):
            with torch.cuda.stream(s_down):
                h_down.copy_(d_down, non_blocking=True)

    for _ in range(2):
        once()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    s_up.wait_event(e0)
    s_down.wait_event(e0)
    for _ in range(a.reps):
        once()
    eu, ed = torch.cuda.Event(), torch.cuda.Event()
    eu.record(s_up)
    ed.record(s_down)
    torch.cuda.current_stream().wait_event(eu)
    torch.cuda.current_stream().wait_event(ed)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    t = torch.tensor([ms], device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
    ms_max = float(t.item())
    if rank == 0:
        up = up_n if a.mode in ("up", "both") else 0
        down = down_n if a.mode in ("down", "both") else 0
        line = {"ranks": world, "mode": a.mode, "affinity": bool(a.affinity and cpus), "numa_node_rank0": node,
                "cpus_rank0": f"{cpus[0]}-{cpus[-1]}" if cpus else None,
                "ms_per_step_max_over_ranks": round(ms_max, 3),
                "up_gbs_per_rank": round(up / ms_max / 1e6, 2), "down_gbs_per_rank": round(down / ms_max / 1e6, 2),
                "up_gbs_total": round(world * up / ms_max / 1e6, 2), "down_gbs_total": round(world * down / ms_max / 1e6, 2),
                # one e2e step per rank moves 256 ch x 2^19 samples: 4 B/sample up, 4.8 B/sample down
                "e2e_ceiling_msamples": round(world * (up_n / 4.0) / (ms_max * 1e-3) / 1e6, 1) if a.mode == "both" else None}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
