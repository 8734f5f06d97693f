python tools/kbench.py fft 4096 2>&1 | tail -2
python bench.py --steps 4 --no-cpu --e2e-channels 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('prio  step',d['ms_per_step'], r['kernel'][:12], r['ms_per_launch'], [ (k['kernel'][:10],k['ms_per_launch']) for k in r['other_kernels']], d['variants'])"
JSDR_SIDE_PRIORITY=0 python bench.py --steps 4 --no-cpu --e2e-channels 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('noprio step',d['ms_per_step'], r['kernel'][:12], r['ms_per_launch'], [ (k['kernel'][:10],k['ms_per_launch']) for k in r['other_kernels']], d['variants'])"
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader
