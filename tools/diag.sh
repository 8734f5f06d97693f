for c in 0 1; do
JSDR_PUMP_CONCURRENT=$c python bench.py --steps 6 --no-cpu --e2e-channels 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('concurrent $c step',d['ms_per_step'], r['kernel'][:12], r['ms_per_launch'], [ (k['kernel'][:10],k['ms_per_launch']) for k in r['other_kernels']], d['variants']['decimator_f32']['ms_per_step'])"
done
