for g in 148 74 37 18; do
SB200_CHAIN_GRID=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('grid',$g, 'ms', d['ms_per_step'], 'warm', 2097152/d['config']['value_l2_warm']/1e6)"
done
