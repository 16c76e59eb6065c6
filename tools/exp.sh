for cfg in 0 2; do for xf in 0 1 2 3; do
SB200_XFLAGS=$xf SB200_STAGGER=6000 SB200_PERSIST_CFG=$cfg timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg',$cfg,'xflags',$xf, 'ms', round(d['ms_per_step'],4), 'warm', round(2097152/d['value_l2_warm']/1e6,4), 'gdof', round(d['value'],2))"
done; done
