b() { python bench.py --steps 2 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['ms_per_frame'],2), 'e2e', round(j['e2e']['value']), 'cost', round(j['stages_ms']['cost_init'],3))"; }
for rep in 1 2; do
for lib in scripts/micro/libs/*.so; do echo -n "$(basename $lib .so): "; TSM_LIB=$PWD/$lib b; done
echo -n "default: "; b
done
