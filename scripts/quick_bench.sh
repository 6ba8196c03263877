for w in "$@"; do
  echo "== $w"
  python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('iter/s', round(d['value'],1), 'b2b', round(d['back_to_back_iter_per_sec'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), d['roofline']['per_kernel_ms'], 'upd/iter', d['config']['updates_per_iteration'], 'create_s', round(d['engine_create_s'],2))"
done
