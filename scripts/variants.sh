for n in "$@"; do
  echo "== variant $n"
  RS_ENGINE_LIB=$PWD/rustsolver_b200/libb200cfr_$n.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['back_to_back_iter_per_sec'],1), round(d['e2e']['value'],1), d['roofline']['per_kernel_ms'])"
done
