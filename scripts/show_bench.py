import json, sys, glob
for f in sorted(glob.glob(sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        pk = d['roofline']['per_kernel_ms']
        print(f.split('/')[-1], 'iter/s', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), {k: round(v, 4) for k, v in pk.items()}, 'whole', round(d['roofline']['whole_iteration_frac'], 4))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-600:])
