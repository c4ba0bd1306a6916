for c in "jpegl 1 512 768" "jpegl 24 512 768" "two_layer_syn2 8 1200 1200" "two_layer_syn2:24 8 1200 1200" "two_layer_syn2:48 8 1200 1200" "mbt2018 24 512 768" "bls2017 2 2160 3840"; do set -- $c; echo "== $c"; timeout 200 python bench.py --config $1 --batch $2 --height $3 --width $4 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -c 3000 | python -c "
import sys,json
for l in sys.stdin:
  l=l.strip()
  if l.startswith('{'):
    d=json.loads(l); print(json.dumps({k:d[k] for k in ('value','ms_per_step')}), 'e2e', round(d['e2e']['value'])); print(json.dumps(d['config']['layers_ms'])); r=d['roofline']; print(r['kernel'], round(r['achieved'],1), round(r['frac'],4))
  else: print(l[-400:])
"; done
