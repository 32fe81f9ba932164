# usage: quick_bench.sh STEPS "ENV=1 ENV2=2" ...   -> one bench.py run per environment string
mkdir -p gpurun_out
steps=$1; shift
for cfg in "$@"; do
  env $cfg python bench.py --steps $steps --cpu-sample-pairs 20000 > gpurun_out/qb.log 2> gpurun_out/qb.err
  python - "$cfg" <<'PY'
import json, sys
l = [x for x in open('gpurun_out/qb.log') if x.startswith('{')]
if not l:
    print(sys.argv[1], 'FAILED', open('gpurun_out/qb.err').read()[-1500:])
else:
    d = json.loads(l[-1])
    print(f"[{sys.argv[1]}] value {d['value']/1e6:.2f} M  e2e {d['e2e']['value']/1e6:.2f} M (wall {d['e2e']['wall_s']:.2f} s)  one-slot {d['value_one_batch_in_flight']/1e6:.2f} M  host {d['host_busy_s']}  parity {d['parity_vs_reference']['identical_frac'] if d['parity_vs_reference'] else None}")
PY
done
