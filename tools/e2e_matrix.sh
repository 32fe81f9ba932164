#!/bin/bash
# development aid: bench.py under different host/device settings, one summary line each
for cfg in "$@"; do
  env $cfg python bench.py --steps ${STEPS:-8} > gpurun_out/bench_m.log 2>gpurun_out/bench_m.err
  python - "$cfg" <<EOF
import json,sys
d=json.loads(open("gpurun_out/bench_m.log").read().strip().split("\n")[-1])
print(sys.argv[1] or "default", "| value %.2fM one %.2fM e2e %.2fM wall %.3f" % (d["value"]/1e6, d["value_one_batch_in_flight"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["wall_s"]), d["host_busy_s"])
EOF
done
