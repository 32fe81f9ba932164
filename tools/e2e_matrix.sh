for cfg in "" "BSB_HOST_THREADS=8" "BSB_HOST_THREADS=6" "BSB_HOST_THREADS=8 BSB_SPIN=1" "BSB_HOST_THREADS=12 BSB_SPIN=1"; do
  env $cfg python bench.py > gpurun_out/bench_m.log 2>gpurun_out/bench_m.err
  python - "$cfg" <<EOF
import json,sys
d=json.loads(open("gpurun_out/bench_m.log").read().strip().split("\n")[-1])
print(sys.argv[1] or "default", "| value %.2fM one %.2fM e2e %.2fM wall %.3f" % (d["value"]/1e6, d["value_one_batch_in_flight"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["wall_s"]), d["host_busy_s"])
EOF
done
