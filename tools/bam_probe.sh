# development aid: device BAM stage -- parity tests, e2e, per-kernel times
timeout 300 python -m pytest tests/test_bam_output.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/e2e_probe.py --bam --batches 12 "" "BSB_GPU_SLOTS=1" 2>&1 | tail -2
BSB_GPU_SLOTS=1 timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_bam|k_bgzf" --csv --log-file gpurun_out/r02_bam_launches.csv python tools/e2e_probe.py --bam --batches 2 --warm-batches 1 "" > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r02_bam_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); mi=hdr.index("Metric Name")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[(r[ki][:40], r[mi])].append(float(r[vi].replace(",","")))
    except: pass
for k,v in agg.items(): print(k, len(v), max(v)/1e6, "ms / M inst (largest launch)")
PY
