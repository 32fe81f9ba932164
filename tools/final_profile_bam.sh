mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02_final_pytest.log 2>&1; tail -3 gpurun_out/r02_final_pytest.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/r02_final_bench.err | grep "^{" > gpurun_out/r02_final_bench.json; tail -c 300 gpurun_out/r02_final_bench.err
BSB_GPU_SLOTS=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_sam|k_bam|k_bgzf" --csv --log-file gpurun_out/r02_final_launches_bam.csv python tools/e2e_probe.py --bam --batches 2 --warm-batches 1 "" > /dev/null 2>&1
BSB_GPU_SLOTS=1 ncu --set full --clock-control none --import-source on -k regex:"k_bgzf_deflate" --launch-skip 1 -c 1 -o gpurun_out/r02_final_deflate -f python tools/e2e_probe.py --bam --batches 2 --warm-batches 1 "" > /dev/null 2>&1
BSB_DF_PROFILE=1 BSB_GPU_SLOTS=1 python tools/e2e_probe.py --bam --batches 2 --warm-batches 1 "" 2>&1 | grep -E "deflate|rc=" | tail -2 > gpurun_out/r02_final_deflate_phases.log
