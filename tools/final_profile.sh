# round-end evidence on one B200: tests, bench line, launch list, ncu captures of the two largest kernels
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02_final_pytest.log 2>&1; tail -3 gpurun_out/r02_final_pytest.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/r02_final_bench.err | grep "^{" > gpurun_out/r02_final_bench.json; tail -c 300 gpurun_out/r02_final_bench.err
python bench.py --impl reference --steps 20 --warmup 5 2> gpurun_out/r02_final_bench_ref.err | grep "^{" > gpurun_out/r02_final_bench_ref.json
BSB_GPU_SLOTS=1 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 0 -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python tools/stage_times.py --batches 2 "" > /dev/null 2>&1
BSB_GPU_SLOTS=1 ncu --set full --clock-control none --import-source on -k regex:"k_seed3<|k_extend_lanes|k_chain_warp" -c 3 -o gpurun_out/r02_final_top3 -f python tools/stage_times.py --batches 1 "" > gpurun_out/r02_final_ncu.log 2>&1
ls -la gpurun_out/r02_final_top3.ncu-rep
