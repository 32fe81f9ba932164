mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/t4_pytest.log 2>&1; tail -3 gpurun_out/t4_pytest.log
python tools/stage_times.py --batches 2 "" > gpurun_out/t4_stage.log 2>&1; tail -2 gpurun_out/t4_stage.log
BSB_GPU_SLOTS=1 ncu --set full --clock-control none --import-source on -k regex:"k_extend_warp|k_chain_warp" -c 2 -o gpurun_out/t4_prof -f python tools/stage_times.py --batches 1 "" > gpurun_out/t4_ncu.log 2>&1
ls -la gpurun_out/t4_prof.ncu-rep
