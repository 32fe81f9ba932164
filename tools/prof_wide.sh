mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/t5_pytest.log 2>&1; tail -3 gpurun_out/t5_pytest.log
BSB_GPU_SLOTS=1 python tools/stage_times.py --batches 2 "" "BSB_REF_BLOCKS=1,BSB_SAMPLED_SA=1" "BSB_REF_BLOCKS=1,BSB_DENSE_SA40=1" > gpurun_out/t5_stage.log 2>&1; grep "^\[" gpurun_out/t5_stage.log
