mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/quick_pytest.log 2>&1; grep "passed\|failed" gpurun_out/quick_pytest.log
BSB_GPU_SLOTS=1 python tools/stage_times.py --batches 2 "$@" > gpurun_out/quick_stage.log 2>&1; grep "^\[" gpurun_out/quick_stage.log || tail -5 gpurun_out/quick_stage.log
