mkdir -p gpurun_out
python -m pytest "$@" -x -q -m gpu > gpurun_out/quick_pytest.log 2>&1; grep -v "^\[M::" gpurun_out/quick_pytest.log | tail -40 | cut -c1-400
