# usage: prof_kernel.sh <kernel regex> <out name>   -> ncu --set full with sources of one launch on the C2 workload
mkdir -p gpurun_out
BSB_GPU_SLOTS=1 ncu --set full --clock-control none --import-source on -k regex:"$1" -c 1 -o gpurun_out/$2 -f python tools/stage_times.py --batches 1 "" > gpurun_out/$2.log 2>&1
ls -la gpurun_out/$2.ncu-rep
