# Round-end evidence on the C2 workload: launch list of the bench command and one ncu --set full pass over every per-batch kernel
mkdir -p gpurun_out
BSB_GPU_SLOTS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/final_launches_bench.log 2>&1
BSB_GPU_SLOTS=1 ncu --set full --clock-control none --import-source on -k regex:"k_seed3|k_chain_warp|k_extend_warp|k_final_pe|k_tasks_dp|k_tasks_finish|k_sa|k_convert|k_pack4|k_pestat|k_sam_write|k_sam_count" -c 14 -o gpurun_out/final_prof -f python tools/stage_times.py --batches 1 "" > gpurun_out/final_prof.log 2>&1
ls -la gpurun_out/final_prof.ncu-rep gpurun_out/final_launches.csv
# the report itself is larger than what travels back: keep the summaries
python tools/ncu_summary.py gpurun_out/final_prof.ncu-rep --md > gpurun_out/final_prof.md 2>&1
python tools/ncu_summary.py gpurun_out/final_prof.ncu-rep --lines k_chain_warp > gpurun_out/final_prof_chain_lines.txt 2>&1
rm -f gpurun_out/final_prof.ncu-rep
