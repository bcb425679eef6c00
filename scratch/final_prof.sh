set -x
export KVFE_NO_GRAPH=1 KVFE_BATCH=32
KVFE_STEPS=10 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_final.csv python profiles/profile_step.py > gpurun_out/pf1.log 2>&1
KVFE_STEPS=7 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r01_metrics_final.csv python profiles/profile_step.py > gpurun_out/pf2.log 2>&1
KVFE_STEPS=8 ncu --set full --clock-control none --import-source on -k regex:"lk_kernel_col|match_kernel|mineig_kernel|rectify_kernel|pyr_level_kernel|sort_greedy" -s 66 -c 11 -f -o gpurun_out/prof_r01_final python profiles/profile_step.py > gpurun_out/pf3.log 2>&1
tail -2 gpurun_out/pf3.log
ls -la gpurun_out/prof_r01_final.ncu-rep
