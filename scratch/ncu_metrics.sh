KVFE_BATCH=32 KVFE_STEPS=7 KVFE_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/metrics_b32.csv python profiles/profile_step.py > gpurun_out/pm.log 2>&1
tail -2 gpurun_out/pm.log
