set -x
python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 300 gpurun_out/final_bench_n1.json
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 400 gpurun_out/final_bench_ref.json
export KVFE_NO_GRAPH=1 KVFE_BATCH=32
KVFE_STEPS=10 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_final.csv python profiles/profile_step.py > gpurun_out/pf1.log 2>&1
KVFE_STEPS=7 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r01_metrics_final.csv python profiles/profile_step.py > gpurun_out/pf2.log 2>&1
tail -1 gpurun_out/pf2.log
