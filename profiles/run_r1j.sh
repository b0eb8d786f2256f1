set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r1j.log 2>&1; tail -3 gpurun_out/pytest_gpu_r1j.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r1j.log 2>&1; tail -2 gpurun_out/smoke_r1j.log
python bench.py > gpurun_out/bench_r1j.log 2> gpurun_out/bench_r1j.err; tail -c 300 gpurun_out/bench_r1j.log
python profiles/bench_configs.py > gpurun_out/bench_configs_r1j.jsonl 2>&1
PK_NO_CERTIFICATE=1 python profiles/bench_configs.py > gpurun_out/bench_configs_r1j_nocert.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_r1j.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pk_step_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1j python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1j.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pk_step_cluster_kernel -s 1 -c 1 -f -o gpurun_out/prof_r1j_c4 python profiles/run_c4.py 18 10 > gpurun_out/ncu_r1j_c4.log 2>&1
