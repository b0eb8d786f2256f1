set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r1g.log 2>&1; tail -3 gpurun_out/pytest_gpu_r1g.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r1g.log 2>&1; tail -2 gpurun_out/smoke_r1g.log
python bench.py > gpurun_out/bench_r1g.log 2> gpurun_out/bench_r1g.err; tail -c 600 gpurun_out/bench_r1g.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1g.log 2>&1; tail -c 400 gpurun_out/bench_ref_r1g.log
python profiles/bench_configs.py > gpurun_out/bench_configs_r1g.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_r1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pk_step_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1g python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pk_upwind_kernel -s 1 -c 1 -f -o gpurun_out/prof_r1g_upwind python profiles/run_upwind.py > gpurun_out/ncu_r1g_upwind.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pk_step_cluster_kernel -s 1 -c 1 -f -o gpurun_out/prof_r1g_c4 python profiles/run_c4.py 18 10 > gpurun_out/ncu_r1g_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
