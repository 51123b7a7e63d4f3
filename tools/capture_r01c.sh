#!/bin/bash
# Round-end GPU call: tests, smoke, both bench arms, the tree workload, launch list of the final build.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_c.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_b200_c.json 2> gpurun_out/bench_b200_c.err; echo "bench rc=$?"; cat gpurun_out/bench_b200_c.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_c.json 2> gpurun_out/bench_ref_c.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_c.json | cut -c1-400
timeout 600 python bench.py --workload tree > gpurun_out/bench_tree_c.json 2> gpurun_out/bench_tree_c.err; echo "tree rc=$?"; cat gpurun_out/bench_tree_c.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --pairs 200000 --steps 2 --warmup 3 --skip-cpu > gpurun_out/launches_r01c.log 2>&1
ls -la gpurun_out | tail -8
