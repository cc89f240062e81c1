set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/bench_all.sh
python bench.py > gpurun_out/r1_final_bench_default.log 2> gpurun_out/r1_final_bench_default.err; tail -c 400 gpurun_out/r1_final_bench_default.log
