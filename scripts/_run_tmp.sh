python -m pytest tests -m gpu -q 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err; tail -c 300 gpurun_out/r1f_bench.err
