python scripts/diag_trace.py 8 48 streams > gpurun_out/r1e_trace_streams.log 2>&1
python scripts/diag_trace.py 8 48 pipeline > gpurun_out/r1e_trace_pipeline.log 2>&1
python bench.py --no-cpu-baseline > gpurun_out/r1e_bench2.json 2> gpurun_out/r1e_bench2.err
tail -c 300 gpurun_out/r1e_bench2.err; head -12 gpurun_out/r1e_trace_streams.log; head -12 gpurun_out/r1e_trace_pipeline.log
