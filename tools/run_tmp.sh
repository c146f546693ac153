mkdir -p gpurun_out/r2c
timeout 200 python tools/wavefront_trace.py gpurun_out/r2c/wavefront_trace_astat.md 2>&1 | grep -v "^$" | tail -22
timeout 200 python tools/trace_step.py gpurun_out/r2c/trace_graph_timeline.md --graph > /dev/null 2>gpurun_out/r2c/trace.err; tail -3 gpurun_out/r2c/trace.err
