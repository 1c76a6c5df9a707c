mkdir -p gpurun_out
timeout 80 python tools/bench_stream.py --frames 200 > gpurun_out/stream_config4.json 2> gpurun_out/stream_config4.err; tail -c 900 gpurun_out/stream_config4.json; tail -3 gpurun_out/stream_config4.err
