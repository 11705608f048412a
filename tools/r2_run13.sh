python tools/stream_battery_report.py > gpurun_out/r2_stream_battery.txt 2>&1; cat gpurun_out/r2_stream_battery.txt
python tools/bench_cvnn_widths.py > gpurun_out/r2_cvnn_widths.jsonl 2> gpurun_out/r2_cvnn_widths.err; cat gpurun_out/r2_cvnn_widths.jsonl; tail -3 gpurun_out/r2_cvnn_widths.err
