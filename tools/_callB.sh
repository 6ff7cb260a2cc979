mkdir -p gpurun_out
# (1) ncu --set full of the new cross-attention kernel (after the plain probe has exited 0)
ROWS=2048 timeout 120 python tools/gpu_probe_xattn.py > gpurun_out/xattn_probe_plain.log 2>&1 && \
ROWS=2048 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dec_cross_tc -s 8 -c 1 -o gpurun_out/r02_xattn_tc python tools/gpu_probe_xattn.py > gpurun_out/ncu_xattn.log 2>&1
ncu -i gpurun_out/r02_xattn_tc.ncu-rep --page raw --csv > gpurun_out/r02_xattn_tc_raw.csv 2>/dev/null
# (2) launch list of the bench command (1 warm-up + 1 step, 16 pages so that the profiled run ends in minutes)
python bench.py --pages 16 --steps 1 --warmup 1 --no-cpu-baseline --no-second-dtype > gpurun_out/bench_16p_plain.json 2> gpurun_out/bench_16p_plain.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_config1.csv python bench.py --pages 16 --steps 1 --warmup 1 --no-cpu-baseline --no-second-dtype > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
python tools/ncu_launch_summary.py gpurun_out/r02_launches_bench_config1.csv | head -40
