timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv -c 1400 --log-file gpurun_out/launches_render_warm.csv python bench.py --workload render --no-graph --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/launches_render_warm.csv
