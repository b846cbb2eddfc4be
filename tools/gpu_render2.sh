timeout 600 python -m pytest tests/test_refine_loss.py tests/test_raster_gpu.py -m gpu -q 2>&1 | tail -8
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:k_backward_rgb -s 3 -c 1 -o gpurun_out/prof_bwd_rgb python bench.py --workload render --no-graph --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/prof_bwd_rgb.ncu-rep --page raw --csv > gpurun_out/prof_bwd_rgb_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_bwd_rgb.ncu-rep --page source --csv > gpurun_out/prof_bwd_rgb_source.csv 2>/dev/null
ls -la gpurun_out/prof_bwd_rgb*
