timeout 600 python -m pytest tests/test_refine_loss.py tests/test_raster_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload render --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_render.json 2> gpurun_out/bench_render.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_render.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d.get('kernel_classes_ms'), d.get('loss_after_timed_iters'))
PY
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -f -k regex:'k_raster_tiles|k_backward_rgb_cta' -s 6 -c 2 -o gpurun_out/prof_raster2 python bench.py --workload render --no-graph --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/prof_raster2.ncu-rep --page raw --csv > gpurun_out/prof_raster2_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_raster2.ncu-rep --page source --csv > gpurun_out/prof_raster2_source.csv 2>/dev/null
ls -la gpurun_out/prof_raster2*
