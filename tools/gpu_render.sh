timeout 600 python -m pytest tests/test_refine_loss.py tests/test_raster_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --workload render --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_render.json 2> gpurun_out/bench_render.err
tail -c 400 gpurun_out/bench_render.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_render.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'], d.get('kernel_classes_ms'), d.get('first_loss'), d.get('loss_after_timed_iters'))
PY
