timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_vae.json 2> gpurun_out/bench_vae.err
tail -c 1500 gpurun_out/bench_vae.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_vae.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'])
print(json.dumps(d['roofline_scatter']))
print(d['kernel_classes_ms'])
PY
