P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["launches_per_step"])'
python -m pytest tests/test_vae_gpu.py tests/test_refine_reference_gpu.py tests/test_integration.py -m gpu -q -x 2>&1 | tail -2
python bench.py --no-extra --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
python bench.py --no-extra --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
