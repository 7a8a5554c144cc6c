show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['solve']; print('$1', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],3), 'it', s['cg_iterations_per_step'], 'ms/it', round(s['ms_per_cg_iteration'],4), 'prod', round(s['ms_product_per_iteration'],4), 'step', round(s['ms_vector_step_per_iteration'],4))"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline 2>&1 | show default
LSS_PQ_IN_STEP=1 timeout 300 python bench.py --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline 2>&1 | show pq_in_step
LSS_NO_L2_WINDOW=1 timeout 300 python bench.py --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline 2>&1 | show no_l2_window
