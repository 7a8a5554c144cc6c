# A/B of the normal-matrix product kernels and the CG vector-step kernels (one gpurun call)
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['solve']; print('$1', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],3), 'it', s['cg_iterations_per_step'], 'ms/it', round(s['ms_per_cg_iteration'],4), 'prod', round(s['ms_product_per_iteration'],4), 'step', round(s['ms_vector_step_per_iteration'],4))" 2>&1 | tail -1; }
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "variants_agree" > gpurun_out/ab2_test.log 2>&1; tail -3 gpurun_out/ab2_test.log
B="timeout 200 python bench.py --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline"
LSS_N_PRODUCT=11 LSS_TMA_STAGES=2 $B 2>&1 | show tmad2_entry
LSS_N_PRODUCT=11 LSS_TMA_STAGES=3 $B 2>&1 | show tmad3_entry_pipe
