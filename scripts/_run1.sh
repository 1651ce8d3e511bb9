set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core_updates or structured_basis_matches_dense or full_size_c2" > gpurun_out/r2_np2_tests.log 2>&1
tail -5 gpurun_out/r2_np2_tests.log
QF_TRACE=1 timeout 300 python scripts/prof_step.py c2 18944 1 > gpurun_out/r2_np2_trace.log 2>&1
tail -3 gpurun_out/r2_np2_trace.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_np2_bench.json 2> gpurun_out/r2_np2_bench.log
cat gpurun_out/r2_np2_bench.json | head -c 600
