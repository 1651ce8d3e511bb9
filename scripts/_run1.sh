set -x
python scripts/ab_hash.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core_updates or full_size_c2 or np_diag_kernels" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_stat_parity.py -x -q -m gpu -k "tensor_core_path" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_np7_bench.json 2> gpurun_out/r2_np7_bench.log
python -c "
import json; d=json.load(open('gpurun_out/r2_np7_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_issued'], d['roofline']['kernel_share_of_step'], d['checks'])"
QF_TRACE=1 timeout 300 python scripts/prof_step.py c2 37888 1 > gpurun_out/r2_np7_trace.log 2>&1
