set -x
python scripts/_dbg_np.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core_updates or full_size_c2 or np_diag_kernels" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_stat_parity.py -x -q -m gpu -k "tensor_core_path or two_phase" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_np10_bench.json 2> gpurun_out/r2_np10_bench.log
python -c "
import json; d=json.load(open('gpurun_out/r2_np10_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_issued'], d['roofline']['kernel_share_of_step'], d['checks'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2d.csv python scripts/prof_step.py c2 37888 1 > gpurun_out/r2_prof_d.log 2>&1
