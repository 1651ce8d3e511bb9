set -x
python scripts/_dbg_np.py 2>&1 | tail -4
python scripts/ab_hash.py 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_np11_bench.json 2> gpurun_out/r2_np11_bench.log
python -c "
import json; d=json.load(open('gpurun_out/r2_np11_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_issued'], d['roofline']['kernel_share_of_step'], d['checks'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2e.csv python scripts/prof_step.py c2 37888 1 > gpurun_out/r2_prof_e.log 2>&1
