set -x
python scripts/ab_hash.py 2>&1 | tail -3 > gpurun_out/ab_default.txt
QF_I8_EPI_STAGE=0 python scripts/ab_hash.py 2>&1 | tail -3 > gpurun_out/ab_nostage.txt
cat gpurun_out/ab_default.txt gpurun_out/ab_nostage.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_np6_bench.json 2> gpurun_out/r2_np6_bench.log
python -c "
import json; d=json.load(open('gpurun_out/r2_np6_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_issued'], d['roofline']['kernel_share_of_step'])"
QF_TRACE=1 timeout 300 python scripts/prof_step.py c2 37888 1 > gpurun_out/r2_np6_trace.log 2>&1
