set -x
python scripts/d2h_probe.py 2>&1 | tail -6
QF_CHUNK=18944 timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_np14_bench.json 2> gpurun_out/r2_np14_bench.log
python -c "
import json; d=json.load(open('gpurun_out/r2_np14_bench.json')); print('chunk18944', d['value'], d['e2e']['value'], d['ms_per_step'])"
