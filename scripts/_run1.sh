set -x
python scripts/_dbg_np.py 2>&1 | tail -4
python scripts/ab_hash.py 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_np16_bench.json 2> gpurun_out/r2_np16_bench.log
python -c "
import json; d=json.load(open('gpurun_out/r2_np16_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['checks'])"
