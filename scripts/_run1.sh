set -x
python scripts/_dbg_np.py 2>&1 | tail -5
timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_np5_bench.json 2> gpurun_out/r2_np5_bench.log
head -c 250 gpurun_out/r2_np5_bench.json; echo
python -c "
import json; d=json.load(open('gpurun_out/r2_np5_bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_issued'], d['roofline']['kernel_share_of_step'])"
QF_CHUNK=18944 timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --batch 75776 > gpurun_out/r2_np5_bench_c1x.json 2> gpurun_out/r2_np5_bench_c1x.log
head -c 250 gpurun_out/r2_np5_bench_c1x.json; echo
QF_CHUNK=75776 timeout 900 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_np5_bench_c4x.json 2> gpurun_out/r2_np5_bench_c4x.log
head -c 250 gpurun_out/r2_np5_bench_c4x.json; echo
