set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_final.csv python scripts/prof_step.py c2 37888 1 > gpurun_out/r2_prof_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:np_diag2 -s 30 -c 1 -o gpurun_out/prof_npdiag2_r2 -f python scripts/prof_step.py c2 37888 1 > gpurun_out/ncu_np2_r2.log 2>&1
QF_TRACE=1 timeout 300 python scripts/prof_step.py c2 37888 1 > gpurun_out/r2_final_trace.log 2>&1
