mkdir -p gpurun_out
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1_tsqr_hr_launches.csv python tools/tsqr_hr_once.py 131072 256 > gpurun_out/r1_tsqr_hr_ncu.log 2>&1
echo "exit $?"; tail -2 gpurun_out/r1_tsqr_hr_ncu.log; wc -l gpurun_out/r1_tsqr_hr_launches.csv
