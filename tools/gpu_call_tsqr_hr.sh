mkdir -p gpurun_out
( timeout 60 python tools/tsqr_hr_debug.py diag > gpurun_out/dbg_diag.txt 2>&1; echo "exit $?" >> gpurun_out/dbg_diag.txt )
( timeout 60 python tools/tsqr_hr_debug.py zero_col > gpurun_out/dbg_zero_col.txt 2>&1; echo "exit $?" >> gpurun_out/dbg_zero_col.txt )
( timeout 150 python -m pytest tests/test_gpu_tsqr_hr.py -q -m gpu --tb=line -p no:cacheprovider > gpurun_out/r1_tsqr_hr_tests.txt 2>&1; echo "exit $?" >> gpurun_out/r1_tsqr_hr_tests.txt )
tail -5 gpurun_out/r1_tsqr_hr_tests.txt
