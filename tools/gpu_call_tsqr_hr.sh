mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_gpu_tsqr_hr.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r1_tsqr_hr_tests.txt 2>&1; echo "exit $?" >> gpurun_out/r1_tsqr_hr_tests.txt )
( timeout 60 python tools/tsqr_qr_time.py 1048576 256 > gpurun_out/r1_tsqr_qr_time.json 2> gpurun_out/r1_tsqr_qr_time.err; echo "exit $?" >> gpurun_out/r1_tsqr_qr_time.err )
tail -30 gpurun_out/r1_tsqr_hr_tests.txt; cat gpurun_out/r1_tsqr_qr_time.json; tail -3 gpurun_out/r1_tsqr_qr_time.err
