mkdir -p gpurun_out
( timeout 280 python -m pytest tests -q -m gpu --tb=line -p no:cacheprovider > gpurun_out/r1_pytest_gpu.txt 2>&1; echo "exit $?" >> gpurun_out/r1_pytest_gpu.txt )
tail -8 gpurun_out/r1_pytest_gpu.txt
