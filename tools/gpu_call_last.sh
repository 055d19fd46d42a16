mkdir -p gpurun_out
( timeout 60 python -m pytest tests/test_gpu_lobpcg_blocks.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r1_lobpcg_blocks_tests.txt 2>&1; echo "exit $?" >> gpurun_out/r1_lobpcg_blocks_tests.txt )
tail -40 gpurun_out/r1_lobpcg_blocks_tests.txt
