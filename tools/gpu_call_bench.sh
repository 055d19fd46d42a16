mkdir -p gpurun_out
( timeout 170 python bench.py > gpurun_out/r1_bench_1gpu.json 2> gpurun_out/r1_bench_1gpu.err; echo "exit $?" >> gpurun_out/r1_bench_1gpu.err )
( timeout 40 python tools/tsqr_qr_time.py 1048576 256 > gpurun_out/r1_tsqr_qr_time.json 2> gpurun_out/r1_tsqr_qr_time.err; echo "exit $?" >> gpurun_out/r1_tsqr_qr_time.err )
tail -c 1500 gpurun_out/r1_bench_1gpu.json; tail -3 gpurun_out/r1_bench_1gpu.err; cat gpurun_out/r1_tsqr_qr_time.json
