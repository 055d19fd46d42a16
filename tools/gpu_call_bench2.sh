mkdir -p gpurun_out
( timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r1_bench_2gpu.json 2> gpurun_out/r1_bench_2gpu.err; echo "exit $?" >> gpurun_out/r1_bench_2gpu.err )
tail -c 700 gpurun_out/r1_bench_2gpu.json; tail -3 gpurun_out/r1_bench_2gpu.err
