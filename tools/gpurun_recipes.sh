#!/usr/bin/env bash
# The gpurun commands this round's numbers came from (each is one call: `gpurun --timeout T -- 'bash tools/gpurun_recipes.sh <name>'`).
# Everything lands in gpurun_out/ (scratch); what is judged is copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
case "${1:-}" in
  tests)      # full GPU parity suite (~2 min)
    timeout 400 python -m pytest tests -q -m gpu --tb=line -p no:cacheprovider > gpurun_out/r1_pytest_gpu.txt 2>&1; echo "exit $?" >> gpurun_out/r1_pytest_gpu.txt
    tail -8 gpurun_out/r1_pytest_gpu.txt ;;
  bench)      # default bench line + stage timing of the tall-skinny route
    timeout 170 python bench.py > gpurun_out/r1_bench_1gpu.json 2> gpurun_out/r1_bench_1gpu.err; echo "exit $?" >> gpurun_out/r1_bench_1gpu.err
    timeout 40 python tools/tsqr_qr_time.py 1048576 256 > gpurun_out/r1_tsqr_qr_time.json 2> gpurun_out/r1_tsqr_qr_time.err
    tail -c 600 gpurun_out/r1_bench_1gpu.json; cat gpurun_out/r1_tsqr_qr_time.json ;;
  bench2)     # gpurun --gpus 2
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r1_bench_2gpu.json 2> gpurun_out/r1_bench_2gpu.err
    tail -c 700 gpurun_out/r1_bench_2gpu.json ;;
  ncu_tsqr)   # launch list of one tall-skinny qr_into
    timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1_tsqr_hr_launches.csv python tools/tsqr_hr_once.py 131072 256 > gpurun_out/r1_tsqr_hr_ncu.log 2>&1
    wc -l gpurun_out/r1_tsqr_hr_launches.csv ;;
  smoke)
    timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.txt 2>&1; echo "exit $?" >> gpurun_out/r1_smoke.txt; tail -3 gpurun_out/r1_smoke.txt ;;
  *) echo "usage: $0 {tests|bench|bench2|ncu_tsqr|smoke}"; exit 2 ;;
esac
