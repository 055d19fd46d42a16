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
  # ---- round 2 ----
  r2_final)   # what profiles/r2_pytest_gpu_final.txt, r2_smoke.txt and r2_bench_1gpu.json came from (one call, ~5.5 min)
    python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.txt 2>&1; tail -4 gpurun_out/r2_pytest_gpu_final.txt
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -2 gpurun_out/r2_smoke.txt
    python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 600 gpurun_out/r2_bench_1gpu.json ;;
  r2_benchN)  # gpurun --gpus N: bash tools/gpurun_recipes.sh r2_benchN N
    N="${2:-2}"
    timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus "$N" --steps 5 --warmup 3 > "gpurun_out/r2_bench_${N}gpu.json" 2> "gpurun_out/r2_bench_${N}gpu.err"
    tail -c 700 "gpurun_out/r2_bench_${N}gpu.json" ;;
  r2_ab)      # the A/B tools of the round, each in one process: GEMM kernels, QR / Cholesky option sweeps, host-view e2e
    timeout 600 python tools/gemm_option_ab.py gemm_tma2:2 gemm_tma2_maxk=100000 | tee gpurun_out/r2_gemm_tma2b.jsonl
    timeout 300 python tools/sweep.py qr 16384 3 "-" "hr_split=0" "qr_fold_t=0" "qr_vt=0" "qr_panel_cholqr=0" "gemm_tma2=0" "prof=1" | tee gpurun_out/r2_qr_ab.jsonl
    timeout 300 python tools/sweep.py chol 16384 5 "-" "gemm_tma2=0" "chol_split_panel=0" "chol_tn=0" | tee gpurun_out/r2_chol_ab.jsonl
    timeout 300 python tools/chol_e2e.py 16384 2 pinned C | tee gpurun_out/r2_chol_e2e_pinned_C.jsonl
    timeout 300 python tools/qr_e2e.py 16384 2 | tee gpurun_out/r2_qr_e2e.jsonl ;;
  r2_trace)   # look-ahead pipelines and the panel kernels' phases
    timeout 100 python tools/factor_once.py qr 16384 16384 qr_trace=1 warm=1 2> gpurun_out/r2_qr_trace.txt
    timeout 100 python tools/factor_once.py chol 16384 16384 chol_trace=1 warm=1 2> gpurun_out/r2_chol_trace.txt
    LFB_PANEL_DBG=1 timeout 100 python tools/factor_once.py qr 128 16384 warm=1 2>&1 | tail -3
    LFB_WAVE_DBG=1 timeout 200 python tools/chol_e2e.py 16384 1 pinned F 2>&1 | grep -v "^{" | head -12 ;;
  r2_ncu)     # ncu evidence of the two-CTA GEMM and the bench launch list
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma2 -s 2 -c 1 -f -o gpurun_out/r2_dgemm_tma2_syrk python tools/gemm_one.py 1 0 8192 8192 512 3 > /dev/null 2>&1
    timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dgemm --csv --log-file gpurun_out/r2b_chol_traffic.csv python tools/factor_once.py chol 16384 16384 warm=1 lookahead=0 > /dev/null 2>&1
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/r2_bench_under_ncu.log 2>&1
    python tools/ncu_agg.py gpurun_out/r2_bench_launches.csv | head -14 ;;
  *) echo "usage: $0 {tests|bench|bench2|ncu_tsqr|smoke|r2_final|r2_benchN N|r2_ab|r2_trace|r2_ncu}"; exit 2 ;;
esac
