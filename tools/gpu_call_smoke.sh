mkdir -p gpurun_out
( timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.txt 2>&1; echo "exit $?" >> gpurun_out/r1_smoke.txt )
tail -5 gpurun_out/r1_smoke.txt
