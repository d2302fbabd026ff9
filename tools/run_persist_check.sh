mkdir -p gpurun_out
(time B2S_LA_PERSIST=1 timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lookahead or (stepping and lookahead) or options or drive_out") > gpurun_out/r02p_tests.log 2>&1; tail -5 gpurun_out/r02p_tests.log
for per in 0 1; do for skip in 1 0; do
  B2S_LA_PERSIST=$per B2S_SKIP=$skip timeout 100 python tools/la_profile.py 8192 8192 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('persist=$per skip=$skip free-running us/pivot', round(d['free_running_us_per_pivot'],2))"
done; done
