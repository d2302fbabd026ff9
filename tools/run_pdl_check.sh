mkdir -p gpurun_out
(time timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lookahead or options") > gpurun_out/r02o_tests.log 2>&1; tail -4 gpurun_out/r02o_tests.log
for pdl in 1 0; do for skip in 1 0; do
  B2S_LA_PDL=$pdl B2S_SKIP=$skip timeout 100 python tools/la_profile.py 8192 8192 50 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pdl=$pdl skip=$skip free', round(d['free_running_us_per_pivot'],2), 'committed', round(d['committed_us']['mean'],1))"
done; done
B2S_LA_PDL=1 B2S_LOOKAHEAD=1 timeout 100 python tools/la_profile.py 2048 2048 50 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('2048 pdl=1 free', round(d['free_running_us_per_pivot'],2), 'committed', round(d['committed_us']['mean'],1))"
timeout 200 python tools/loop_mode_sweep.py 2048,2048 4096,2048 4096,4096 8192,4096 > gpurun_out/r02o_loop_sweep.jsonl 2>&1; cat gpurun_out/r02o_loop_sweep.jsonl
