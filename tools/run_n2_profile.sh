mkdir -p gpurun_out
(time timeout 500 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q) > gpurun_out/r02h_tests_n2.log 2>&1; tail -4 gpurun_out/r02h_tests_n2.log
: > gpurun_out/r02h_prof_n2.jsonl
for cfg in "8 8" "16 8" "8 4" "16 4"; do set -- $cfg
  echo "{\"helpers\": $1, \"la_u\": $2}" >> gpurun_out/r02h_prof_n2.jsonl
  B2S_LA_HELPERS=$1 B2S_LA_U=$2 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29560 tools/la_profile.py 8192 8192 200 2>/dev/null | grep '"rank": 0' >> gpurun_out/r02h_prof_n2.jsonl
done
for cfg in "8 4" "16 4"; do set -- $cfg
  echo "{\"n1_helpers\": $1, \"la_u\": $2}" >> gpurun_out/r02h_prof_n2.jsonl
  B2S_LA_HELPERS=$1 B2S_LA_U=$2 timeout 100 python tools/la_profile.py 8192 8192 200 2>/dev/null >> gpurun_out/r02h_prof_n2.jsonl
done
python - <<'PY'
import json
for ln in open('gpurun_out/r02h_prof_n2.jsonl'):
    d=json.loads(ln)
    if 'kernel_us' not in d: print(d); continue
    print(d['world'], 'free', round(d['free_running_us_per_pivot'],1), 'kernel', round(d['kernel_us']['mean'],1), [round(d[k]['mean'],1) for k in ('rhs_row_us','entering_known_us','leaving_known_us','pivot_row_complete_us','proposal_ready_us','committed_us')])
PY
