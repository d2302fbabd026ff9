# One 8-GPU box: sharded parity at world 4 and 8, look-ahead chain profiles at N=8 / N=4, then the bench line at N=8.
mkdir -p gpurun_out
TAG=${TAG:-r02i}
(time timeout 400 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q -k "p2p-lookahead and (8 or 4)") > gpurun_out/${TAG}_tests_n8.log 2>&1; tail -4 gpurun_out/${TAG}_tests_n8.log
: > gpurun_out/${TAG}_prof.jsonl
for cfg in "8 16 8" "8 16 4" "8 8 4" "4 16 8" "4 16 4"; do set -- $cfg
  echo "{\"world\": $1, \"helpers\": $2, \"la_u\": $3}" >> gpurun_out/${TAG}_prof.jsonl
  B2S_LA_HELPERS=$2 B2S_LA_U=$3 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29561 tools/la_profile.py 8192 8192 200 2>/dev/null | grep '"rank": 0' >> gpurun_out/${TAG}_prof.jsonl
done
python - <<PY
import json
for ln in open('gpurun_out/${TAG}_prof.jsonl'):
    d=json.loads(ln)
    if 'kernel_us' not in d: print(d); continue
    print(d['world'], 'free', round(d['free_running_us_per_pivot'],1), 'kernel', round(d['kernel_us']['mean'],1), [round(d[k]['mean'],1) for k in ('rhs_row_us','entering_known_us','leaving_known_us','pivot_row_complete_us','proposal_ready_us','committed_us')])
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n8.json').read().strip().splitlines()[-1])
print('N=8 value', d['value'], 'e2e', d.get('e2e',{}).get('value'), d.get('e2e',{}).get('step'))
print('large', d.get('large_config')); print('parity', d.get('parity'))
PY
