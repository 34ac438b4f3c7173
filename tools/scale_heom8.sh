# one 8-GPU box: sharded FMO hierarchy with the automatic kernel choice at 8 and 4 ranks
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@" --no-cpu 2>>gpurun_out/scale_heom8.err | grep "^{" ; }
run 8 --workload heom_fmo --depth 4 --rk-steps 400 --steps 3 --warmup 3 >> gpurun_out/r02_scale_heom_8.jsonl
run 8 --workload heom_fmo --depth 6 --rk-steps 50 --steps 3 --warmup 3 >> gpurun_out/r02_scale_heom_8.jsonl
run 4 --workload heom_fmo --depth 4 --rk-steps 400 --steps 3 --warmup 3 >> gpurun_out/r02_scale_heom_8.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r02_scale_heom_8.jsonl'):
    d=json.loads(l); print(d['n_gpus'], d['config']['n_ado'], d['config']['sharding'][:60], '%.3g'%d['value'], '%.2f ms'%d['ms_per_step'], d['check'].get('sharded_vs_single_relerr'))
PY
tail -3 gpurun_out/scale_heom8.err
