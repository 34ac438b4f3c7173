# sharded FMO hierarchy: kernel variants at N ranks (N = $NG); EX = list of exchange modes, DEPTHS = list of depths
NG=${NG:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG "$@" --no-cpu 2>>gpurun_out/scale_heom.err | grep "^{" ; }
for ex in ${EX:-auto p2p}; do
  for d in ${DEPTHS:-4 6}; do
    if [ $d = 4 ]; then rk=400; else rk=50; fi
    run --workload heom_fmo --depth $d --rk-steps $rk --exchange $ex --steps 3 --warmup 3 >> gpurun_out/r02_scale_heom_$NG.jsonl
  done
done
python - <<'PY'
import json,os
for l in open('gpurun_out/r02_scale_heom_%s.jsonl' % os.environ.get('NG','2')):
    d=json.loads(l); print(d['n_gpus'], d['config']['n_ado'], d['config']['sharding'][:60], '%.3g'%d['value'], '%.2f ms'%d['ms_per_step'], d['check'].get('sharded_vs_single_relerr'))
PY
tail -3 gpurun_out/scale_heom.err
