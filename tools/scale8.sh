run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@" --no-cpu 2>>gpurun_out/scale8.err | grep "^{" ; }
for n in 8 4 2; do
  run $n --steps 3 --warmup 3 >> gpurun_out/scale_jc.jsonl
  run $n --workload heom_fmo --exchange p2p --steps 3 --warmup 3 >> gpurun_out/scale_fmo4_p2p.jsonl
  run $n --workload heom_fmo --exchange nccl --steps 3 --warmup 3 >> gpurun_out/scale_fmo4_nccl.jsonl
  run $n --workload heom_fmo --depth 6 --rk-steps 20 --exchange p2p --steps 3 --warmup 3 >> gpurun_out/scale_fmo6_p2p.jsonl
  run $n --workload heom_fmo --depth 6 --rk-steps 20 --exchange nccl --steps 3 --warmup 3 >> gpurun_out/scale_fmo6_nccl.jsonl
done
timeout 300 python -m pytest tests/test_sharded.py -q 2>&1 | tail -2 > gpurun_out/scale8_tests.log
