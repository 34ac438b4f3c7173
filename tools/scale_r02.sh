# round-2 scaling runs on one multi-GPU box: the default bench line (jc_lindblad + the HEOM suite with the in-run
# sharded-vs-single parity check) at N = 8/4/2, plus the 2-GPU tests of tests/test_sharded.py
NG=${NG:-8}
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@" 2>>gpurun_out/scale_r02.err | grep "^{" ; }
for n in 8 4; do
  [ $n -le $NG ] || continue
  run $n --steps 3 --warmup 3 --no-cpu >> gpurun_out/r02_scale_default.jsonl
done
timeout 300 python -m pytest tests/test_sharded.py -q -m gpu 2>&1 | tail -3 > gpurun_out/r02_scale_tests.log
cat gpurun_out/r02_scale_tests.log
