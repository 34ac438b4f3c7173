#!/bin/bash
# Round-2 fourth GPU pass (2 GPUs): the 2-GPU tests of tests/test_sharded.py and the default bench line at N = 2
# (sharded FMO hierarchy: dataflow kernel at depth 4, barrier kernel + packed gather at depth 6, in-run sharded-vs-single check)
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 240 -- 'bash tools/gpu_r02_pass4_2gpu.sh'
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
stamp() { echo "== $1 at +$(( $(date +%s) - T0 )) s" | tee -a $O/r02i_timeline.log; }
stamp bench2
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu 2>$O/r02i_bench2.err | grep '^{' > $O/r02i_bench_default_2gpu.json
python -c "
import json
d=json.loads(open('gpurun_out/r02i_bench_default_2gpu.json').readline())
print('default N=2:', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'])
for h in d.get('heom', []): print('  ', h['label'], h.get('value'), h.get('check'), h.get('error'), h.get('roofline',{}).get('kernel'))"
tail -5 $O/r02i_bench2.err
stamp tests2
timeout 60 python -m pytest tests/test_sharded.py -m gpu -q --timeout 50 -p no:cacheprovider -k "nccl_world2" > $O/r02i_pytest_sharded.log 2>&1
tail -3 $O/r02i_pytest_sharded.log
stamp done
