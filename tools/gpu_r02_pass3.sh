#!/bin/bash
# Round-2 third GPU pass (1 GPU): test suite, then the HEOM stage-kernel modes after the coefficient-table re-layout
# (0 table walk, 1 packed gather in the generic tile code, 2 heom_stage_fast_kernel + global table, 3 + n_k x base from
# shared memory) on the 38 760-ADO hierarchy and the batch of 64 FMO hierarchies, and one ncu capture of the fastest.
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/gpu_r02_pass3.sh'
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
stamp() { echo "== $1 at +$(( $(date +%s) - T0 )) s" | tee -a $O/r02h_timeline.log; }
line() { grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); d['tag']='$1'; print(json.dumps(d))"; }

stamp tests
timeout 300 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider > $O/r02h_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/r02h_pytest.log
tail -4 $O/r02h_pytest.log
grep -E "^(FAILED|ERROR)" $O/r02h_pytest.log | head

stamp heom_modes
: > $O/r02h_heom_modes.jsonl
hrun() {
  local tag=$1; shift
  timeout 120 python bench.py --workload heom_fmo "$@" --steps 3 --warmup 3 --no-cpu 2>>$O/r02h_heom.err | line "$tag" >> $O/r02h_heom_modes.jsonl
}
for cfg in b64 d6; do
  if [ $cfg = d6 ]; then A="--depth 6 --batch 1 --rk-steps 50"; else A="--depth 4 --batch 64 --rk-steps 8"; fi
  for m in 0 1 2 3; do LIMEB200_HEOM_STAGE_MODE=$m hrun ${cfg}_mode$m $A; done
done
python - <<'E' | tee -a gpurun_out/r02h_timeline.log
import json
for l in open('gpurun_out/r02h_heom_modes.jsonl'):
    d = json.loads(l)
    print(d['tag'], '%.4g' % d['value'], '%.4f' % d['roofline']['frac'], d.get('check'))
E
MODE=$(python - <<'E'
import json
best, bv = 0, 0.0
for l in open('gpurun_out/r02h_heom_modes.jsonl'):
    d = json.loads(l)
    if d['tag'].startswith('b64_mode') and d['value'] > bv:
        best, bv = int(d['tag'][-1]), d['value']
print(best)
E
)
echo "fastest stage mode (batch 64): $MODE" | tee -a $O/r02h_timeline.log

if [ "$MODE" != "0" ]; then
  stamp ncu_heom
  LIMEB200_HEOM_STAGE_MODE=$MODE timeout 150 ncu --set full --clock-control none --import-source on -k regex:heom_stage -s 8 -c 1 -f -o $O/r02h_heom_stage_b64 \
    python bench.py --workload heom_fmo --depth 4 --batch 64 --rk-steps 4 --steps 1 --warmup 0 --no-cpu > $O/r02h_ncu_heom.log 2>&1
  python tools/ncu_summary.py $O/r02h_heom_stage_b64.ncu-rep > $O/r02h_heom_stage_b64_summary.txt 2>&1
  head -32 $O/r02h_heom_stage_b64_summary.txt
fi
stamp done
