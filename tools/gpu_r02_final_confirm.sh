#!/bin/bash
# Round-2 final confirmation (1 GPU): default bench line of the final defaults, then the GPU tests (HEOM first)
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
stamp() { echo "== $1 at +$(( $(date +%s) - T0 )) s" | tee -a $O/r02j_timeline.log; }
stamp bench_default
timeout 120 python bench.py 2>$O/r02j_bench.err | grep '^{' > $O/r02j_bench_default_1gpu.json
python -c "
import json
d=json.loads(open('gpurun_out/r02j_bench_default_1gpu.json').readline())
print('default:', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['check'].get('oracle_spot_check_relerr'), d['cpu_baseline']['value'])
for h in d.get('heom', []): print('  ', h['label'], h.get('value'), h.get('roofline', {}).get('frac'), h.get('check'), h.get('error'))"
stamp tests_heom
timeout 100 python -m pytest tests -m gpu -q --timeout 60 -p no:cacheprovider -k "heom" > $O/r02j_pytest_heom.log 2>&1
tail -2 $O/r02j_pytest_heom.log
stamp tests_rest
timeout 100 python -m pytest tests -m gpu -q --timeout 60 -p no:cacheprovider -k "not heom" > $O/r02j_pytest_rest.log 2>&1
tail -2 $O/r02j_pytest_rest.log
stamp done
