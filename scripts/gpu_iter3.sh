#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-iter}
if [ -n "${TESTS:-}" ]; then timeout 1200 python -m pytest $TESTS -q -m gpu -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/${tag}_tests.log; fi
if [ -n "${BENCH:-}" ]; then timeout 900 python bench.py $BENCH > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d.get('e2e_parity'))
    print('roofline frac', d['roofline']['frac'], 'traffic', d['roofline'].get('traffic'))
    for k,v in (d.get('extra') or {}).items(): print(k, json.dumps(v)[:900])
except Exception as e: print('parse error', e)
PY
fi
