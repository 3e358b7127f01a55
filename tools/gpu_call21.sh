#!/bin/bash
timeout 900 python -m pytest tests/test_train_ops_gpu.py tests/test_train_model_gpu.py -q -x 2>&1 | tail -4
timeout 300 python tools/wgrad_probe.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('WGRAD_PROBE'):
        d=json.loads(l[len('WGRAD_PROBE '):]); print('  %-22s %7.1f us %7.1f TF plan %s' % (d['name'], d['us'], d['tflops'], d['plan(m_tiles,n_tiles,splits,kb/split,m_sub,stages,grid,atoms)']))
"
timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-inference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], d['ms_per_step'], d['train_step']['fwd_ms'], d['train_step']['bwd_ms'])"
