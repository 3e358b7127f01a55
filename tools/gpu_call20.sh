#!/bin/bash
for k in 4 8 16; do echo "== PRN_WGRAD_MINKB=$k"; PRN_WGRAD_MINKB=$k timeout 300 python tools/wgrad_probe.py l0_1x1_64_256 l2_1x1_256_1024 l2_1x1_1024_256 l3_1x1_2048_512 l2_3x3_256 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('WGRAD_PROBE'):
        d=json.loads(l[len('WGRAD_PROBE '):]); print('  %-22s %7.1f us %7.1f TF plan %s' % (d['name'], d['us'], d['tflops'], d['plan(m_tiles,n_tiles,splits,kb/split,m_sub,stages,grid,atoms)']))
"; done
for k in 8 16; do PRN_WGRAD_MINKB=$k timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-inference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('minkb=$k value', d['value'], d['ms_per_step'], d['train_step']['fwd_ms'], d['train_step']['bwd_ms'])"; done
