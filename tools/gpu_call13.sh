#!/bin/bash
for m in none fields all; do timeout 200 python tools/infer_e2e.py $m 2>&1 | grep INFER_E2E; done
PRN_CONV_TMA=0 timeout 200 python tools/infer_e2e.py all 2>&1 | grep INFER_E2E
PRN_PDL=0 timeout 200 python tools/infer_e2e.py all 2>&1 | grep INFER_E2E
timeout 200 python tools/infer_e2e.py all 2 2>&1 | grep INFER_E2E
timeout 300 python -m pytest tests/test_pointwise_gpu.py -x -q 2>&1 | tail -3
