#!/bin/bash
# last check of a build: GPU parity suite, smoke, top-k slice + full timing of the product library
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_test_gpu.log 2>&1; echo "exit $?"; tail -4 gpurun_out/r2_test_gpu.log | cut -c1-300
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo -n "== slice: "; timeout 300 python bench.py --topk-only --topk 151552x1000000x128x100 --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'parity', t['parity_check']['ok'])"
echo -n "== full: "; timeout 600 python bench.py --topk-only --topk-steps 3 2>/dev/null | python -c "import json,sys; t=json.load(sys.stdin); print('ms', round(t['ms_per_step'],2), 'frac', t['roofline']['frac'], 'parity', t['parity_check']['ok'])"
