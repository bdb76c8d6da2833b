#!/bin/bash
mkdir -p gpurun_out
python scripts/topk_prof.py 151552 1000000 2>&1 | tail -3
python scripts/topk_prof.py 151552 125000 2>&1 | tail -3
