#!/bin/bash
# per-phase cycle accounting of the banded kernels: rebuilds the library with EP_PHASE_TIMING=1 (tools only; run it on the GPU box)
EP_PHASE_TIMING=1 python -m eventpretrain_b200.build --force > /dev/null 2>&1
EP_PRINT_TIMING=1 python tools/quick_bin.py --batch 64 --methods banded --steps 1 "$@" 2>&1 | tail -12
