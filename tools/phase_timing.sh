#!/bin/bash
# per-phase cycle accounting of the banded kernels (library must be built with EP_PHASE_TIMING=1)
EP_PROFILE_METHOD=banded EP_PRINT_TIMING=1 EP_PROFILE_BATCH=64 python tools/profile_binning.py 2>&1 | tail -12
