#!/bin/bash
mkdir -p gpurun_out
export ENV_AB_SKIP_NEAREST=1
for b in 30 24 21 18; do TWG_SORT_BITS=$b python scripts/env_ab.py 2>&1 | tail -1; done > gpurun_out/s14_sortbits.log
cat gpurun_out/s14_sortbits.log
for b in 30 24 18; do TWG_SORT_BITS=$b python scripts/prof_part.py winding 4e6 3 2>&1 | tail -1; done
