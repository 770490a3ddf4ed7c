#!/bin/bash
mkdir -p gpurun_out
export ENV_AB_SKIP_NEAREST=1
for cfg in "16 32" "16 64" "32 64" "64 128" "64 256"; do set -- $cfg; echo "front=$1 group=$2 $(TWG_ENV_FRONT=$1 TWG_ENV_GROUP=$2 python scripts/env_ab.py 2>&1 | tail -1 | cut -c1-260)"; done > gpurun_out/s16_front.log
cat gpurun_out/s16_front.log
