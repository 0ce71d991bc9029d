#!/bin/bash
# round 2, GPU call 35 (last GPU minutes of the round): whole -m gpu suite with the new tests -- golden loops of the example scripts with
# sliding constraints (02 / 06 / 08), the README script through the host mirror at 512x256x256 (its timings go to a json)
set -u
mkdir -p gpurun_out
O=gpurun_out
export ESIMP_HOST_MIRROR_OUT=$O/r2_host_mirror_c5.json
timeout 450 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -40 | tee $O/r2_call35_gpu_suite.log
