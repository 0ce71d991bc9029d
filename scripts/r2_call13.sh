#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/sanitizer_target.py > gpurun_out/r2_sanitizer_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/r2_sanitizer_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitizer_target.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/r2_sanitizer_memcheck.log
