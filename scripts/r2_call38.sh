#!/bin/bash
# round 2, GPU call 38: does FP64 pipe time hide the issue slots of other instruction classes? (scripts/microbench/fp64_issue_overlap.cu)
set -u
mkdir -p gpurun_out
timeout 60 scripts/microbench/fp64_issue_overlap | tee gpurun_out/r2_fp64_issue_overlap.json
