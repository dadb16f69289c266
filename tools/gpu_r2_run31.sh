#!/bin/bash
# round-2 GPU call 31: final single-GPU validation (smoke, whole GPU suite, default bench line) + profile captures
cd "$GRAFT_REPO_ROOT" || exit 1
bash tools/gpu_r2_run30_final.sh
bash tools/gpu_r2_run28_profiles.sh
