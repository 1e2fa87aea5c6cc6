#!/bin/bash
# Build liblgx.so in-tree (fail hard), then run a command on the B200 box:  scripts/gpu.sh [--timeout S] -- '<cmd>'
set -e
cd /root/repo
python -m factors_of_serendipity_recommendation_b200.build
exec /usr/local/graft/bin/gpurun "$@"
