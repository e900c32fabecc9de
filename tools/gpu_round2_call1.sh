set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 120 python tools/fc_call_latency.py 2>&1 | tail -4
timeout 300 python tools/hm_latency.py 2>&1 | tail -8
timeout 300 python tools/hm_latency.py --no-fused 2>&1 | tail -8
