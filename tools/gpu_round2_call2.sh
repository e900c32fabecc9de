cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_hm.py -m gpu -x -q 2>&1 | tail -5
timeout 120 python tools/fc_call_latency.py 2>&1 | tail -3
PNN_FC_STAMPS=1 timeout 120 python tools/fc_call_latency.py 2>&1 | tail -8
