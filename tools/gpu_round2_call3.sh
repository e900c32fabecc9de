cd $GRAFT_REPO_ROOT
timeout -s KILL 200 python -m pytest tests/test_gpu_hm.py -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 1300 python hm/config4.py --out gpurun_out/r2_config3_hm.json 2>&1 | tail -40
