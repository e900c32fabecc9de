cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -8
timeout -s KILL 400 python hm/run_switch_multi.py 8 2560x1600 32 direct 2>&1 | tee gpurun_out/r2_config4_hm_switch_8gpu.json | tail -3
