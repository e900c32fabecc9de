cd $GRAFT_REPO_ROOT
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
( echo "== tools/fc_call_latency.py (pnn_predict_hm_context from ctypes)"; PNN_FC_STAMPS=1 timeout -s KILL 100 python tools/fc_call_latency.py 2>&1 | tail -6
  echo "== tools/hm_latency.py (pnn_set_context + pnn_predict_hm from Python)"; timeout -s KILL 100 python tools/hm_latency.py 2>&1 | tail -6
  echo "== tools/hm_latency.py --no-fused (plain CUDA graphs)"; timeout -s KILL 100 python tools/hm_latency.py --no-fused 2>&1 | tail -6
  echo "== tools/ref_backend_latency.py (CPU baseline backend)"; timeout -s KILL 200 python tools/ref_backend_latency.py 2>&1 | tail -8 ) > gpurun_out/r2_inloop_latency.txt 2>&1
cat gpurun_out/r2_inloop_latency.txt
timeout -s KILL 300 python bench.py --config conv64 --steps 3 --warmup 3 > gpurun_out/r2_bench_conv64.json 2> gpurun_out/r2_bench_conv64.err; tail -c 1500 gpurun_out/r2_bench_conv64.json; tail -3 gpurun_out/r2_bench_conv64.err
