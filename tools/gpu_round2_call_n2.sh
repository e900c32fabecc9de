cd $GRAFT_REPO_ROOT
timeout -s KILL 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_strong.json 2> gpurun_out/r2_bench_n1_strong.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_strong.json')); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['scaling'], d['clocks'])"
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_strong.json 2> gpurun_out/r2_bench_n2_strong.err
tail -3 gpurun_out/r2_bench_n2_strong.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_strong.json')); print('N=2 strong', d['value'], d['ms_per_step'], d['e2e']['value'], d['scaling'], d['e2e']['mean_psnr_pnn'], d['e2e']['frequency_win_pnn'])"
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --scaling weak > gpurun_out/r2_bench_n2_weak.json 2> gpurun_out/r2_bench_n2_weak.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_weak.json')); print('N=2 weak', d['value'], d['ms_per_step'], d['e2e']['value'], d['scaling'])"
