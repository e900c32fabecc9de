"""BASELINE.json configs[4]: HM-16.15 switch variant, independent synthetic 2560x1600 frames, one process per GPU
(replicas only: no collective).  Prints the wall time of encoding N frames on N GPUs at once next to one frame alone."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w, h = (int(v) for v in (sys.argv[2].split('x') if len(sys.argv) > 2 else ('2560', '1600')))
qp = sys.argv[3] if len(sys.argv) > 3 else '32'
backend = sys.argv[4] if len(sys.argv) > 4 else 'direct'


def launch(gpu, seed):
    # one process per GPU, pinned as SURVEY.md section 8(d) prescribes (CUDA_VISIBLE_DEVICES; the library then sees device 0)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(gpu), PNN_DEVICE='0')
    return subprocess.Popen([sys.executable, os.path.join(ROOT, 'hm', 'run_hm.py'), '--variant', 'switch', '--width', str(w),
                             '--height', str(h), '--qps', qp, '--seed', str(seed), '--backend', backend], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                            text=True, env=env)


t0 = time.time()
p = launch(0, 0)
out1 = p.communicate()[0]
t_one = time.time() - t0
t0 = time.time()
procs = [launch(g, g) for g in range(n)]
outs = [q.communicate()[0] for q in procs]
t_all = time.time() - t0
rows = [json.loads(o.strip().split('\n')[-1]) for o in [out1] + outs if o.strip()]
print(json.dumps({'config': 'configs[4]: HM-16.15 switch, %d independent synthetic %dx%d frames, one process per GPU, QP %s' % (n, w, h, qp),
                  'wall_one_frame_one_gpu_s': t_one, 'wall_%d_frames_%d_gpus_s' % (n, n): t_all,
                  'backend': backend, 'all_hash_ok': all(r.get('decoder_hash_ok') for r in rows), 'all_recon_equal': all(r.get('recon_enc_equals_dec') for r in rows),
                  'encoder_wall_s': [r.get('encoder_wall_s') for r in rows], 'pnn_encoder_total': [r.get('pnn_encoder', [''])[-1] for r in rows]}))
