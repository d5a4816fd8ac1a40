"""Diagnostic for tests/test_gpu_multi.py: run its worker on 1 and 2 GPUs and print the differences it asserts on."""
import os, pickle, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_multi as T
tmp = tempfile.mkdtemp()
res = {}
for world in (1, 2):
    out = os.path.join(tmp, "w%d.pkl" % world)
    script = os.path.join(tmp, "worker%d.py" % world)
    open(script, "w").write(T.WORKER % dict(root=ROOT, out=out))
    cmd = [sys.executable, script] if world == 1 else [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29655", script]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print("world", world, "rc", p.returncode, p.stderr[-800:] if p.returncode else "")
    res[world] = pickle.load(open(out, "rb"))
a1, a2 = res[1]['a'], res[2]['a']
print("a n_iter equal", np.array_equal(a1['n_iter'], a2['n_iter']), a1['n_iter'].tolist(), a2['n_iter'].tolist())
print("a Psi", np.abs(a1['Psi'] - a2['Psi']).max(), "gain", np.abs(a1['gain'] - a2['gain']).max(),
      "fdr calls equal", np.array_equal(a1['fdr'] < 0.05, a2['fdr'] < 0.05), "losses shape", a1['losses'].shape, a2['losses'].shape)
if a1['losses'].shape == a2['losses'].shape:
    print("a losses rel", np.abs(a1['losses'] - a2['losses']).max() / np.abs(a1['losses']).max())
b1, b2 = res[1]['b'], res[2]['b']
print("b losses shape", b1['losses'].shape, b2['losses'].shape)
if b1['losses'].shape == b2['losses'].shape:
    print("b losses rel", np.abs(b1['losses'] - b2['losses']).max() / np.abs(b1['losses']).max())
for k in ('Psi', 'gc', 'ic', 'sg', 'cc'):
    d = np.abs(b1[k] - b2[k])
    print("b", k, "median %.3g q99 %.3g max %.3g" % (np.median(d), np.quantile(d, 0.99), d.max()))
print("b lg rel", np.abs(b1['lg'] - b2['lg']).max() / np.abs(b1['lg']).max())
