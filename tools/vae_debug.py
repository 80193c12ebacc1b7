"""Debug aid: one VAE training step on the GPU beside the float64 restatement (prints per-array gradient errors)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from flexs_b200 import _native
from oracle import vae_oracle as vo

L, A, I, Z, B = 14, 4, 50, 2, 10
rng = np.random.default_rng(1)
ws = vo.init_weights(L, A, I, Z, seed=3)
idx = rng.integers(0, A, size=(B, L), dtype=np.uint8)
sw = rng.uniform(0.0, 2.0, size=B)
eps = rng.normal(size=(B, Z))
m1 = (rng.random((B, I)) >= 0.3) / 0.7
m2 = (rng.random((B, I)) >= 0.3) / 0.7
loss, grads, nm, nv = vo.loss_and_grads(ws, idx, A, sw, eps, m1, m2)
c = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
vae = _native.NativeVAE(L, A, I, Z)
vae.set_weights(ws)
t = [c(idx, np.uint8), c(sw, np.float32), c(m1, np.float32), c(m2, np.float32), c(eps, np.float32)]
got = vae.train_step_dev(t[0].data_ptr(), t[1].data_ptr(), B, t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr())
print("loss gpu", got, "oracle", loss)
for name, g, ref in zip(vo.NAMES, vae.get_gradients(), grads):
    sc = max(float(np.abs(ref).max()), 1e-12)
    print(f"{name:9s} max|ref| {sc:.3e}  max err/scale {np.abs(g - ref).max() / sc:.3e}")
# inference pieces
z = rng.normal(size=(3, Z))
out = torch.empty((3, L * A), dtype=torch.float32, device="cuda")
vae.set_weights(ws)
vae.decode_dev(c(z, np.float32).data_ptr(), 3, out.data_ptr())
torch.cuda.synchronize()
print("decode err", np.abs(out.cpu().numpy() - vo.decode(ws, z)).max())
lp = torch.empty(B, dtype=torch.float64, device="cuda")
vae.log_prob_dev(t[0].data_ptr(), B, 0, lp.data_ptr())
torch.cuda.synchronize()
print("logp", lp.cpu().numpy()[:4], vo.log_probability(ws, idx, A)[:4])
