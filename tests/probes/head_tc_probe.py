import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import wakeword_jupyterlab_b200 as ww
from oracle import recipe as R
g = np.load('tests/golden/model_trained.npz')
sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
def load(mode):
    net = ww.WakewordModel(ww.ModelConfig, ww.AudioConfig).cuda().eval()
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    net.conv_mode = mode
    return net
rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
net = load("fp32")
base = R.make_clips(int(g["n"]), seed=int(g["clip_seed"]))
for n in (int(g["n"]), 1, 129, 300):
    clips = np.tile(base, ((n + len(base) - 1) // len(base), 1))[:n]
    x = torch.from_numpy(clips).cuda()
    os.environ.pop("WW_HEAD_KERNEL", None)
    lt, pt, dt = ww.score_clips(x, net, normalize=True)
    os.environ["WW_HEAD_KERNEL"] = "fp32"
    lf, pf, df = ww.score_clips(x, net, normalize=True)
    os.environ.pop("WW_HEAD_KERNEL")
    print(n, "tc vs fp32", rel(lt.cpu().numpy(), lf.cpu().numpy()), "prob", float((pt - pf).abs().max()), "dec eq", bool(torch.equal(dt, df)))
    if n == int(g["n"]):
        print("  tc vs ref", rel(lt.cpu().numpy(), g["logits_reference"]), "fp32 vs ref", rel(lf.cpu().numpy(), g["logits_reference"]),
              "tc vs f64", rel(lt.cpu().numpy(), g["logits_f64"]) if "logits_f64" in g.files else None,
              "fp32 vs f64", rel(lf.cpu().numpy(), g["logits_f64"]) if "logits_f64" in g.files else None)
