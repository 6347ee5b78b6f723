import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests import helpers as H
from tests.test_gpu_grad import _oracle_full, _native_full, _rel, _grad_ok, _loss_weights
from oracle import rodynrf_oracle as O
from dynrf_b200 import backward as B
from dynrf_b200.autograd import _split_ray_out
torch.set_printoptions(linewidth=200, precision=6, sci_mode=False)
g = torch.Generator().manual_seed(3)
n, s = 37, 150
rgb_s, rgb_d = torch.rand(n, s, 3, generator=g), torch.rand(n, s, 3, generator=g)
sig_s, sig_d = torch.relu(torch.randn(n, s, generator=g)) * 2, torch.relu(torch.randn(n, s, generator=g)) * 2
dists = torch.rand(n, s, generator=g) * 0.3
blend = torch.rand(n, s, generator=g)
z = torch.sort(torch.rand(n, s, generator=g), -1).values
rays, _ = H.synth_rays(n, "ndc", g)
ins = [rgb_s, sig_s, rgb_d, sig_d, dists, blend, z, rays]
o = O.raw2outputs(*ins[:7], ins[7], is_train=True, ray_type="ndc", white_bg_draw=False)
b = [t.clone().cuda() for t in ins]
oo, wf, wsn, wd = B.CompositeFn.apply(*b, False, "ndc")
rf, df, af, rs, ds, as_, rd, dd, ad, dyn = _split_ray_out(oo)
o2 = (rf, df, af, wf, rs, ds, as_, wsn, rd, dd, ad, wd, dyn)
names = ["rgb_full", "depth_full", "acc_full", "w_full", "rgb_s", "depth_s", "acc_s", "w_s", "rgb_d", "depth_d", "acc_d", "w_d", "dyn"]
for nme, x, y in zip(names, o, o2):
    d = (y.cpu() - x).abs()
    print(nme, "maxabs", float(d.max()), "ref range", float(x.min()), float(x.max()))
print("depth_full ref", o[1][:6], "got", o2[1][:6].cpu())
print("acc_full ref", o[2][:6], "got", o2[2][:6].cpu())

case = H.build_pair([21, 24, 18], 96, 40, ray_type="ndc", static_mode="MLP_Fea", seed=7 * 96 + 40)
ss, sd, _, fs, fd, o = _oracle_full(case)
shapes = [t.shape for t in o] + [fd[5].shape, fd[2].shape, fs[4].shape]
ws = _loss_weights(shapes, 99)
sum((t * w).sum() for t, w in zip(list(o) + [fd[5], fd[2], fs[4]], ws)).backward()
static, dyn, _, fs2, fd2, o2 = _native_full(case)
sum((t * w.cuda()).sum() for t, w in zip(list(o2) + [fd2[5], fd2[2], fs2[4]], ws)).backward()
for tag, mod, sdict in (("s", static, ss), ("d", dyn, sd)):
    for name, p in mod.named_parameters():
        if name.startswith("scene_flow"): continue
        r = _grad_ok(p.grad.cpu(), sdict[name].grad)
        if not r[0] or r[1] > 1e-3: print(tag, name, r, tuple(p.shape))
