import os, sys, torch, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests import helpers as H
from dynrf_b200 import ops, backward as B, lib as L
from dynrf_b200.ops import _p, _stream
torch.set_printoptions(linewidth=200, precision=5, sci_mode=False)
case = H.build_pair([21, 24, 18], 96, 40, ray_type="ndc", static_mode="MLP_Fea", seed=7 * 96 + 40)
static, dyn = H.native_models(case, "cuda:0")
rays, ts = case["rays"].cuda(), case["ts"].cuda()
cfg = dyn.native_cfg("ndc")
zvec = ops.make_zvec(cfg, 40, "ndc", False, rays.device)
xyz, z, valid = ops.sample_points(cfg, rays, zvec)
r = ops.dynamic_forward_raw(dyn, rays, ts, xyz, z, valid, "ndc", save_feat=True)
n, s = 96, 40
cnt = int(r["cnt"]); idx = r["idx"][:cnt].long()
print("active", cnt)
g = torch.Generator().manual_seed(1); grgb = torch.randn(n, s, 3, generator=g).cuda()
st = B._HeadBwdState(dyn).refresh(dyn)
ld = r["ldf"]; b = st.scratch(ld, rays.device)
for v in b.values(): v.fill_(float("nan"))
gw3v = torch.zeros(3, 3, device="cuda"); gb3 = torch.zeros(3, device="cuda")
lib = L.load()
L.check(lib.dynrf_dyn_head_bwd_tc(C.byref(cfg), _p(st.wq), _p(st.cst), _p(r["rays"]), _p(r["xyz"]), _p(r["tfeat"]), _p(r["featT"]), _p(grgb), _p(r["idx"]), _p(r["cnt"]), n, s, ld, _p(b["X1T"]), _p(b["H1T"]), _p(b["H2T"]), _p(b["DH2T"]), _p(b["DH1T"]), _p(b["DOT"]), _p(b["DFT"]), _p(gw3v), _p(gb3), _stream()), "head")
torch.cuda.synchronize()
# reference
rm = dyn.renderModule; l1, l2, l3 = rm.mlp[0], rm.mlp[2], rm.mlp_view[0]
feat = r["featT"][:27, :cnt].t()
xn = dyn.normalize_coord(xyz.reshape(-1, 3)[idx])
tt = ts[idx // s][:, None]
from models.tensorBase import positional_encoding as PE
X = torch.cat([feat, xn, PE(xn, 10), tt, PE(tt, 8)], -1)
vd = torch.nn.functional.normalize(rays[idx // s, 3:6], dim=-1)
with torch.no_grad():
    h1 = torch.relu(l1(X)); h2 = torch.relu(l2(h1)); o = l3(torch.cat([h2, vd], -1)); rgb = torch.sigmoid(o)
    do = grgb.reshape(-1, 3)[idx] * rgb * (1 - rgb)
    dh2 = (do @ l3.weight[:, :128]) * (h2 > 0); dh1 = (dh2 @ l2.weight) * (h1 > 0); df = dh1 @ l1.weight[:, :27]
def cmp(name, got, ref):
    d = (got - ref).abs().max(); print(f"{name:6s} maxdiff {float(d):.3e} refmax {float(ref.abs().max()):.3e} nan {int(torch.isnan(got).sum())}")
cmp("H1T", b["H1T"][:, :cnt].t(), h1); cmp("H2T", b["H2T"][:, :cnt].t(), h2); cmp("DOT", b["DOT"][:3, :cnt].t(), do)
cmp("DH2T", b["DH2T"][:, :cnt].t(), dh2); cmp("DH1T", b["DH1T"][:, :cnt].t(), dh1); cmp("DFT", b["DFT"][:27, :cnt].t(), df)
cm = st.colmap.long(); rows = torch.nonzero(cm >= 0).reshape(-1)
Xmine = torch.zeros(cnt, 107, device="cuda"); Xmine[:, cm[rows]] = b["X1T"][rows, :cnt].t()
cmp("X1T", Xmine, X); print("ones row", float(b["X1T"][112, :cnt].min()), float(b["X1T"][112, :cnt].max()))
print("rgb fwd vs recompute", float((r["rgb"].reshape(-1, 3)[idx] - rgb).abs().max()))
