import sys, torch
sys.path.insert(0, '.')
from oracle import port
from tests.util import make_case
from mli_nerf_b200 import config
from mli_nerf_b200.model import Model
cu = lambda t: t.contiguous().cuda()
cfg = config.experiment("syn_hotdog_b", dict_size=14)
cfg.data.val.image_size = [40, 50]; cfg.model.render.rand_rays_val = 700
model = Model(cfg.model, cfg.data); case = make_case(R=8); model.load_state_dict(case["params"]); model = model.cuda()
pose = torch.tensor([[[1, 0, 0, 0.0], [0, -1, 0, 0.0], [0, 0, -1, 3.0]]], dtype=torch.float32)
intr = torch.tensor([[[60.0, 0, 25], [0, 60.0, 20], [0, 0, 1]]])
pose_light = torch.tensor([[[1, 0, 0, 1.0], [0, 1, 0, -2.0], [0, 0, 1, 3.0]]], dtype=torch.float32)
data = dict(pose=cu(pose), intr=cu(intr), pose_light=cu(pose_light), idx=torch.zeros(1).long())
out = model.inference(data)
c, ray, l = port.rays_from_pose(pose, intr, pose_light, (40, 50), torch.arange(2000)[None])
ref = port.render_rays(case["params"], case["ocfg"], c, torch.nn.functional.normalize(ray, dim=-1), l, rands=None, training=False, progress=1.0, keep=True)
dd = (out["dists"].cpu()[0, :, :, 0] - ref["dists"][0, :, :, 0]).abs().amax(dim=1)
print("dist diff quantiles", torch.quantile(dd, torch.tensor([0.5, 0.9, 0.99, 1.0])))
for k in ("rgb", "opacity", "o_r", "o_s", "gradient", "weights"):
    d = (out[k].cpu()[0] - ref[k][0].detach()).abs()
    d = d.reshape(2000, -1).amax(1)
    print(k, "max", float(d.max()), "n>2e-3", int((d > 2e-3).sum()), "corr with distdiff:", [ (float(dd[i]), float(d[i])) for i in d.topk(4).indices])
print("outside", int(ref["outside"].sum()), "opacity mean", float(ref["opacity"].mean()))
