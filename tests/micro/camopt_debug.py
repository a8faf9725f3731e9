import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from helpers import product_bundle, product_model
from cropnerf_b200 import engine, synthetic
from cropnerf_b200.fruit_nerf import CameraOptimizer
from oracle import cases
dev = torch.device("cuda:0")
R, num_images = 256, 12
cfg = cases.make_config({}, small=True)
_, state = cases.build_oracle(cfg, num_images, 0, 0.5)
rays = synthetic.make_rays(R, seed=6, num_cameras=num_images)
targets = {k: v.to(dev) for k, v in synthetic.make_targets(R, seed=3).items()}
jit = synthetic.make_jitter(R, 3, seed=2)
g = torch.Generator().manual_seed(5)
pose0 = torch.randn((num_images, 6), generator=g) * 0.02
res = {}
for name, eager, overlap in (("eager", True, "1"), ("fused", False, "1"), ("fused2", False, "0"), ("fused3", False, "0"), ("eager2", True, "0")):
    model = product_model(cfg, state, num_images, dev, True, precision="fp32")
    model.camera_optimizer = CameraOptimizer(num_images, "SO3xR3").to(dev)
    with torch.no_grad():
        model.camera_optimizer.pose_adjustment.copy_(pose0.to(dev))
    feed = synthetic.JitterFeed(jit)
    model.proposal_sampler.initial_sampler.rand_fn = feed
    model.proposal_sampler.pdf_sampler.rand_fn = feed
    tr = engine.Trainer(model, force_proposal_update=True)
    tr._camopt_eager = eager
    grads = {}
    orig = tr.optimizer_step
    def spy(step, tr=tr, grads=grads, orig=orig, **kw):
        for n, grp in tr.groups.items():
            grads[n] = grp.grad.clone()
        orig(step, **kw)
    tr.optimizer_step = spy
    tr.train_iteration(2000, product_bundle(rays, dev), targets)
    res[name] = grads["camera_opt"][: num_images * 6].view(num_images, 6).cpu()
for k, v in res.items():
    print(k, v[:3])
print("eager vs fused", (res["eager"] - res["fused"]).norm() / res["eager"].norm())
for k in ("fused2","fused3","eager2"):
    print("eager vs", k, float((res["eager"] - res[k]).norm() / res["eager"].norm()))
print("diff rows", (res["eager"] - res["fused"])[:4])
