import os, sys, time, cProfile, pstats, io
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from cropnerf_b200.rays import RayBundle
dev = torch.device("cuda:0")
model = bench.build_model(dev, "mixed").eval()
Rr = 32768
rhost = bench.host_batch(Rr, seed=999)
hb = RayBundle(rhost["origins"], rhost["directions"], rhost["pixel_area"], rhost["camera_indices"])
for _ in range(5):
    model.get_outputs_for_camera_jagged_ray_bundle(hb)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    model.get_outputs_for_camera_jagged_ray_bundle(hb)
torch.cuda.synchronize()
print("e2e ms per call", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    model.get_outputs_for_camera_jagged_ray_bundle(hb)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
