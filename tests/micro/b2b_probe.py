"""Per-step device times of back-to-back (no flush, no sync) training steps under torchrun -- debugging aid for the pipelined exchange."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from cropnerf_b200 import engine
from cropnerf_b200.rays import RayBundle

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
model = bench.build_model(dev, "mixed")
trainer = engine.Trainer(model, world_size=world, cuda_graph=True, force_proposal_update=True, ddp=os.environ.get("CNB_DDP", "auto"))
R = 4096
host = [bench.host_batch(R, seed=100 * rank + i) for i in range(4)]
res = [bench.to_bundle(b, dev, non_blocking=False) for b in host]
step = 2000
def one(step):
    rb, tg = res[step % 4]
    rb = RayBundle(rb.origins, rb.directions, rb.pixel_area, rb.camera_indices)
    return trainer.train_iteration(step, rb, tg)
for _ in range(5):
    one(step); step += 1
torch.cuda.synchronize(); dist.barrier()
n = 24
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
host_t = []
ev[0].record()
for i in range(n):
    t0 = time.perf_counter()
    one(step); step += 1
    host_t.append((time.perf_counter() - t0) * 1e3)
    ev[i + 1].record()
trainer.wait_deferred_update()
torch.cuda.synchronize()
# end-to-end flavour: pinned host batch in, loss read back every step; host time of the call vs the read
t_call, t_item = [], []
for i in range(n):
    b = host[step % 4]
    rb = RayBundle(b["origins"], b["directions"], None, b["camera_indices"]); tg = {"image": b["image"], "fruit_mask": b["fruit_mask"]}
    t0 = time.perf_counter()
    st_ = trainer.train_iteration(step, rb, tg); step += 1
    t1 = time.perf_counter()
    float(st_["loss"].item())
    t2 = time.perf_counter()
    t_call.append((t1 - t0) * 1e3); t_item.append((t2 - t1) * 1e3)
if rank == 0:
    print("e2e host ms in train_iteration:", [round(t, 3) for t in t_call[4:]])
    print("e2e host ms in loss.item():", [round(t, 3) for t in t_item[4:]])
    print("mode", trainer.ddp, "device ms per step:", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(n)])
    print("host ms per step:", [round(t, 3) for t in host_t])
dist.barrier(); dist.destroy_process_group()
