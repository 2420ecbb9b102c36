"""host-side timing of one config-5 step (debug helper; torchrun --nproc-per-node 2)"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easygaussiansplatting_b200.gsfunction import Camera
from easygaussiansplatting_b200.parallel import MultiViewStep
from easygaussiansplatting_b200.scene import ring_camera, synthetic_scene, upstream_gradient
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N5, V5, W, H = 2_000_000, 8, 1920, 1080
sc = synthetic_scene(N5, W, H, sh_dim=48, seed=1)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
mv = MultiViewStep(T(sc["pws"]), T(sc["rots"]), T(sc["scales"]), T(sc["shs"]), T(sc["alphas"][:, None]))
cams = []
for v in [v for v in range(V5) if v % world == rank]:
    Rcw, tcw, twc = ring_camera(v, V5)
    cams.append(Camera(W, H, sc["fx"], sc["fy"], sc["cx"], sc["cy"], T(Rcw), T(tcw), T(twc)))
dl = T(upstream_gradient(W, H, rank) * (3.0 * W * H))
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    marks = []
    for cam in cams:
        image, ctx = mv.render(cam); marks.append(("render", time.perf_counter()))
        mv.backward(ctx, dl); marks.append(("backward", time.perf_counter()))
    g = mv.reduce(); marks.append(("reduce_host", time.perf_counter()))
    torch.cuda.synchronize(); marks.append(("sync", time.perf_counter()))
    if rank == 0:
        prev = t0; s = []
        for n, t in marks:
            s.append("%s %.2f" % (n, (t - prev) * 1e3)); prev = t
        print("it", it, "total %.2f ms |" % ((marks[-1][1] - t0) * 1e3), " ".join(s), flush=True)
if world > 1:
    dist.destroy_process_group()
