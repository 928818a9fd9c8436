"""torchrun check of the sample-split multi-GPU path: reduced image == single-GPU image (fp32 summation order)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import cuda_pt_b200 as cpt
from cuda_pt_b200 import distributed as cd
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
scene = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref", "cbox_wfpt_256.cpts")
ref = np.fromfile(scene.replace(".cpts", "_wfpt_f0064.bin"), dtype=np.float32).reshape(256, 256, 4)
for split in ("samples", "tiles"):
    r = cpt.WavefrontRenderer(scene, lr, 0)
    mean = cd.render_distributed(r, 64, split=split)
    if rank == 0:
        m = mean.cpu().numpy()
        rel = np.linalg.norm((m - ref)[..., :3]) / np.linalg.norm(ref[..., :3])
        exact = float((m[..., :3] == ref[..., :3]).all(-1).mean())
        print(f"[ddp_check] world={dist.get_world_size()} split={split}: rel_l2 vs reference 64-spp frame = {rel:.3e}, "
              f"max abs = {np.abs(m-ref)[...,:3].max():.3e}, bit-identical pixels = {exact*100:.2f}%")
        assert rel < 1e-6
    r.close()
dist.destroy_process_group()
