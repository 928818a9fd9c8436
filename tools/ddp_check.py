"""torchrun check of the multi-GPU paths on real GPUs (tests/test_gpu.py::test_two_rank_distributed_render launches it):
  * render_distributed, sample split and static tile split: reduced image == the reference's committed 64-spp frame
    (sample split: up to fp32 summation order; static tiles: bit for bit),
  * DistributedRenderer (one overlapped reduce per step; tile split with row bands that follow the measured batch times):
    after 8 steps of 8 spp the reduced SUM image / 64 == the same frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import cuda_pt_b200 as cpt
from cuda_pt_b200 import distributed as cd
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
world = dist.get_world_size()
scene = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref", "cbox_wfpt_256.cpts")
ref = np.fromfile(scene.replace(".cpts", "_wfpt_f0064.bin"), dtype=np.float32).reshape(256, 256, 4)


def report(tag, m, want_exact):
    rel = np.linalg.norm((m - ref)[..., :3]) / np.linalg.norm(ref[..., :3])
    exact = float((m[..., :3] == ref[..., :3]).all(-1).mean())
    print(f"[ddp_check] world={world} {tag}: rel_l2 vs the reference's 64-spp frame = {rel:.3e}, max abs = {np.abs(m - ref)[..., :3].max():.3e}, "
          f"bit-identical pixels = {exact * 100:.2f}%", flush=True)
    assert rel < 1e-6, tag
    if want_exact:
        assert exact == 1.0, tag


for split in ("samples", "tiles"):
    r = cpt.WavefrontRenderer(scene, lr, 0)
    mean = cd.render_distributed(r, 64, split=split)
    if rank == 0:
        report(f"render_distributed split={split}", mean.cpu().numpy(), split == "tiles")
    r.close()
for split in ("samples", "tiles"):
    r = cpt.WavefrontRenderer(scene, lr, 0)
    host = torch.empty((256, 256, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
    dr = cd.DistributedRenderer(r, split=split, dst=0, host_out=host)
    for _ in range(8):
        dr.step(8)
    dr.finish()
    if rank == 0:
        report(f"DistributedRenderer split={split} (8 steps x 8 spp, overlapped reduce" + (f", final row bands {dr.rows}" if split == "tiles" else "") + ")",
               (dr.reduced() / 64.0).cpu().numpy(), False)
        report(f"DistributedRenderer split={split}, host copy of the reduced frame", host.numpy() / 64.0, False)
    r.close()
dist.barrier()
if rank == 0:
    print("[ddp_check] OK", flush=True)
dist.destroy_process_group()
