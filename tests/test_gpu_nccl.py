"""SURVEY.md section 8e parity check ON HARDWARE: N NCCL ranks (one per GPU) shard the views round-robin, every rank
accumulates its views' gradients in place in its GradArena, one ncclAllReduce(SUM) -- and the reduced buffer must
equal the SUM OVER ALL VIEWS of the reference kernels' single-view gradients (oracle/_ref) within the gradient
tolerance.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a single-GPU box (the gloo twin of the host logic is
tests/test_view_parallel.py)."""
import math
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_VIEWS, P, W, H, S, SEED = 6, 60_000, 480, 320, 16, 23
SLOTS = ("means3D", "opacities", "scales", "rotations", "sh", "semantics")
REF_KEYS = dict(means3D="dL_dmeans3D", opacities="dL_dopacity", scales="dL_dscales", rotations="dL_drotations",
                sh="dL_dsh", semantics="dL_dsemantics")


def _paths():
    for p in (ROOT, os.path.join(ROOT, "goi-hyperplane_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene(dev):
    _paths()
    from goi_b200.scenes import make_loss_weights, make_orbit_scene
    g, cams, bg = make_orbit_scene(P, W, H, S, N_VIEWS, SEED, px_sigma=2.5)
    ws = [make_loss_weights(S, W, H, SEED + v, device=dev) for v in range(N_VIEWS)]
    return g.to(dev), [c.to(dev) for c in cams], bg.to(dev), ws


def _worker(rank, world, port, out_path):
    _paths()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gaussian_renderer import render
    from goi_b200 import view_parallel as vp
    from goi_b200.scenes import PipeFlags
    g, cams, bg, ws = _scene(dev)
    g = g.requires_grad_(True)
    arena = vp.GradArena({"means3D": g.get_xyz, "opacities": g.get_opacity, "scales": g.get_scaling,
                          "rotations": g.get_rotation, "sh": g.get_features, "semantics": g.get_semantics})
    arena.flat.fill_(float("nan"))                    # the first view must overwrite, not add
    outs = ("render", "semantics", "depth", "alpha")
    mine = vp.shard_views(N_VIEWS, rank, world)
    for i, v in enumerate(mine):
        out = render(cams[v], g, PipeFlags(), bg)
        with arena.accumulating(i > 0):
            torch.autograd.backward([out[k] for k in outs], [ws[v][k] for k in outs])
    arena.all_reduce()                                 # ncclAllReduce(SUM, f32) on the flat buffer
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({k: arena.slots[k].cpu() for k in SLOTS}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_all_reduced_arena_equals_sum_of_reference_view_gradients(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _paths()
    import common
    from oracle import refshim
    out = str(tmp_path / "arena.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    g, cams, bg, ws = _scene("cuda:0")
    want = None
    for v in range(N_VIEWS):                           # every view, one at a time, through the reference's kernels
        if refshim.available(S):
            gr = common.run_reference_cuda(g, cams[v], bg, ws[v])["grads"]
        else:                                          # (no prebuilt reference: this build's own single-view path)
            gr = common.run_cuda(g, cams[v], bg, ws[v])["grads"]
        gr = {k: gr[REF_KEYS[k]].detach().double().cpu().reshape(-1) for k in SLOTS}
        want = gr if want is None else {k: want[k] + gr[k] for k in SLOTS}
    for k in SLOTS:
        scale = float(want[k].abs().max())
        err = float((got[k].double() - want[k]).abs().max())
        assert scale > 0 and math.isfinite(err)
        assert err <= common.GRAD_RTOL * scale, f"{k}: all-reduced sum differs by {err / scale:.2e} of max"
