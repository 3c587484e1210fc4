"""CPU test of the multi-GPU path's host logic: two gloo ranks shard the views round-robin, accumulate
per-Gaussian gradients into one flat buffer each, all-reduce, and must reproduce the single-process sum.
The per-view gradient producer is the CPU oracle (the CUDA path needs a GPU)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleRender(torch.autograd.Function):
    """semantics -> pseudo-loss through the CPU oracle (only dL/dsemantics and dL/dopacity are wired)."""

    @staticmethod
    def forward(ctx, semantics, opacity, g, cam, bg, w):
        import common
        from oracle import oracle
        arrs = common.gaussian_arrays(g)
        arrs["semantics"], arrs["opacities"] = semantics.detach().numpy(), opacity.detach().numpy()
        res = oracle.forward(**arrs, **common.cam_arrays(cam, bg))
        ctx.res, ctx.w = res, w
        L = (res.color * w["render"].numpy()).sum() + (res.semantics * w["semantics"].numpy()).sum()
        return torch.tensor(float(L))

    @staticmethod
    def backward(ctx, gout):
        from oracle import oracle
        w = ctx.w
        gr = oracle.backward(ctx.res, w["render"].numpy(), w["semantics"].numpy(), None, None, wide=True)
        return (torch.tensor(gr["dL_dsemantics"]) * gout, torch.tensor(gr["dL_dopacity"]) * gout, None, None, None, None)


def _scene():
    for p in (ROOT, os.path.join(ROOT, "goi-hyperplane_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from goi_b200.scenes import make_loss_weights, make_orbit_scene
    g, cams, bg = make_orbit_scene(150, 48, 32, 4, n_views=5, seed=31, px_sigma=3.0)
    ws = [make_loss_weights(4, 48, 32, 31 + v) for v in range(len(cams))]
    return g, cams, bg, ws


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, cams, bg, ws = _scene()
    from goi_b200 import view_parallel as vp
    sem = g.get_semantics.clone().requires_grad_(True)
    opa = g.get_opacity.clone().requires_grad_(True)
    views = list(range(len(cams)))
    flat, losses = vp.accumulate_views(views, [sem, opa], lambda v: _OracleRender.apply(sem, opa, g, cams[v], bg, ws[v]))
    assert vp.shard_views(len(views), rank, world) == views[rank::world]
    assert len(losses) == len(views[rank::world])
    assert sem.grad.data_ptr() == flat.flat.data_ptr()          # .grad aliases the flat all-reduce buffer
    if rank == 0:
        np.save(out_path, flat.flat.numpy())
    dist.destroy_process_group()


def test_two_rank_view_sharding_equals_single_process_sum(tmp_path):
    out = str(tmp_path / "flat.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    g, cams, bg, ws = _scene()
    from goi_b200 import view_parallel as vp
    sem = g.get_semantics.clone().requires_grad_(True)
    opa = g.get_opacity.clone().requires_grad_(True)
    flat, losses = vp.accumulate_views(list(range(len(cams))), [sem, opa],
                                       lambda v: _OracleRender.apply(sem, opa, g, cams[v], bg, ws[v]), rank=0, world=1)
    want = flat.flat.numpy()
    assert len(losses) == len(cams)
    assert np.abs(want).max() > 0
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


def test_shard_views_partition():
    from goi_b200.view_parallel import shard_views
    for world in (1, 2, 3, 8):
        allv = sorted(v for r in range(world) for v in shard_views(64, r, world))
        assert allv == list(range(64))
    assert shard_views(3, 5, 8) == []


def test_grad_arena_slots_are_aligned_and_only_aliased_grads_are_dropped():
    """GradArena hands 16-byte-aligned slices to the C ABI whatever P is (P % 4 != 0 put the rotation slot
    at an odd float offset), keeps its padding zero, and on entering the context drops only the `.grad`s that alias
    the arena (a gradient left by other code must survive)."""
    from goi_b200.view_parallel import GradArena
    P = 1001
    params = {"means3D": torch.zeros(P, 3, requires_grad=True), "opacities": torch.zeros(P, 1, requires_grad=True),
              "scales": torch.zeros(P, 3, requires_grad=True), "rotations": torch.zeros(P, 4, requires_grad=True)}
    arena = GradArena(params)
    base = arena.flat.data_ptr()
    for name, p in params.items():
        off = (arena.slots[name].data_ptr() - base) // 4
        assert off % 4 == 0 and arena.slots[name].numel() == p.numel(), name
    assert arena.flat.numel() % 4 == 0 and float(arena.flat.abs().sum()) == 0.0
    params["means3D"].grad = arena.slots["means3D"].view(P, 3)          # aliases the arena (left by the previous view)
    foreign = torch.ones(P, 1)
    params["opacities"].grad = foreign                                  # does not
    arena.clear_grads(only_aliased=True)
    assert params["means3D"].grad is None and params["opacities"].grad is foreign
    arena.clear_grads()
    assert params["opacities"].grad is None
