import sys, torch
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,R+'/goi-hyperplane_b200')
from gaussian_renderer import render
from goi_b200.scenes import PipeFlags, make_loss_weights, make_scene
from goi_b200 import view_parallel as vp
g, cam, bg = make_scene(20000, 320, 200, 16, 1)
g = g.to('cuda').requires_grad_(True); cam=cam.to('cuda'); bg=bg.cuda()
w = make_loss_weights(16, 320, 200, 1, device='cuda')
arena = vp.GradArena({"means3D": g.get_xyz, "opacities": g.get_opacity, "scales": g.get_scaling, "rotations": g.get_rotation, "sh": g.get_features, "semantics": g.get_semantics})
outs=("render","semantics","depth","alpha")
out = render(cam, g, PipeFlags(), bg)
with arena:
    torch.autograd.backward([out[k] for k in outs],[w[k] for k in outs])
print("alias:", {k: (p.grad.data_ptr()==arena.slots[k].data_ptr()) for k,p in arena.named.items()})
print(float(arena.flat.abs().sum()), float(sum(p.grad.abs().sum() for p in arena.named.values())))
