import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from multimodal_edema_prediction_b200 import ops
from multimodal_edema_prediction_b200.graph import CudaGraphStep
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
B = int(sys.argv[1]); L = int(sys.argv[2]); side = int(sys.argv[3])
bench.WORK.update(B=B, L=L)
model, flat, opt, red = bench.build(dev, 1)
hb = bench.synth_host_batch(B, 1)[1]
x = model.feats_to_input((hb["x_ts"], hb["x_static"], list(hb["bin_ends"])), B)
y = hb["y"].to(dev)
def fn():
    opt.zero_grad(); red.start_step()
    loss = model._supervised_loss(model.forward(x), y)
    loss.backward()
    opt.step(grad_scale=red.finish())
    return loss
try:
    if side:
        gs = CudaGraphStep(fn, {}, warmup=3)
        gs(); torch.cuda.synchronize()
        print("side-stream warmup + capture OK", float(gs.out))
    else:
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        g.replay(); torch.cuda.synchronize()
        print("main-stream warmup + capture OK", float(out))
except Exception:
    traceback.print_exc(limit=25)
