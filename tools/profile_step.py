"""The command profiled under ncu for profiles/r02_*: N eager training steps of a bench.py workload (default c2: DuETT base,
B=256, bf16, fwd + loss + bwd + AdamW) between cudaProfilerStart/Stop, after 3 un-profiled warm-up steps.
usage: ncu --profile-from-start off ... python tools/profile_step.py [config] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = dict(bench.CONFIGS[key], key=key)
device = torch.device("cuda", 0)
torch.cuda.set_device(0)
wl = bench.Workload(cfg, device)
B = cfg["B"]
hb = bench.synth_host_batch(B, 1234, cfg["dims"], with_cxr=wl.task == "kd")
st = wl.static_from(hb, B)
for _ in range(3):
    wl.train_step(st)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(steps):
    loss = wl.train_step(st)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", steps, "steps of", key, "loss", float(loss.detach()))
