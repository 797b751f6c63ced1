import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import morig_b200
from morig_b200 import synth
dev = torch.device("cuda:0")
kw = synth.ARCH_KWARGS["jointnet_motion"]
model = morig_b200.jointnet_motion(**kw).eval()
model.load_state_dict(synth.seeded_state_dict(model, 1)); model = model.to(dev)
for b, n in ((1, 1024), (1, 4096), (4, 4096)):
    data = synth.make_batch(b, n, seed=0).to(dev)
    for graph in (True, False):
        model.use_cuda_graph = graph
        with torch.no_grad():
            for _ in range(4): model(data, data.pred_flow)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(20): model(data, data.pred_flow)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print(f"jointnet {b} x {n}: {'graph replay' if graph else 'eager launches'} {dt*1e3:.3f} ms/forward (wall)")
