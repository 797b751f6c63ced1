"""Where does the gap between the device-resident and the host-to-host step time come from?  (one GPU)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morig_b200  # noqa: E402
from morig_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
kw = synth.ARCH_KWARGS["jointnet_motion"]
model = morig_b200.jointnet_motion(**kw).eval()
model.load_state_dict(synth.seeded_state_dict(model, 1))
model = model.to(dev)
host = synth.make_batch(4, 4096, seed=0).pin_memory()
res = host.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream()


def total_ms(fn_loop, steps):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record(stream)
    fn_loop(steps)
    e.record(stream)
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps, 1e3 * (time.perf_counter() - t0) / steps


with torch.no_grad():
    for _ in range(4):
        model(res, res.pred_flow)

    def resident(steps, do_flush=True):
        for _ in range(steps):
            if do_flush:
                flush.fill_(1)
            model(res, res.pred_flow)
    print("resident, flush in the timed region  : %.3f ms/step (host %.3f)" % total_ms(resident, 30))
    print("resident, no flush                   : %.3f ms/step (host %.3f)" % total_ms(lambda n: resident(n, False), 30))

    for depth in (1, 2, 3):
        pipe = morig_b200.HostPipeline(model, depth=depth)

        def piped(steps, do_flush=True):
            for _ in range(steps):
                if do_flush:
                    flush.fill_(1)
                if pipe.in_flight == depth:
                    pipe.result()
                pipe.submit(host, host.pred_flow)
            while pipe.in_flight:
                pipe.result()
            pipe.join(stream)
        piped(3)
        print("pipeline depth %d, flush              : %.3f ms/step (host %.3f)" % ((depth,) + total_ms(piped, 30)))
        print("pipeline depth %d, no flush           : %.3f ms/step (host %.3f)" % ((depth,) + total_ms(lambda n: piped(n, False), 30)))

    # copies alone
    d = {k: torch.empty_like(v, device=dev) for k, v in host.__dict__.items() if torch.is_tensor(v)}
    outs = model(res, res.pred_flow)
    ho = [torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs]

    def h2d(steps):
        for _ in range(steps):
            for k, v in d.items():
                v.copy_(getattr(host, k), non_blocking=True)

    def d2h(steps):
        for _ in range(steps):
            for h, o in zip(ho, outs):
                h.copy_(o, non_blocking=True)
    print("H2D of one batch alone               : %.3f ms (host %.3f)" % total_ms(h2d, 20))
    print("D2H of one result alone              : %.3f ms (host %.3f)" % total_ms(d2h, 20))
