"""Experiment: software-pipelining the step across streams.  The fused kernel is dispatch/latency bound (HBM at
55 %), the SAD-of-Sobel scan is LSU bound and soft-argmin is HBM-read bound: with two steps in flight on two
streams (own workspaces and outputs) the block scheduler can put CTAs of different kernels on one SM.
usage: python profiles/overlap_streams.py [N pairs] [steps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__
__graft_entry__.build()
from msnets_b200 import cbmv, regression
from tests._synth import bordered_pair

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
H, W, D, B = 540, 960, 192, 10
pairs = [bordered_pair(H, W, 1234 + i, border=B) for i in range(N)]
l = torch.stack([torch.from_numpy(p[0]) for p in pairs]).cuda()
r = torch.stack([torch.from_numpy(p[1]) for p in pairs]).cuda()
exs = [cbmv.MSFeatureExtractor(N, H + 2 * B, W + 2 * B, maxdisp=D, board_h=B, board_w_left=B, board_w_right=B) for _ in range(2)]
outs = [e.empty_output() for e in exs]
logits = [torch.randn(N, D, H, W, device="cuda") for _ in range(2)]
disp = [torch.empty(N, H, W, device="cuda") for _ in range(2)]
res = {}


def timed(name, fn, drain):
    for i in range(4):
        fn(i)
    drain()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        fn(i)
    drain()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    res[name] = {"ms_per_step": round(ms, 4), "pairs_per_s": round(N / ms * 1e3, 1)}
    print(name, res[name], flush=True)


def serial(i):
    b = i & 1
    exs[0](l, r, out=outs[0])
    regression.soft_argmin(logits[b], out=disp[b])


timed("serial_one_stream", serial, lambda: None)

for prio_sa in (0, -1):
    s_ms = [torch.cuda.Stream(), torch.cuda.Stream()]
    s_sa = torch.cuda.Stream(priority=prio_sa)
    cur = torch.cuda.current_stream()

    def fork():
        for s in s_ms + [s_sa]:
            s.wait_stream(cur)

    def join():
        for s in s_ms + [s_sa]:
            cur.wait_stream(s)

    def piped(i):
        b = i & 1
        if i == 0:
            fork()
        with torch.cuda.stream(s_ms[b]):
            exs[b](l, r, out=outs[b])
        with torch.cuda.stream(s_sa):
            regression.soft_argmin(logits[b], out=disp[b])

    # the fork must come after e0.record() on the current stream: redo it per timed run via a wrapper
    state = {"first": True}

    def piped_run(i):
        if i == 0:
            fork()
        b = i & 1
        with torch.cuda.stream(s_ms[b]):
            exs[b](l, r, out=outs[b])
        with torch.cuda.stream(s_sa):
            regression.soft_argmin(logits[b], out=disp[b])

    timed("two_steps_in_flight_sa_prio%d" % prio_sa, piped_run, join)

    def sa_only_overlap(i):
        if i == 0:
            fork()
        b = i & 1
        with torch.cuda.stream(s_ms[0]):
            exs[0](l, r, out=outs[0])
        with torch.cuda.stream(s_sa):
            regression.soft_argmin(logits[b], out=disp[b])

    timed("softargmin_on_own_stream_prio%d" % prio_sa, sa_only_overlap, join)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/overlap_streams.json", "w"), indent=1)
