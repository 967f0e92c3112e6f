"""Experiment: the fused kernel's two bulk phases (fp32 reads, then fp32 writes) are each bound by one direction of HBM
traffic and every CTA of a launch is in the same phase at the same time.  Does running two half-batches on two streams,
offset in time, mix reads and writes well enough to raise total throughput?  Prints tokens/s for
  (a) one stream, M = 16384;  (b) two streams, M = 8192 each, started together;  (c) same with stream 2 delayed."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from outlier_suppression_b200 import ops

torch.manual_seed(0)
dev = "cuda"
SITES = [(768, 2304), (768, 768), (768, 3072), (3072, 768)]
LAYERS = 12


def make_sites():
    out = []
    for k, n in SITES:
        w = torch.randn(n, k, device=dev) * 0.05
        ws = (w.abs().amax(1) / 31.5).contiguous(); wz = torch.zeros(n, dtype=torch.int32, device=dev)
        codes, rowsum = ops.pack_weight(w, ws, wz, -32, 31)
        out.append((k, n, codes, ws, rowsum, torch.randn(n, device=dev)))
    return out


def make_bufs(m):
    acts = {768: [torch.randn(m, 768, device=dev) for _ in range(4)], 3072: [torch.randn(m, 3072, device=dev) for _ in range(2)]}
    outs = {768: [torch.empty(m, 768, device=dev) for _ in range(4)], 2304: [torch.empty(m, 2304, device=dev) for _ in range(2)],
            3072: [torch.empty(m, 3072, device=dev) for _ in range(2)]}
    return acts, outs


a_scale = torch.tensor([0.1], device=dev); a_zp = torch.tensor([31.0], device=dev)
sites = make_sites()


def make_step(m):
    acts, outs = make_bufs(m)
    cnt = {768: 0, 3072: 0, 2304: 0}

    def step():
        for _ in range(LAYERS):
            for k, n, codes, ws, rowsum, bias in sites:
                a = acts[k][cnt[k] % len(acts[k])]; cnt[k] += 1
                o = outs[n][cnt[n] % len(outs[n])]
                ops.fused_fq_linear(a, a_scale, a_zp, 0, 63, codes, ws, rowsum, bias, lsq_grad_factor=1e-4, out=o)
    return step


def graph_of(fn, stream):
    with torch.cuda.stream(stream):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            fn()
    return g


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
steps = [make_step(16384), make_step(8192), make_step(8192)]  # keep the closures (and their buffers) alive
g_full = graph_of(steps[0], s1)
g_a = graph_of(steps[1], s1)
g_b = graph_of(steps[2], s2)
torch.cuda.synchronize()
res = {}
STEPS = 10


def timed(run):
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def one_stream():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    with torch.cuda.stream(s1):
        for _ in range(STEPS):
            g_full.replay()
    cur.wait_stream(s1)


def two_streams(delay_cycles):
    def run():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            for _ in range(STEPS):
                g_a.replay()
        with torch.cuda.stream(s2):
            if delay_cycles:
                torch.cuda._sleep(delay_cycles)
            for _ in range(STEPS):
                g_b.replay()
        cur.wait_stream(s1); cur.wait_stream(s2)
    return run


ms = timed(one_stream); res["one_stream_M16384_tok_s"] = STEPS * 16384 / (ms * 1e-3)
for d in (0, 20000, 40000, 60000, 100000):
    ms = timed(two_streams(d)); res["two_streams_M8192_delay%d_tok_s" % d] = STEPS * 16384 / (ms * 1e-3)
print(json.dumps(res, indent=1))
