"""Manual probe: CUDA-graph capture of module forward+backward vs eager."""
import sys, pathlib, time, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200, bench
dev = torch.device("cuda:0")
inp = bench.make_c2(dev, 0, torch.bfloat16)
m = xfmr_b200.InfomationNoiseContrastiveEstimationLoss(sigma=5.0, margin=0.5)
q = inp["user_embed"].detach().requires_grad_(True)
v = inp["item_embed"].detach().requires_grad_(True)
def step():
    loss = m(q, v, inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"])
    dq, dv = torch.autograd.grad(loss, (q, v))
    return loss, dq, dv
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = step()
torch.cuda.synchronize()
ref = step()
g.replay(); torch.cuda.synchronize()
print("graph loss", float(out[0]), "eager loss", float(ref[0]), "dq match", torch.equal(out[1], ref[1]), "dv match", torch.equal(out[2], ref[2]))
flush = torch.zeros(64 << 20, device=dev)
def timeit(fn, n=50):
    ts = []
    for _ in range(n):
        flush.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)
print("eager ms/step", timeit(step), " graph ms/step", timeit(g.replay))
