"""Manual probe: compares the CUDA semi-hard selection with the oracle's on the C1 shape."""
import ctypes, sys, pathlib
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200 as xb
from xfmr_b200 import _lib, synthetic
from xfmr_b200.losses import _make_desc, _loss_fwd
from oracle import losses_oracle

dev = torch.device("cuda:0")
B, N, d, P, K = 1024, 3706, 64, 32, 4
inp = {k: v.to(dev) for k, v in synthetic.make_loss_inputs(B, N, d, P, n_catalog=N, seed=0).items()}
losses, ws = _loss_fwd(inp["user_embed"], inp["item_embed"], inp["target"], inp["item_idx"], inp["pos_idx"], None, K, 1.0, 1.0, 1 << 1, 1)
desc = _make_desc(B, N, d, P, 0, 1, K, 1 << 1, 1.0, 1.0, False)
def region(r, dtype):
    off, nb = ctypes.c_size_t(), ctypes.c_size_t()
    _lib.check(_lib.lib.xb_debug_loss_region(ctypes.byref(desc), r, ctypes.byref(off), ctypes.byref(nb)), "region")
    return ws[off.value: off.value + nb.value].view(dtype)
selcol = region(0, torch.int32).view(B, K).long()
q, v = inp["user_embed"].double(), inp["item_embed"].double()
logits = -losses_oracle.half_squared_distance(q, v) * inp["target"].sign().double().unsqueeze(1)
mask = losses_oracle.negative_mask(inp["item_idx"], inp["pos_idx"], B)
sel = losses_oracle.semi_hard_selection(logits, mask.clone(), K)
ora = [set(sel[i].nonzero().flatten().tolist()) for i in range(B)]
mine = [set(c for c in selcol[i].tolist() if c >= 0) for i in range(B)]
bad = [i for i in range(B) if ora[i] != mine[i]]
print("rows with different selection:", len(bad), "of", B)
R = logits - logits.diagonal().unsqueeze(1)
cand = region(4, torch.int64).view(B, -1)
for i in bad[:6]:
    print("row", i, "oracle", sorted(ora[i]), "cuda", sorted(mine[i]))
    print("   R oracle:", [f"{R[i, j].item():+.3e}" for j in sorted(ora[i])], " R cuda:", [f"{R[i, j].item():+.3e}" for j in sorted(mine[i])])
    print("   masked? cuda cols:", [bool(~mask[i, j]) for j in sorted(mine[i])])
    cols = [(~(c & 0xffffffff)) & 0xffffffff for c in cand[i].tolist() if c != 0]
    print("   n candidates", len(cols), "oracle cols in candidates:", [j in cols for j in sorted(ora[i])])
    nsemi = int(((R[i] < 0) & mask[i]).sum()); print("   #semi-hard valid", nsemi, "#valid", int(mask[i].sum()))
    half = len(cand[i]) // 2
    for nm, part in (("A", cand[i][:half]), ("B", cand[i][half:])):
        cc = [(~(c & 0xffffffff)) & 0xffffffff for c in part.tolist() if c != 0]
        print("   list", nm, [f"{R[i, j].item():+.2e}" for j in cc])
    Rm = R[i].clone(); Rm[~mask[i]] = float("nan")
    neg = Rm[Rm < 0].sort(descending=True).values[:6]; pos = Rm[Rm >= 0].sort().values[:6]
    print("   true closest semi-hard:", [f"{x:+.2e}" for x in neg.tolist()], " closest hard:", [f"{x:+.2e}" for x in pos.tolist()])
