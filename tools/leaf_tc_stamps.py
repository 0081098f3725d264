import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from hier_slam_b200 import _lib, losses
lib = _lib.load()
losses.LEAF_KERNEL = "tcgen05"
S, L, hh, ww = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (74, 550, 480, 640)))
g = torch.Generator().manual_seed(0)
sem = torch.randn(S, hh, ww, generator=g).cuda(); w = (0.3 * torch.randn(L, S, generator=g)).cuda(); b = torch.randn(L, generator=g).cuda()
lab = torch.randint(0, L, (hh, ww), generator=g).int().cuda(); loss = torch.zeros((), device="cuda"); grad = torch.empty_like(sem)
for _ in range(2): losses._run_leaf(lib, sem, lab, w, b, 1.0, hh * ww, loss, grad, False, False)
st = torch.zeros(64, dtype=torch.int64, device="cuda")
lib.hs_leaf_tc_debug(ctypes.c_void_p(st.data_ptr()))
losses._run_leaf(lib, sem, lab, w, b, 1.0, hh * ww, loss, grad, False, False)
torch.cuda.synchronize(); lib.hs_leaf_tc_debug(None)
v = st.cpu().tolist(); v = [x for x in v if x]
print("deltas (cycles):", [v[i + 1] - v[i] for i in range(len(v) - 1)])
