"""One top-k select at DiT-XL/2 size (for an ncu launch list of the K2b kernels)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfron_b200 as sfr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 675_129_632
dev = torch.device("cuda:0")
x = torch.empty(n, device=dev).normal_(0, 1e-2, generator=torch.Generator(device=dev).manual_seed(0))
hp = sfr.HotPath(n, dev, sfr.OptConfig())
for _ in range(3):
    m = hp.topk_mask(x, n // 2)
torch.cuda.synchronize()
st = hp.select_state()
print("k", n // 2, "selected", int(m.sum()), "count_eq", st.count_eq, "tie_budget", st.tie_budget)
