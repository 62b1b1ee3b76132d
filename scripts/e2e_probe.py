"""Where does the gap between the device-resident loop and the host-buffer call go?"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smoothsde_b200 import devgen
eng, par, info = devgen.make_ctcrw_device(1024, 100000, seed=20260103, device=0)
n = info["n"]
for _ in range(3):
    eng.eval(par, 1)
wall, dev = [], []
for _ in range(20):
    t0 = time.perf_counter(); eng.eval(par, 1); wall.append(time.perf_counter() - t0); dev.append(eng.last_eval_ms)
print("host-buffer eval: wall ms", np.round(np.array(wall) * 1e3, 3)[:8], "median", np.median(wall) * 1e3)
print("events around the kernels (ms):", np.round(dev, 3)[:8], "median", np.median(dev))
dev_t = torch.device("cuda", 0)
par_dev = torch.as_tensor(par, device=dev_t); out = torch.zeros(eng.n_par + 2, dtype=torch.float64, device=dev_t)
st = torch.cuda.Stream(device=dev_t); torch.cuda.set_stream(st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    eng.eval_device(par_dev.data_ptr(), out.data_ptr(), 1, st.cuda_stream)
e1.record(); torch.cuda.synchronize()
print("device loop ms/eval", e0.elapsed_time(e1) / 20)
# device loop with a sync after every evaluation
t0 = time.perf_counter()
for _ in range(20):
    eng.eval_device(par_dev.data_ptr(), out.data_ptr(), 1, st.cuda_stream); torch.cuda.synchronize()
print("device call + sync, wall ms/eval", (time.perf_counter() - t0) / 20 * 1e3)
