import sys, torch
sys.path.insert(0, ".")
import qwtdh_b200 as qw
prob = qw.Problem(n=16, topology="circular", schedule="lin", n_steps=1024)
T = torch.tensor([16.0], dtype=torch.float64, device="cuda"); g = torch.tensor([1.0], dtype=torch.float64, device="cuda")
best = 1e9
for _ in range(200):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); p, _ = qw.rk4_batch(prob, T, g); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("C1 min over 200: %.1f us, p=%.14f" % (best * 1e3, float(p[0])))
