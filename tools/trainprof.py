"""host-side profile of the GCL training step (where do the 16 ms go?)"""
import cProfile, pstats, os, sys, io, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcl_b200.training import GclTrainStep
dev = torch.device("cuda:0")
ts = GclTrainStep(dev, samples=4)
for _ in range(8): ts.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): ts.step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue {t_host / 10 * 1e3:.2f} ms/step, incl. GPU drain {t_all / 10 * 1e3:.2f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(10): ts.step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:5000])
