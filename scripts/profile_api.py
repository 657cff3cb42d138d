"""cProfile of the API-level step (host numpy in/out) at the bench size."""
import sys, os, io, contextlib, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
api = bench.ApiStep(n)
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(3):
        api.step()
    import torch
    torch.cuda.synchronize()
    t = time.perf_counter()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        api.step()
    torch.cuda.synchronize()
    pr.disable()
    dt = (time.perf_counter() - t) / 3
print('api step %.1f ms' % (dt * 1e3))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:9000])
