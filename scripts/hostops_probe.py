"""Host-side bulk copy / accumulate throughput on this box: numpy, torch and csrc/hostops.cpp at several thread counts."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np
    import torch
    from femo_b200 import _hostops as H
    n = 32_000_000
    a = torch.rand(n, dtype=torch.float64).pin_memory().numpy()
    b = torch.zeros(n, dtype=torch.float64).pin_memory().numpy()
    def t(fn, reps=5):
        fn(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        return (time.perf_counter() - t0) / reps * 1e3
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    print('threads %s: femo copy %.1f ms, femo axpy %.1f ms, torch copy_ %.1f ms, torch add_ %.1f ms, np.copyto %.1f ms  (256 MB vectors)' % (
        sys.argv[1], t(lambda: H.copy(b, a)), t(lambda: H.iadd(b, a, -1.0)), t(lambda: tb.copy_(ta)), t(lambda: tb.add_(ta, alpha=-1.0)),
        t(lambda: np.copyto(b, a))), flush=True)
else:
    for th in ('1', '2', '4', '8', '16', ''):
        env = dict(os.environ)
        if th:
            env['OMP_NUM_THREADS'] = th
        subprocess.run([sys.executable, __file__, th or 'default'], env=env)
