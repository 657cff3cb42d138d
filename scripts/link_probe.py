"""Micro-benchmarks of the multi-GPU transport (torchrun, one rank per GPU): latency of a halo exchange, of a scalar
all-reduce, and the fine-level V-cycle operator with and without halo/interior overlap.  FEMO_COMM=link|nccl."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from femo_b200 import engine as E, dist as fd
from femo_b200._lib import lib, check

lr = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
rank, R = fd.init(lr)
n = int(os.environ.get('N', '4096'))
p = fd.SlabProblem(E.FAMILY_NLPOISSON_P1, n, n * R, rank, R, lo=(0.0, 0.0), hi=(1.0, float(R)))
p.enable_multigrid()
p.upload(lr)
u, f = p.new_vector(p.N, 0.0), p.new_vector(p.M[0], 0.1)
p.set_coefficient(0, u); p.set_coefficient(1, f)
p.newton_solve(kind='SNES', krylov_rtol=1e-10, precond=2, cheb_degree=2)      # sets up the hierarchy


def timeit(fn, reps):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, (time.perf_counter() - t0) / reps * 1e6


x = p.new_vector(p.N, 1.0)
out = {}
out['halo_us (gpu, host)'] = timeit(lambda: p.halo(x), 500)
out['dia PLAIN (mode 0) overlap'] = timeit(lambda: p.vcycle_op_probe(0), 200)
os.environ['FEMO_NO_OVERLAP'] = '1'
out['dia PLAIN (mode 0) no overlap'] = timeit(lambda: p.vcycle_op_probe(0), 200)
os.environ.pop('FEMO_NO_OVERLAP')
vals, _ = p.assemble_jacobian()
b = p.new_vector(p.N, 1.0)
l0 = p.launch_count(); s0 = fd.stats()
t0 = time.perf_counter()
xs, info = p.linear_solve(vals, b, rtol=1e-10, precond=2, cheb_degree=2)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
s1 = fd.stats()
out['pcg solve'] = dict(ms=dt * 1e3, its=info['iterations'], launches=p.launch_count() - l0,
                        halos=s1['halo_exchanges'] - s0['halo_exchanges'], allreduces=s1['allreduces'] - s0['allreduces'])
if rank == 0:
    print('ranks', R, 'backend', fd.stats()['backend'])
    for k, v in out.items():
        print(k, v)
fd.finalize()
dist.destroy_process_group()
