"""Debug helper (torchrun, 2 ranks): the API-level partitioned run of tests/dist_check_api.py with the adjoint solve taken apart."""
import contextlib, io, os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from femo_b200 import engine as E
from femo_b200 import dist as fd

same = os.environ.get('FEMO_DIST_SAME_DEVICE') == '1'
lr = 0 if same else int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(lr)
dist.init_process_group('gloo' if same else 'nccl', **({} if same else dict(device_id=torch.device('cuda', lr))))
rank, R = fd.init(lr)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
from femo_b200.fea.fea_b200 import FEA, createUnitSquareMesh, FunctionSpace, Function, TestFunction
from femo_b200.forms.nonlinear_poisson import pdeRes, outputForm
from femo_b200.csdl_opt import FEAModel, Simulator
mesh = createUnitSquareMesh(n)
fea = FEA(mesh)
f = Function(FunctionSpace(mesh, ('DG', 0)))
Vu = FunctionSpace(mesh, ('CG', 1))
u = Function(Vu)
res = pdeRes(u, TestFunction(Vu), f)
fea.add_input('f', f)
fea.add_state(name='u', function=u, residual_form=res, arguments=['f'])
fea.add_output(name='l2_functional', type='scalar', form=outputForm(u, f), arguments=['f', 'u'])
fea.PDE_SOLVER = 'SNES'
fea.REPORT = False
model = FEAModel(fea=[fea], debug_mode=False)
nf = fea.inputs_dict['f']['shape']
rows = n // R
gcell = np.arange(2 * n * n, dtype=np.float64).reshape(n, 2 * n)[rank * rows:(rank + 1) * rows].ravel()
model.create_input('f', shape=nf, val=0.1 + 0.05 * np.sin(gcell * 0.37))
sim = Simulator(model)
with contextlib.redirect_stdout(io.StringIO()):
    sim.run()
p = res.fam.problem
def st(name, t):
    print('[rank %d] %s: finite %s norm %.6e size %d' % (rank, name, bool(torch.isfinite(t).all()), float(t.norm()), t.numel()), flush=True)
st('u', p.coefficient(0)); st('f', p.coefficient(1))
vals, _ = p.assemble_jacobian()
st('vals', vals)
b = p.assemble_output_grad(0, 0)
st('dJdu', b)
for tr in (False, True):
    for pre, restart in ((0, 0), (2, 1), (2, 0)):
        x = p.new_vector(p.N, 0.0)
        try:
            _, info = p.linear_solve(vals, b, x, transpose=tr, rtol=1e-12, precond=pre, restart=restart, cheb_degree=2, max_it=300)
            print('[rank %d] transpose %s precond %d restart %d -> %r' % (rank, tr, pre, restart, info), flush=True)
        except Exception as e:
            print('[rank %d] transpose %s precond %d restart %d -> EXC %s' % (rank, tr, pre, restart, e), flush=True)
        st('x', x)
try:
    with contextlib.redirect_stdout(io.StringIO()):
        tot = sim.compute_totals('l2_functional', 'f')[('l2_functional', 'f')]
    print('[rank %d] compute_totals ok, norm %.6e' % (rank, float(np.linalg.norm(tot))), flush=True)
except Exception as e:
    print('[rank %d] compute_totals EXC %s' % (rank, e), flush=True)
torch.cuda.synchronize()
fd.finalize()
dist.destroy_process_group()
