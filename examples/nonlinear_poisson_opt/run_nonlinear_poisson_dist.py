"""The nonlinear Poisson state + adjoint gradient of config 2 on several GPUs THROUGH THE femo API: the same script on every
rank (the reference's nominal MPI.COMM_WORLD, femo/fea/utils_dolfinx.py:32,140-153).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 run_nonlinear_poisson_dist.py --nel 1024

`createUnitSquareMesh` returns this rank's slab (owned rows + a one-cell ghost layer); input / state / gradient arrays hold
the rank's OWNED dofs; halo exchanges and all-reduces happen inside libfemo_b200 (peer-memory transport, FEMO_COMM=nccl
for NCCL)."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from femo_b200 import dist as fd                                                          # noqa: E402
from femo_b200.fea.fea_b200 import *                                                      # noqa: F401,F403,E402
from femo_b200.csdl_opt import FEAModel, Simulator                                        # noqa: E402
from femo_b200.forms.nonlinear_poisson import pdeRes, outputForm                          # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--nel', default='512')
args = parser.parse_args()
same = os.environ.get('FEMO_DIST_SAME_DEVICE') == '1'          # all ranks on cuda:0 (1-GPU box): the link transport still works
local_rank = 0 if same else int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
if same:
    dist.init_process_group('gloo')
else:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
rank, nranks = fd.init(local_rank)

mesh = createUnitSquareMesh(int(args.nel))                    # this rank's y-slab
fea = FEA(mesh)
f = Function(FunctionSpace(mesh, ('DG', 0)))
V = FunctionSpace(mesh, ('CG', 1))
u = Function(V)
fea.add_input('f', f)
fea.add_state(name='u', function=u, residual_form=pdeRes(u, TestFunction(V), f), arguments=['f'])
fea.add_output(name='l2_functional', type='scalar', form=outputForm(u, f), arguments=['f', 'u'])
fea.PDE_SOLVER, fea.REPORT = 'SNES', False
model = FEAModel(fea=[fea], debug_mode=False)
model.create_input('f', shape=fea.inputs_dict['f']['shape'], val=0.1)
sim = Simulator(model)
sim.run()
g = sim.compute_totals('l2_functional', 'f')[('l2_functional', 'f')]
n2 = torch.tensor([float(np.dot(g, g))], device='cpu' if same else 'cuda')
dist.all_reduce(n2)
if rank == 0:
    print('ranks %d, %d x %d cells: J = %.12g, |dJ/df| = %.6e (owned dofs on rank 0: %d of %d)'
          % (nranks, int(args.nel), int(args.nel), float(np.ravel(sim['l2_functional'])[0]), float(n2.sqrt()),
             fea.states_dict['u']['shape'], (int(args.nel) + 1) ** 2))
fd.finalize()
dist.destroy_process_group()
