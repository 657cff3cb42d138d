"""Regenerates tests/golden/symbolic/*.npz: the exact (sympy) values of R, dR/du, dR/dm, J and its partials for every form
family on small distorted meshes, derived in tests/test_oracle_symbolic.py from the reference's weak forms.  The same test
module, run without FEMO_SYMBOLIC_DUMP, checks that the committed files are reproduced.

    python scripts/make_symbolic_golden.py
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
env = dict(os.environ, FEMO_SYMBOLIC_DUMP='1')
sys.exit(subprocess.call([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_oracle_symbolic.py'), '-q'], env=env, cwd=ROOT))
