"""Keep the reference's import lines: `import femo_b200.compat` (or `femo_b200.compat.install()`) registers the module names
of RuruX/femo -- `femo`, `femo.fea`, `femo.fea.fea_dolfinx`, `femo.fea.utils_dolfinx`, `femo.csdl_opt`,
`femo.csdl_opt.fea_model / state_model / output_model` -- as aliases of their femo_b200 counterparts, so that

    from femo.fea.fea_dolfinx import *                      # FEA, Function, FunctionSpace, update, assemble, ...
    from femo.csdl_opt.fea_model import FEAModel
    from femo.csdl_opt.state_model import StateModel

in a user script resolve to the B200 engine.  Nothing is installed if a real `femo` package is importable (the reference
itself, over dolfinx) unless force=True.  The lower face is also what the reference's OWN femo/csdl_opt modules need: with only
`femo.fea.fea_dolfinx` aliased (install(csdl_opt=False)) femo's unmodified state / output operations run on top of
femo_b200.fea (their call sequence against this lower face is pinned in tests/test_upper_face.py)."""
import importlib
import importlib.util
import sys
import types

_ALIASES = {
    'femo.fea.fea_dolfinx': 'femo_b200.fea.fea_b200',
    'femo.fea.utils_dolfinx': 'femo_b200.fea.utils_b200',
    'femo.csdl_opt.fea_model': 'femo_b200.csdl_opt.fea_model',
    'femo.csdl_opt.state_model': 'femo_b200.csdl_opt.state_model',
    'femo.csdl_opt.output_model': 'femo_b200.csdl_opt.output_model',
}


def install(force=False, csdl_opt=True):
    """Register the aliases; returns the list of module names installed (empty if a real femo is present)."""
    if not force and 'femo' not in sys.modules:
        try:
            if importlib.util.find_spec('femo') is not None:
                return []
        except (ImportError, ValueError):
            pass
    done = []

    def package(name):
        if name not in sys.modules or force:
            m = types.ModuleType(name)
            m.__path__ = []
            m.__femo_b200_alias__ = True
            sys.modules[name] = m
            done.append(name)
        return sys.modules[name]
    package('femo')
    package('femo.fea')
    if csdl_opt:
        package('femo.csdl_opt')
    for alias, target in _ALIASES.items():
        if alias.startswith('femo.csdl_opt') and not csdl_opt:
            continue
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        parent, _, leaf = alias.rpartition('.')
        setattr(sys.modules[parent], leaf, mod)
        done.append(alias)
    setattr(sys.modules['femo'], 'fea', sys.modules['femo.fea'])
    if csdl_opt:
        setattr(sys.modules['femo'], 'csdl_opt', sys.modules['femo.csdl_opt'])
    return done


def uninstall():
    for name in list(sys.modules):
        if name == 'femo' or name.startswith('femo.'):
            m = sys.modules[name]
            if getattr(m, '__femo_b200_alias__', False) or name in _ALIASES:
                del sys.modules[name]
