"""
Model-file parser: psfMC's model-definition syntax, unchanged.

A model file is Python; every *bare expression statement* at top level that
evaluates to a component becomes a model component, in file order; component and
prior names are available without imports; relative file names inside the model
are resolved against the model file's directory
(reference: /root/reference/psfMC/model_parser.py:26-66).

Implementation: each top-level expression statement is rewritten to a call of a
collector injected into the execution namespace. Explicit
``from psfMC.ModelComponents import ...`` / ``from psfMC.distributions import ...``
lines in existing model files keep working: when the real psfMC package is not
importable those module names are aliased to this package's equivalents.
"""
import ast
import os
import sys
import types

from . import components as _components
from . import distributions as _distributions

_COLLECT = '__psfmc_collect__'


def install_psfmc_aliases():
    """Make ``import psfMC.ModelComponents`` / ``psfMC.distributions`` resolve to
    this package when the reference package is not installed."""
    if 'psfMC' in sys.modules:
        return
    try:
        import psfMC  # noqa: F401  (a real installation wins)
        return
    except Exception:
        pass
    root = types.ModuleType('psfMC')
    root.__path__ = []
    root.ModelComponents = _components
    root.distributions = _distributions
    sys.modules['psfMC'] = root
    sys.modules['psfMC.ModelComponents'] = _components
    sys.modules['psfMC.distributions'] = _distributions


class _CollectExpressions(ast.NodeTransformer):
    """Top-level ``expr`` -> ``__psfmc_collect__(expr)``; nested scopes untouched."""

    def visit_Module(self, node):
        body = []
        for stmt in node.body:
            if isinstance(stmt, ast.Expr):
                call = ast.Call(func=ast.Name(id=_COLLECT, ctx=ast.Load()),
                                args=[stmt.value], keywords=[])
                stmt = ast.copy_location(ast.Expr(value=call), stmt)
            body.append(stmt)
        node.body = body
        return node


def component_list_from_file(filename):
    """Execute a model file and return its components in file order."""
    install_psfmc_aliases()
    with open(filename) as fobj:
        tree = ast.parse(fobj.read(), filename=filename)
    tree = ast.fix_missing_locations(_CollectExpressions().visit(tree))

    collected = []
    namespace = {'__file__': os.path.abspath(filename), '__name__': '__psfmc_model__'}
    for module in (_components, _distributions):
        for name in getattr(module, '__all__', None) or dir(module):
            if not name.startswith('_'):
                namespace[name] = getattr(module, name)
    namespace[_COLLECT] = collected.append

    previous_dir = os.getcwd()
    model_dir = os.path.dirname(os.path.abspath(filename))
    try:
        os.chdir(model_dir)
        exec(compile(tree, filename, 'exec'), namespace)
    finally:
        os.chdir(previous_dir)
    return [obj for obj in collected if _is_component(obj)]


def _is_component(obj):
    if isinstance(obj, _components.ComponentBase):
        return True
    # components of a real psfMC installation, if the model file imported them
    return hasattr(obj, '_priors') and hasattr(obj, '_constants')
