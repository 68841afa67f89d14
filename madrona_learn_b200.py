"""Import shim: `import madrona_learn_b200` loads the package that lives in the directory
`madrona-learn_b200/` (a hyphen is not importable, the task fixes the directory name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'madrona-learn_b200')
_spec = importlib.util.spec_from_file_location(
    'madrona_learn_b200', os.path.join(_dir, '__init__.py'), submodule_search_locations=[_dir])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules['madrona_learn_b200'] = _pkg
_spec.loader.exec_module(_pkg)
