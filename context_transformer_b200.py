"""Import shim: the package directory is ``context-transformer_b200/`` (hyphenated, as the
layout contract names it), which Python cannot import by name.  Importing this module loads
that directory as the package ``context_transformer_b200`` and replaces this shim in
``sys.modules`` so that ``import context_transformer_b200.detection`` etc. work."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "context-transformer_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
