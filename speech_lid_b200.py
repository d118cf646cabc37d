"""Import alias: ``import speech_lid_b200`` loads the package that lives in ``speech-lid_b200/``.

The directory name is fixed by the project layout and is not a valid Python identifier, so this
one-file shim registers it under an importable name (sub-modules resolve through ``__path__``).
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "speech-lid_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
