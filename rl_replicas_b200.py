"""Import shim: ``import rl_replicas_b200`` loads the package that lives in the directory
``reinforcement-learning-replications_b200/`` (named after the reference repo; hyphens are not importable)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reinforcement-learning-replications_b200")
_spec = importlib.util.spec_from_file_location("rl_replicas_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rl_replicas_b200"] = _mod
_spec.loader.exec_module(_mod)
