"""Import alias: the package directory carries the (un-importable) long hyphenated name the build contract asks for;
`import nngp_b200` loads it under this short name."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "improving-performances-of-mcmc-for-nearest-neighbor-gaussian-process-models-with-full-data-augmentat_b200")
_spec = importlib.util.spec_from_file_location("nngp_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["nngp_b200"] = _mod
_spec.loader.exec_module(_mod)
