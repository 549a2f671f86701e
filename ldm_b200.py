"""Import shim: the package directory is named `oxford-102-flower-gan-vae-latent-diffusion_b200` (hyphens
are not valid in a Python identifier), so it is loaded here under the module name `ldm_b200`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oxford-102-flower-gan-vae-latent-diffusion_b200")
_spec = importlib.util.spec_from_file_location("ldm_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ldm_b200"] = _mod
_spec.loader.exec_module(_mod)
