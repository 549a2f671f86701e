"""Import the unmodified reference script as a module (build container only).

/root/reference/v2/model_train_test.py imports matplotlib and imageio at the
top (v2:6,14); neither is installed here.  They are only used by plotting
code that is outside the hot path, so empty stand-in modules are enough.
tqdm is silenced.  No reference source is copied: the file is executed from
where it lies.  On the GPU box /root/reference does not exist and
`available()` is False; tests then rely on tests/golden/*.npz.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("LDM_REFERENCE_ROOT", "/root/reference")
_cache = {}


def available(version="v2"):
    return os.path.isfile(os.path.join(REF_ROOT, version, "model_train_test.py"))


def load(version="v2"):
    """Return the reference script as a module object (cached)."""
    if version in _cache:
        return _cache[version]
    path = os.path.join(REF_ROOT, version, "model_train_test.py")
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    import torch  # noqa: F401  (make sure torch is initialised before the script reseeds it)
    for name in ("matplotlib", "matplotlib.pyplot", "imageio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if version == "v3":   # v3:25-26 mounts a Colab drive at import time
        for name in ("google", "google.colab", "google.colab.drive"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        sys.modules["google"].colab = sys.modules["google.colab"]
        sys.modules["google.colab"].drive = sys.modules["google.colab.drive"]
        sys.modules["google.colab.drive"].mount = lambda *a, **k: None
    spec = importlib.util.spec_from_file_location("_ldm_reference_" + version, path)
    mod = importlib.util.module_from_spec(spec)
    rng_state = torch.get_rng_state()
    real_flowers = None
    if version in ("v4", "v5"):   # v4:27-32 downloads Flowers102 at import time: an empty stand-in dataset, restored below
        import torch.utils.data
        import torchvision.datasets

        class _NoFlowers(torch.utils.data.Dataset):
            classes = []

            def __init__(self, *a, **k):
                pass

            def __len__(self):
                return 1

            def __getitem__(self, i):
                raise IndexError(i)

        real_flowers = torchvision.datasets.Flowers102
        torchvision.datasets.Flowers102 = _NoFlowers
    try:
        spec.loader.exec_module(mod)          # runs torch.manual_seed(42) (v2:17)
    finally:
        if real_flowers is not None:
            torchvision.datasets.Flowers102 = real_flowers
    torch.set_rng_state(rng_state)
    mod.tqdm = lambda it, **kw: it        # v2:596 wraps the sampling loop in tqdm
    _cache[version] = mod
    return mod
