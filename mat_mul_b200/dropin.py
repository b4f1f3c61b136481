"""Make the reference's training.py / model.py pick up this package.

The reference imports its environment with `from utils import *`,
`from datasets import *`, `from act import *` (training.py:12-14, act.py:4-5,
datasets.py:9).  install() registers this package's mirrors under those
top-level names, so with the reference's model.py and training.py on sys.path
they run unchanged on the B200 kernels:

    import mat_mul_b200.dropin; mat_mul_b200.dropin.install()
    import training; training.TensorGameTrainingApp().main()
"""
import importlib
import sys

NAMES = ("utils", "datasets", "act")


def install() -> None:
    # act last: it tries to import the reference's model.py, which itself imports nothing from the environment
    for name in NAMES:
        sys.modules[name] = importlib.import_module(f"mat_mul_b200.{name}")


def uninstall() -> None:
    for name in NAMES:
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__name__", "").startswith("mat_mul_b200."):
            del sys.modules[name]
