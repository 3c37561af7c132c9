"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference from /root/reference (this container only).

/root/reference does not exist on the GPU box; nothing on the GPU-side run path (``-m gpu`` tests, smoke(),
bench.py) may import this module.  It exists to (1) pin ``oracle/port.py`` against the real reference and
(2) generate ``tests/golden/*`` (see gen_golden.py).

How the reference is made importable on a CPU-only box without touching it:
  * ``tinycudann`` -> stub module whose ``Encoding`` is oracle.torch_hashgrid.TorchHashGrid
  * ``matplotlib``/``termcolor`` -> empty stubs (only used by visualisation helpers)
  * ``nerf_util.sample_dists`` default ``device="cuda"`` (projects/nerf/utils/nerf_util.py:20) re-bound to cpu
  * the literal ``{DATASET_FOLDER}`` placeholder in the YAMLs is substituted textually before ``Config()``
"""
import os
import sys
import tempfile
import types
from functools import partial

REF_ROOT = os.environ.get("MLI_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "projects", "NeuralLumen"))


_loaded = {}


def _install_stubs():
    from oracle.torch_hashgrid import TorchHashGrid
    if "tinycudann" not in sys.modules:
        tcnn = types.ModuleType("tinycudann")
        tcnn.Encoding = TorchHashGrid
        sys.modules["tinycudann"] = tcnn
    for name in ("matplotlib", "matplotlib.pyplot", "termcolor"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod = types.ModuleType(name)
                if name == "termcolor":
                    mod.colored = lambda s, *a, **k: s
                sys.modules[name] = mod
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def load():
    """Returns a namespace with the reference modules (imported once)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from projects.nerf.utils import nerf_util
    nerf_util.sample_dists = partial(nerf_util.sample_dists, device="cpu")
    from projects.nerf.utils import camera, render
    from projects.neuralangelo.utils import misc as angelo_misc
    from projects.NeuralLumen.utils import utils as lumen_utils
    from projects.NeuralLumen import model as lumen_model
    from imaginaire.config import Config
    _loaded.update(nerf_util=nerf_util, camera=camera, render=render, angelo_misc=angelo_misc,
                   lumen_utils=lumen_utils, lumen_model=lumen_model, Config=Config)
    return types.SimpleNamespace(**_loaded)


def load_config(name="syn_hotdog_b", overrides=None):
    """Config() of an as-shipped YAML with the {DATASET_FOLDER} placeholder substituted (not valid YAML as is)."""
    ref = load()
    src = os.path.join(REF_ROOT, "projects", "NeuralLumen", "configs", name + ".yaml")
    text = open(src).read().replace("{DATASET_FOLDER}", "/nonexistent")
    cwd = os.getcwd()
    os.chdir(REF_ROOT)  # _parent_ paths are relative to the reference root
    try:
        with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
            f.write(text)
        cfg = ref.Config(f.name)
    finally:
        os.chdir(cwd)
        os.unlink(f.name)
    for key, val in (overrides or {}).items():
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = getattr(node, p)
        setattr(node, parts[-1], val)
    return cfg


def build_model(cfg, progress=0.5, iteration=None, training=True):
    """Reference Model on CPU with the trainer-side attributes set by hand (neuralangelo/trainer.py:30-34,65-76)."""
    ref = load()
    model = ref.lumen_model.Model(cfg.model, cfg.data)
    model.progress = progress
    sdf = model.neural_sdf
    sdf.warm_up_end = cfg.optim.sched.warm_up_end
    if cfg.model.object.sdf.encoding.coarse2fine.enabled:
        sdf.set_active_levels(iteration if iteration is not None else int(progress * cfg.max_iter))
    sdf.set_normal_epsilon()
    model.train(training)
    if hasattr(model, "bounding_box_aabb"):
        model.bounding_box_aabb = model.bounding_box_aabb.float()
    return model
