"""Drive the UNMODIFIED reference step loops on the CPU (TEST INFRASTRUCTURE).

Only usable where /root/reference exists (the build container); nothing that runs on the
GPU box imports this module.  It is the pin for oracle/steps.py and the source of
tests/golden/*.json (see oracle/make_golden.py).

What is stubbed, and why:
  * ``torchinfo`` and ``matplotlib`` -- imported by train/dcgan_trainer.py:6,26 but absent from
    this image.  ``summary`` returns '' ; ``pyplot.plot`` records its arguments, which is how the
    trainer's local ``losses_d`` / ``losses_g`` lists (dcgan_trainer.py:232-233) are recovered.
  * ``Metrics`` -- needs ./save/iception_v3/loss_bset.pt and a CIFAR-100 download
    (metrics.py:51,56); replaced by constants.
  * the data preprocessor -- replaced by an object whose ``get_data_loader()`` returns a list
    of pre-built batches (the trainer only iterates, len()s and next(iter())s it).
  * ``torch.randn`` / ``torch.rand`` -- replay the caller's tensors in draw order so the oracle
    and the CUDA path can be fed the identical noise; ``torch.save`` -- no-op (75 MB per call).
  * ``torch.nn.functional.dropout`` (CGAN only) -- replays keep-masks.
The step arithmetic itself is the reference's own code, untouched.
"""
import contextlib
import os
import sys
import tempfile
import types

import torch

REF_ROOT = os.environ.get("JCK_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "train", "dcgan_trainer.py"))


class _Replay:
    def __init__(self, tensors):
        self.q = list(tensors)
        self.i = 0

    def pop(self, shape):
        if self.i >= len(self.q):
            raise RuntimeError(f"replay queue exhausted at draw {self.i} (shape {tuple(shape)})")
        t = self.q[self.i]
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"replay draw {self.i}: reference asked for {tuple(shape)}, "
                               f"queue holds {tuple(t.shape)}")
        self.i += 1
        return t.clone()


def _shape_of(args):
    if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)):
        return tuple(args[0])
    return tuple(args)


@contextlib.contextmanager
def _patched(randn_q, rand_q, drop_q=None):
    plots = []
    plt = types.ModuleType("matplotlib.pyplot")

    def _noop(*a, **k):
        return None

    for name in ("axis", "title", "imshow", "savefig", "clf", "xlabel", "ylabel", "legend",
                 "close"):
        setattr(plt, name, _noop)
    plt.plot = lambda *a, **k: plots.append((a, k))

    class _Fig:
        def add_subplot(self, *a, **k):
            return None

    plt.figure = lambda *a, **k: _Fig()
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    tinfo = types.ModuleType("torchinfo")
    tinfo.summary = lambda *a, **k: ""

    saved_mods = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot", "torchinfo")}
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "torchinfo": tinfo})
    sys.path.insert(0, REF_ROOT)

    o_randn, o_rand, o_save, o_drop = torch.randn, torch.rand, torch.save, torch.nn.functional.dropout
    torch.randn = lambda *a, **k: randn_q.pop(_shape_of(a))
    torch.rand = lambda *a, **k: rand_q.pop(_shape_of(a))
    torch.save = lambda *a, **k: None
    if drop_q is not None:
        def _drop(x, p=0.5, training=True, inplace=False):
            if not training:
                return x
            return x * drop_q.pop(x.shape) / (1.0 - p)
        torch.nn.functional.dropout = _drop
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="jck_ref_")
    os.chdir(tmp)
    try:
        yield plots, tmp
    finally:
        os.chdir(cwd)
        torch.randn, torch.rand, torch.save = o_randn, o_rand, o_save
        torch.nn.functional.dropout = o_drop
        sys.path.remove(REF_ROOT)
        for k, v in saved_mods.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        # leave no reference modules cached: the product package reuses the names
        for k in list(sys.modules):
            if k.split(".")[0] in ("train", "model", "preprocess", "logger", "metrics", "utils",
                                   "enums", "change_randomseed"):
                f = getattr(sys.modules[k], "__file__", "") or ""
                # train/ and logger/... may be namespace packages (no __init__.py): match __path__
                paths = [str(p) for p in (getattr(sys.modules[k], "__path__", None) or [])]
                if f.startswith(REF_ROOT) or any(p.startswith(REF_ROOT) for p in paths):
                    del sys.modules[k]


class _FakeMetrics:
    def __init__(self, *a, **k):
        pass

    def inception_score(self, *a, **k):
        return 1.0

    def fid(self, *a, **k):
        return 1.0

    def intra_fid(self, *a, **k):
        return 1.0


class _FakeData:
    idx_to_labels = {i: str(i) for i in range(100)}

    def __init__(self, batches):
        self.batches = batches

    def get_data_loader(self):
        return self.batches, None


def _args(lr, batch, tmp):
    return types.SimpleNamespace(epoch=1, max_learning_rate=lr, model_path="ref", log_file=0,
                                 batch_size=batch, num_worker=0, save_path=os.path.join(tmp, "save"))


def _hook_convs(module, store):
    """Record every conv raw output (and its gradient) per forward call, in call order."""
    handles = []
    for name, sub in module.named_modules():
        if name.startswith("conv"):
            def fwd(mod, inp, outp, name=name):
                rec = {"out": outp.detach().clone(), "grad": None}
                store.setdefault(name, []).append(rec)
                if outp.requires_grad:
                    outp.register_hook(lambda gr, rec=rec: rec.__setitem__("grad", gr.detach().clone()))
            handles.append(sub.register_forward_hook(fwd))
    return handles


def run_dcgan(real_batches, rng_steps, fixed_noise, lr, seed=12345, hook=False):
    """Run DCGANTrainer.train() (dcgan_trainer.py:130-239) for len(real_batches) iterations.

    real_batches: list of [B,3,64,64] tensors; rng_steps: list of dicts as oracle.steps.make_rng.
    Returns losses (from the captured plt.plot call), final state and optional hook captures.
    """
    randn_list = [fixed_noise]
    rand_list = []
    for r in rng_steps:
        randn_list += [r["noise_real"], r["z"], r["noise_fake"]]
        rand_list += [r["alpha"]]
    with _patched(_Replay(randn_list), _Replay(rand_list)) as (plots, tmp):
        torch.manual_seed(seed)
        from model import DCGAN as ref_dcgan
        from train import dcgan_trainer as ref_tr
        ref_tr.Metrics = _FakeMetrics
        g, d = ref_dcgan.Generator(), ref_dcgan.Discriminator()
        trainer = ref_tr.DCGANTrainer(_args(lr, real_batches[0].shape[0], tmp), g, d,
                                      _FakeData([(x,) for x in real_batches]))
        caps = {"d": {}, "g": {}}
        handles = (_hook_convs(trainer.model_d, caps["d"]) + _hook_convs(trainer.model_g, caps["g"])) if hook else []
        trainer.train()
        for h in handles:
            h.remove()
        (_, losses_d), (_, losses_g) = plots[0][0], plots[1][0]
        return {
            "losses_d": list(losses_d), "losses_g": list(losses_g),
            "g_state": {k: v.detach().clone() for k, v in trainer.model_g.state_dict().items()},
            "d_state": {k: v.detach().clone() for k, v in trainer.model_d.state_dict().items()},
            "opt_d": trainer.optimizer_d.state_dict(), "opt_g": trainer.optimizer_g.state_dict(),
            "hooks": caps,
        }


def run_cgan(real_batches, label_batches, rng_steps, fixed_noise_list, lr, seed=12345):
    """Run CGANTrainer.train() (cgan_trainer.py:134-270).  fixed_noise_list: the 100 [10,100,1,1]
    draws of cgan_trainer.py:144-150."""
    randn_list = list(fixed_noise_list)
    rand_list, drops = [], []
    for r in rng_steps:
        randn_list += [r["noise_real"], r["z"], r["noise_fake"]]
        rand_list += [r["alpha"]]
        drops += list(r["drop"])
    with _patched(_Replay(randn_list), _Replay(rand_list), _Replay(drops)) as (plots, tmp):
        torch.manual_seed(seed)
        from model import CGAN as ref_cgan
        from train import cgan_trainer as ref_tr
        ref_tr.Metrics = _FakeMetrics
        g, d = ref_cgan.Generator(), ref_cgan.Discriminator()
        trainer = ref_tr.CGANTrainer(_args(lr, real_batches[0].shape[0], tmp), g, d,
                                     _FakeData(list(zip(real_batches, label_batches))))
        trainer.train()
        (_, losses_d), (_, losses_g) = plots[0][0], plots[1][0]
        return {
            "losses_d": list(losses_d), "losses_g": list(losses_g),
            "g_state": {k: v.detach().clone() for k, v in trainer.model_g.state_dict().items()},
            "d_state": {k: v.detach().clone() for k, v in trainer.model_d.state_dict().items()},
        }


def reference_modules(seed=12345):
    """Fresh reference G, D after seeding + weights_init (dcgan_trainer.py:54-55)."""
    with _patched(_Replay([]), _Replay([])):
        torch.manual_seed(seed)
        from model import DCGAN as ref_dcgan
        g, d = ref_dcgan.Generator(), ref_dcgan.Discriminator()
        g.apply(ref_dcgan.weights_init)
        d.apply(ref_dcgan.weights_init)
        return g, d
