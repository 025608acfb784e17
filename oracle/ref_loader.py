"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference model files.

This module imports lifeiteng/vall-e's own `valle/modules/{scaling,activation,
embedding,transformer}.py` and `valle/models/{macros,valle}.py` straight from
`/root/reference` (read-only mount) so that the CPU restatement in
`oracle/valle_oracle.py` can be pinned against the real thing and golden vectors
can be generated (`oracle/gen_golden.py`).  Nothing is copied: the files are
exec'd where they lie.  `/root/reference` does not exist on the GPU box, so
nothing under `-m gpu`, `smoke()` or `bench.py` may call this loader.

`import valle` as shipped fails here because `valle/__init__.py:1` drags in
icefall / lhotse / encodec / torchmetrics / matplotlib, none of which is
installed (no network).  The stubs below contain NO model arithmetic except
`make_pad_mask`, which is one comparison (icefall.utils.make_pad_mask, called at
valle/models/valle.py:804-805).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VALLE_REFERENCE_ROOT", "/root/reference")

_PREFIX = "_valle_ref"  # private top-level name: never shadows the real `valle`


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "valle", "models", "valle.py"))


def _mod(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def _install_stubs():
    import torch
    import torch.nn as nn

    # --- icefall.utils -------------------------------------------------
    if "icefall" not in sys.modules:
        icefall = _mod("icefall")
        iu = _mod("icefall.utils")
        icefall.utils = iu

        def make_pad_mask(lengths: torch.Tensor, max_len: int = 0) -> torch.Tensor:
            assert lengths.ndim == 1, lengths.ndim
            max_len = max(max_len, int(lengths.max()))
            n = lengths.size(0)
            seq = torch.arange(0, max_len, device=lengths.device)
            return seq.unsqueeze(0).expand(n, max_len) >= lengths.unsqueeze(-1)

        class AttributeDict(dict):
            def __getattr__(self, k):
                if k in self:
                    return self[k]
                raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v

        def str2bool(v):
            if isinstance(v, bool):
                return v
            return str(v).lower() in ("yes", "true", "t", "y", "1")

        iu.make_pad_mask = make_pad_mask
        iu.AttributeDict = AttributeDict
        iu.str2bool = str2bool

    # --- torchmetrics.classification -----------------------------------
    if "torchmetrics" not in sys.modules:
        tm = _mod("torchmetrics")
        tmc = _mod("torchmetrics.classification")
        tm.classification = tmc

        class MulticlassAccuracy(nn.Module):
            """top-k micro accuracy with ignore_index (torchmetrics semantics for
            the arguments valle.py:157-163 uses).  Metric only; not on the token path."""

            def __init__(self, num_classes, top_k=1, average="micro",
                         multidim_average="global", ignore_index=None):
                super().__init__()
                self.top_k = top_k
                self.ignore_index = ignore_index

            def forward(self, logits, target):
                # logits [N, C, T], target [N, T]
                topk = logits.topk(self.top_k, dim=1).indices  # [N,k,T]
                hit = (topk == target.unsqueeze(1)).any(dim=1)
                if self.ignore_index is not None:
                    keep = target != self.ignore_index
                    hit = hit & keep
                    denom = keep.sum().clamp(min=1)
                else:
                    denom = torch.tensor(target.numel())
                return hit.sum().float() / denom.float()

        tmc.MulticlassAccuracy = MulticlassAccuracy
        tmc.BinaryAccuracy = MulticlassAccuracy


def load_reference():
    """Return a namespace with the reference's own classes:
    .VALLE .VALLF .topk_sampling .TransformerEncoder ... (unmodified code)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if _PREFIX + ".models.valle" in sys.modules:
        return _namespace()

    _install_stubs()
    import torch.nn as nn

    root = os.path.join(REFERENCE_ROOT, "valle")
    pkg = _mod(_PREFIX)
    pkg.__path__ = []  # mark as package
    for sub in ("utils", "modules", "models", "data"):
        m = _mod(f"{_PREFIX}.{sub}")
        m.__path__ = []
        setattr(pkg, sub, m)

    # valle.utils.{Transpose, make_pad_mask} (valle/utils/__init__.py re-exports)
    utils = sys.modules[f"{_PREFIX}.utils"]

    class Transpose(nn.Identity):
        def forward(self, input):
            return input.transpose(1, 2)

    utils.Transpose = Transpose
    utils.make_pad_mask = sys.modules["icefall.utils"].make_pad_mask

    # valle.data.input_strategies.PromptedFeatures (isinstance check valle.py:793)
    data_is = _mod(f"{_PREFIX}.data.input_strategies")

    class PromptedFeatures:
        def __init__(self, prompts, features):
            self.prompts = prompts
            self.features = features

        @property
        def data(self):
            return (self.prompts, self.features)

    data_is.PromptedFeatures = PromptedFeatures
    sys.modules[f"{_PREFIX}.data"].input_strategies = data_is

    # valle.models.visualizer.visualize -> no-op (real one needs matplotlib)
    vis = _mod(f"{_PREFIX}.models.visualizer")
    vis.visualize = lambda *a, **k: None

    # The reference uses absolute `valle.*` imports in models/valle.py; alias the
    # private package under the public name only while exec'ing, then restore.
    saved = {k: v for k, v in sys.modules.items() if k == "valle" or k.startswith("valle.")}
    for k in list(saved):
        del sys.modules[k]
    try:
        for k, v in list(sys.modules.items()):
            if k == _PREFIX or k.startswith(_PREFIX + "."):
                sys.modules["valle" + k[len(_PREFIX):]] = v

        def _load(modname: str, relpath: str):
            full = f"{_PREFIX}.{modname}"
            spec = importlib.util.spec_from_file_location(
                "valle." + modname, os.path.join(root, relpath))
            m = importlib.util.module_from_spec(spec)
            sys.modules["valle." + modname] = m
            sys.modules[full] = m
            spec.loader.exec_module(m)
            parent, _, leaf = modname.rpartition(".")
            setattr(sys.modules[f"{_PREFIX}.{parent}"], leaf, m)
            return m

        _load("modules.scaling", "modules/scaling.py")
        _load("modules.activation", "modules/activation.py")
        _load("modules.embedding", "modules/embedding.py")
        _load("modules.transformer", "modules/transformer.py")
        _load("models.macros", "models/macros.py")
        _load("models.valle", "models/valle.py")
    finally:
        for k in [k for k in sys.modules if k == "valle" or k.startswith("valle.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return _namespace()


def _namespace():
    ns = types.SimpleNamespace()
    mv = sys.modules[f"{_PREFIX}.models.valle"]
    ns.VALLE = mv.VALLE
    ns.VALLF = mv.VALLF
    ns.topk_sampling = mv.topk_sampling
    ns.top_k_top_p_filtering = mv.top_k_top_p_filtering
    ns.valle_module = mv
    ns.transformer = sys.modules[f"{_PREFIX}.modules.transformer"]
    ns.activation = sys.modules[f"{_PREFIX}.modules.activation"]
    ns.embedding = sys.modules[f"{_PREFIX}.modules.embedding"]
    ns.PromptedFeatures = sys.modules[f"{_PREFIX}.data.input_strategies"].PromptedFeatures
    return ns
