"""GPU: the reference's own probe head (avex/models/probes/linear_probe.py:LinearProbe, base_probes.py:197-206, :300-323) on top of
the drop-in BEATs model -- SURVEY 8 row f2.  The probe stays stock torch; what this path provides is the layer-wise features:
device-pooled `[B,768]` per hooked layer for `aggregation="mean"`, frame-level hooks for the list path with its softmax-weighted
layer sum.  Expectations are formed from the REFERENCE's hook outputs (tests/golden) pushed through the same probe head.
The avex package is imported from baseline/_ref (installed there by the reference-arm recipe; it travels with the snapshot)."""
import os
import sys

import numpy as np
import pytest
import torch

from tests.golden import cases
from tests.test_beats_gpu import _build

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _linear_probe_cls():
    sys.path.insert(0, ROOT)
    import bench

    if bench.import_reference() is None:
        pytest.skip("the reference package is not installed under baseline/_ref")
    from avex.models.probes.linear_probe import LinearProbe

    return LinearProbe


def test_reference_linear_probe_on_layerwise_features():
    LinearProbe = _linear_probe_cls()
    cname = "L2_2x1s"
    case = cases.beats_cases()[cname]
    g = np.load(os.path.join(G, f"beats_{cname}.npz"))
    model, _ = _build(case["layers"], case["wseed"])
    model.register_hooks_for_layers(["all"])
    wav = torch.from_numpy(case["wav"]).cuda()
    ref_hooks = [torch.from_numpy(g[f"hook{i}"]).cuda() for i in range(3)]  # the reference's [B,N,768] per hooked layer

    # aggregation="mean": one [B, 3*768] tensor, each third formed in an fc2 / projection epilogue on the device
    torch.manual_seed(0)
    probe = LinearProbe(base_model=model, layers=["all"], num_classes=7, device="cuda", aggregation="mean", target_length=16000)
    with torch.no_grad():
        logits = probe(wav)
        want = probe.classifier(torch.cat([h.mean(dim=1) for h in ref_hooks], dim=1))
    assert logits.shape == (2, 7)
    scale = max(1.0, want.abs().max().item())
    assert (logits - want).abs().max().item() <= 5e-3 * scale, (logits - want).abs().max().item()

    # list path: frame-level hooks -> per-layer projectors -> softmax(layer_weights)-weighted sum -> head
    probe2 = LinearProbe(base_model=model, layers=["all"], num_classes=7, device="cuda", aggregation="none", target_length=16000)
    assert hasattr(probe2, "layer_weights") and probe2.layer_weights.numel() == 3
    with torch.no_grad():
        probe2.layer_weights.copy_(torch.tensor([0.3, -0.2, 0.5], device="cuda"))
        logits2 = probe2(wav)
        want2 = probe2.classifier(probe2._combine_or_reshape_embeddings(ref_hooks))
    scale2 = max(1.0, want2.abs().max().item())
    assert (logits2 - want2).abs().max().item() <= 2e-2 * scale2, (logits2 - want2).abs().max().item()
