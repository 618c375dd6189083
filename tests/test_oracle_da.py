"""oracle/da.py (restatement of Clip_Discriminator + GRL + clip-level CDAN loss) against the fixture produced by the
reference's own modules (tests/golden/ada.npz, tests/make_golden_ada.py), and against the live reference if present."""
import numpy as np
import torch

from conftest import has_reference
from helpers import golden, rel_l2
from oracle import da as oda


def _run_oracle():
    oc = oda.OracleClipDiscriminator()
    oda.seeded_disc_init(oc, 3)
    oc.train()
    f_s = oda.seeded_features(2, 31).requires_grad_(True)
    f_t = oda.seeded_features(3, 32).requires_grad_(True)
    loss, d = oda.cdan_clip_loss(oc, f_s, f_t, 500)
    loss.backward()
    return oc, f_s, f_t, loss, d


def test_restatement_matches_reference_fixture():
    g = golden("ada.npz")
    oc, f_s, f_t, loss, d = _run_oracle()
    assert [str(k) for k in g["keys"]] == list(oc.state_dict().keys())
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    assert np.abs(d.detach().numpy() - g["p_train"]).max() < 1e-6
    assert abs(oda.grl_coeff(500) - float(g["coeff"])) < 1e-12 and int(g["iter_after"]) == 501
    assert rel_l2(f_s.grad.numpy()[:, ::7, ::5], g["df_s"]) < 1e-4 and rel_l2(f_t.grad.numpy()[:, ::7, ::5], g["df_t"]) < 1e-4
    for name, p in oc.named_parameters():
        ref = g["g_" + name]
        got = p.grad.reshape(-1).numpy()
        got = got[:: max(1, got.size // 2048)][:2048]
        if not (name.startswith("conv_") and name.endswith(".bias")):   # conv biases ahead of BatchNorm: rounding noise
            assert rel_l2(got, ref) < 1e-3, name
    for k, v in oc.state_dict().items():
        if "running" in k:
            assert np.abs(v.numpy() - g["s_" + k]).max() < 1e-6, k


def test_grl_coefficient_schedule():
    assert oda.grl_coeff(0) == 0.0
    assert abs(oda.grl_coeff(1000) - (2 / (1 + np.exp(-1.0)) - 1)) < 1e-12
    assert 0.99 < oda.grl_coeff(10 ** 6) <= 1.0


def test_against_live_reference_when_present():
    if not has_reference():
        import pytest
        pytest.skip("reference tree not mounted")
    import sys
    sys.path.insert(0, "/root/reference/src")
    np.float = float
    from models.CRNN_GRL import Clip_Discriminator
    oc, f_s, f_t, loss, d = _run_oracle()
    ref = Clip_Discriminator(256)
    oc2 = oda.OracleClipDiscriminator()
    oda.seeded_disc_init(oc2, 3)
    ref.load_state_dict(oc2.state_dict())
    ref.train()
    with torch.no_grad():
        p = ref(torch.cat((f_s, f_t))).reshape(-1)
    assert np.abs(p.numpy() - d.detach().numpy()).max() < 1e-6
