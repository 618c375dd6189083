"""Generate tests/golden/ada.npz by EXECUTING THE REFERENCE's Clip_Discriminator + cdan_frame loss (imported from
/root/reference/src) on seeded inputs.      python tests/make_golden_ada.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
np.float = float   # the reference's DA/grl.py:64 uses np.float (removed in NumPy 1.24)

from oracle import da as oda  # noqa: E402


def main():
    from models.CRNN_GRL import Clip_Discriminator
    from DA.cdan_frame import ConditionalDomainAdversarialLoss
    torch.manual_seed(0)
    ref = Clip_Discriminator(256)
    oc = oda.OracleClipDiscriminator()
    oda.seeded_disc_init(oc, 3)
    ref.load_state_dict(oc.state_dict())
    ref.train()
    f_s = oda.seeded_features(2, 31).requires_grad_(True)
    f_t = oda.seeded_features(3, 32).requires_grad_(True)
    loss_mod = ConditionalDomainAdversarialLoss(ref, entropy_conditioning=False, randomized=False, reduction='mean')
    loss_mod.grl.iter_num = 500
    g_s, g_t = torch.rand(2, 313, 20), torch.rand(3, 313, 20)
    loss = loss_mod(g_s, f_s, g_t, f_t)
    loss.backward()
    with torch.no_grad():
        ref.eval()
        p_eval = ref(torch.cat((f_s, f_t))).reshape(-1)
        ref.train()
    rec = {"loss": float(loss), "coeff": oda.grl_coeff(500), "iter_after": loss_mod.grl.iter_num,
           "df_s": f_s.grad.numpy()[:, ::7, ::5], "df_t": f_t.grad.numpy()[:, ::7, ::5],
           "df_s_norm": float(f_s.grad.norm()), "df_t_norm": float(f_t.grad.norm()), "p_eval": p_eval.numpy()}
    # train-mode probabilities (recomputed: the loss module does not return them)
    ref2 = Clip_Discriminator(256)
    ref2.load_state_dict(oc.state_dict())
    ref2.train()
    with torch.no_grad():
        rec["p_train"] = ref2(torch.cat((f_s, f_t))).reshape(-1).numpy()
    for name, p in ref.named_parameters():
        g = p.grad.reshape(-1).numpy()
        rec["g_" + name] = g[:: max(1, g.size // 2048)][:2048]
        rec["gn_" + name] = float(np.linalg.norm(g))
    for k, v in ref.state_dict().items():
        if "running" in k or "num_batches" in k:
            rec["s_" + k] = v.numpy()
    rec["keys"] = np.array(list(ref.state_dict().keys()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ada.npz"), **rec)
    print("wrote ada.npz: loss", float(loss), "p_train", rec["p_train"])


if __name__ == "__main__":
    main()
