"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/ref_saved_model.pth (+ expectations) with the UNMODIFIED reference:
a small trained model saved by the reference's own `GPMDM.save` (gpmdm.py:1307-1346), and the reference's predictions
on a few query points, so that `gpmdm_b200.GPMDM.load` can be checked against a file the reference wrote.
Run in the build container:   python -m oracle.make_saved_model"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from gpmdm_b200 import synthetic  # noqa: E402


def main():
    ref = ref_shim.load_reference()
    wl = synthetic.make_sequences(2, 12, 2, 30, seed=31, n_test_trials=1, test_frames=4)
    hp = synthetic.notebook_hyperparameters(12, 3, 1e-1)
    m = ref.GPMDM(D=12, d=3, n_classes=2, dyn_target="full", dyn_back_step=1, **hp)
    for c in range(2):
        for s in wl.sequences[c]:
            m.add_data(s, c)
    m.init_X()
    X0 = m.X.detach().numpy().copy()
    torch.manual_seed(0)
    losses = m.train_adam(5, 0, lr=0.01)
    out = os.path.join(ROOT, "tests", "golden", "ref_saved_model.pth")
    m.save(out)
    with torch.no_grad():
        m.set_evaluation_mode()
        g = torch.Generator().manual_seed(1)
        xs = m.X.detach()[torch.randint(0, m.X.shape[0], (16,), generator=g)] + 0.1 * torch.randn(16, 3, dtype=torch.float64, generator=g)
        mu, var = m.map_x_to_y(xs)
        dm, dv = m.map_x_dynamics_for_class(xs, 1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_saved_model_expect.npz"), xs=xs.numpy(), mu=mu.numpy(),
                        var=var.numpy(), dyn_mean=dm.numpy(), dyn_var=dv.numpy(), losses=np.array(losses), X0=X0,
                        state={k: v.numpy() for k, v in m.state_dict().items()})
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
