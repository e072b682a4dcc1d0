"""TEST INFRASTRUCTURE ONLY -- long-horizon training trajectories of the UNMODIFIED reference (`GPMDM.train_adam`,
gpmdm.py:817-885) on seeded synthetic sequences, written to tests/golden/ref_train_trajectory_<name>.npz: the loss of every
Adam step, the trained hyper-parameters and the trained latents.  Run in the build container (needs /root/reference):
    python -m oracle.make_train_trajectory [name ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from gpmdm_b200 import synthetic  # noqa: E402

# name: (C, d, D, sequences per class, frames, Adam steps, synthetic observation noise, seed, sigma_n init)
CASES = {
    "n360_300steps": (2, 3, 35, 3, 60, 300, 0.3, 41, 1e-2),
    # the reference's published training shape (notebooks/train_gpmdm.ipynb: d = 4, D = 35, 19 sequences, ~2 000 frames)
    "n1995_200steps": (2, 4, 35, (10, 9), 105, 200, 0.3, 3, 1e-2),
}


def build(model_cls, cfg):
    C, d, D, spc, frames, steps, noise, seed, sigma_n = cfg
    per_class = list(spc) if isinstance(spc, tuple) else [spc] * C
    wl = synthetic.make_sequences(C, D, max(per_class), frames, seed=seed, n_test_trials=1, test_frames=4, noise=noise)
    hp = synthetic.notebook_hyperparameters(D, d, sigma_n)
    m = model_cls(D=D, d=d, n_classes=C, dyn_target="full", dyn_back_step=1, **hp)
    for c in range(C):
        for s in wl.sequences[c][:per_class[c]]:
            m.add_data(s, c)
    m.init_X()
    return m


def main():
    ref = ref_shim.load_reference()
    torch.set_num_threads(8)
    for name in (sys.argv[1:] or list(CASES)):
        cfg = CASES[name]
        m = build(ref.GPMDM, cfg)
        X0 = m.X.detach().numpy().copy()
        t0 = time.time()
        losses = m.train_adam(cfg[5], 0, lr=0.01)
        wall = time.time() - t0
        path = os.path.join(ROOT, "tests", "golden", f"ref_train_trajectory_{name}.npz")
        state = {k: v.detach().numpy().copy() for k, v in m.state_dict().items() if k != "X"}
        np.savez_compressed(path, losses=np.array(losses), X0=X0, X=m.X.detach().numpy(), wall_s=np.array(wall),
                            threads=np.array(torch.get_num_threads()), **{"p_" + k: v for k, v in state.items()})
        print(f"wrote {path}: {os.path.getsize(path) / 1e3:.0f} kB, {cfg[5]} steps in {wall:.1f} s on {torch.get_num_threads()} threads")


if __name__ == "__main__":
    main()
