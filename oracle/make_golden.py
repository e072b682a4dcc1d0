"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

For each case: build a reference `GPMDM` on seeded synthetic sequences (gpmdm_b200/synthetic.py),
`init_X()` (+ optionally a few `train_adam` steps), construct the reference `GPMDM_PF` with injected
initial indices, then run `update(z)` for a few frames with injected raw draws
(`oracle/ref_shim.InjectedDraws`), recording every stage output of the reference itself.
The reference sources are imported from where they lie; nothing is copied into this repository.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from gpmdm_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (C, d, D, seqs_per_class, frames, P, steps, sigma_n, adam_steps, seed)
    "c2_d3_n240_p64": (2, 3, 62, 2, 60, 64, 3, 1e-2, 0, 11),
    "c3_d3_n270_p96_trained": (3, 3, 62, 3, 30, 96, 3, 1e-1, 8, 12),
    "c2_d4_n200_p50": (2, 4, 35, 2, 50, 50, 2, 1e-2, 0, 13),  # the notebook's d=4, D=35 shape; P % C == 0
    "c3_d3_n180_p64_ragged": (3, 3, 20, 2, 30, 64, 2, 1e-1, 0, 14),  # P % C != 0
}


def build_reference_model(ref, C, d, D, seqs_per_class, frames, sigma_n, adam_steps, seed):
    wl = synthetic.make_sequences(C, D, seqs_per_class, frames, seed=seed, n_test_trials=1, test_frames=8)
    hp = synthetic.notebook_hyperparameters(D, d, sigma_n)
    model = ref.GPMDM(D=D, d=d, n_classes=C, dyn_target="full", dyn_back_step=1, **hp)
    for c in range(C):
        for s in wl.sequences[c]:
            model.add_data(s, c)
    model.init_X()
    if adam_steps:
        torch.manual_seed(seed)
        model.train_adam(adam_steps, 0, lr=0.01)
    return model, wl


# The reference's own operating point (notebooks/test_gpmdm_pf.ipynb: 100 particles; BASELINE configs[0] shape) at
# N_train = 2 000.  Compact fixture: the reference's inputs and stage outputs only -- no inverses (32 MB) and no Y (regenerated
# from the seeded generator, pinned by its sha256) -- so the CUDA path is compared with the reference directly, using its
# own factors.
SCALE_CASES = {
    "scale_cfg1_n2000_p100": (2, 3, 62, 10, 100, 100, 3, 1e-1, 0, 21),
}


def run_case(name, cfg, compact=False):
    C, d, D, spc, frames, P, steps, sigma_n, adam_steps, seed = cfg
    ref = ref_shim.load_reference()
    import gpmdm.gpmdm_pf as ref_pf_module  # the reference module (resolved via ref_shim's sys.path)

    model, wl = build_reference_model(ref, C, d, D, spc, frames, sigma_n, adam_steps, seed)
    out = {}
    with torch.no_grad():
        model.set_evaluation_mode()
        out["X"] = model.X.detach().numpy().copy()
        out["Y"] = np.concatenate(model.observations_list, 0)  # float32, as fed to add_data
        out["seq_lengths"] = np.array([[len(s) for s in cls] for cls in model.class_aware_observations_list])
        for k in ("y_log_lengthscales", "y_log_lambdas", "y_log_sigma_n", "x_log_lengthscales",
                  "x_log_lambdas", "x_log_sigma_n", "x_log_lin_coeff"):
            out[k] = getattr(model, k).detach().numpy().copy()
        if compact:
            import hashlib

            out["Y_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(out["Y"]).tobytes()).digest(), dtype=np.uint8)
            out["gen_cfg"] = np.array([C, D, spc, frames, seed])  # synthetic.make_sequences(C, D, spc, frames, seed=seed, 1, 8)
            del out["Y"]
        else:
            out["Ky_inv"] = model.Ky_inv.detach().numpy().copy()
        # diagonal blocks of the reference's dense per-class inverses (+ check the off-block claim)
        s = 0
        for c in range(0 if compact else C):
            n = sum(len(q) - 1 for q in model.class_aware_observations_list[c])
            full = model.Kx_inv_class[c].detach()
            out[f"Kx_inv_block_{c}"] = full[s:s + n, s:s + n].numpy().copy()
            off = full.clone()
            off[s:s + n, s:s + n] = 0
            Nx = full.shape[0]
            expect = 1e6 * torch.eye(Nx, dtype=full.dtype)
            expect[s:s + n, s:s + n] = 0
            assert torch.allclose(off, expect, rtol=1e-12, atol=0), "off-class part is not 1e6*I"
            s += n
        T = synthetic.markov_matrix(C)
        out["T_f32"] = T.numpy().copy()
        parts = [P // C + (1 if i < P % C else 0) for i in range(C)]
        g = torch.Generator().manual_seed(seed + 100)
        init_idx = []
        for c in range(C):
            ncls = sum(len(q) for q in model.class_aware_observations_list[c])
            init_idx.append(torch.randint(0, ncls, (parts[c],), generator=g))
        for c in range(C):
            out[f"init_idx_{c}"] = init_idx[c].numpy().copy()
        with ref_shim.InjectedDraws(ref_pf_module, init_idx=init_idx):
            pf = ref.GPMDM_PF(model, T, P)
        out["init_states"] = pf._particle_states.numpy().copy()
        out["init_classes"] = pf._particle_classes.numpy().copy()
        cls_true, trial = wl.test_trials[0]
        rec_obs = {}
        orig_map = model.map_x_to_y

        def recording_map(Xstar, flg_noise=False):
            mean, var = orig_map(Xstar, flg_noise)
            rec_obs["mu"], rec_obs["var"] = mean.clone(), var.clone()
            return mean, var

        model.map_x_to_y = recording_map
        for t in range(steps):
            E, eps, u = synthetic.raw_draws(P, C, d, seed * 1000 + t)
            z = trial[t]
            with ref_shim.InjectedDraws(ref_pf_module, E=E, eps=eps, u=u) as inj:
                pf.update(z)
            dyn_mean = torch.zeros(P, d, dtype=torch.float64)
            dyn_std = torch.zeros(P, d, dtype=torch.float64)
            for c, (rows, val) in inj.record.get("dyn_mean", {}).items():
                dyn_mean[rows] = val
            for c, (rows, val) in inj.record.get("dyn_std", {}).items():
                dyn_std[rows] = val
            pre = f"s{t}_"
            out[pre + "z"] = np.asarray(z)
            out[pre + "E"], out[pre + "eps"], out[pre + "u"] = E.numpy(), eps.numpy(), u.numpy()
            out[pre + "c_new"] = inj.record["new_classes"].numpy()
            out[pre + "dyn_mean"] = dyn_mean.numpy()
            out[pre + "dyn_var"] = (dyn_std ** 2).numpy()  # NB sqrt then square: ~1 ulp from var
            out[pre + "dyn_std"] = dyn_std.numpy()
            out[pre + "mu"] = rec_obs["mu"].numpy()
            out[pre + "var"] = rec_obs["var"].numpy()
            out[pre + "ll"] = pf._log_likelihoods.numpy().copy()
            out[pre + "lw"] = pf._log_weights.numpy().copy()
            out[pre + "w"] = pf._weights.numpy().copy()
            out[pre + "anc"] = inj.record["ancestors"].numpy()
            out[pre + "cdf"] = inj.record["cdf"].numpy()
            out[pre + "states_post"] = pf._particle_states.numpy().copy()
            out[pre + "classes_post"] = pf._particle_classes.numpy().copy()
            out[pre + "class_prob"] = pf.class_probabilities().numpy().copy()
            out[pre + "argmax"] = np.array(pf.get_most_likely_class())
            out[pre + "state_mean"] = pf.current_state_mean().numpy().copy()
            out[pre + "log_likelihood"] = np.array(pf.log_likelihood())
        out["steps"] = np.array(steps)
        out["P"] = np.array(P)
        out["true_class"] = np.array(cls_true)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e3:.0f} kB")


def main():
    torch.set_num_threads(8)
    for name, cfg in CASES.items():
        run_case(name, cfg)
    for name, cfg in SCALE_CASES.items():
        run_case(name, cfg, compact=True)


if __name__ == "__main__":
    main()
