"""GPU: how sparse is the cross-kernel K* of the observation GP for the particle tiles the filter actually produces?

Runs the benchmark filter (BASELINE configs[2] model by default) for a few frames and, for the states the observation kernel
sees at each frame, computes per (64-particle tile, 16-row chunk of training rows) the largest K* entry.  A chunk whose
entries are all below `thr` contributes less than thr * max|K^-1| to k^T K^-1 k and to the means and could be skipped; a
256-column tile all of whose chunks are below it drops out of the quadratic form entirely.  Prints one JSON line with the
live fractions and the implied share of the triangular (k-chunk, column-tile) work, in the filter's own particle order and
with the particles additionally sorted along the first latent coordinate inside each class, and with BOTH the training
rows and the particles in Z-order (Morton code of the scaled latents).
   python tools/kstar_sparsity.py [--frames 6] [--particles 37888]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from gpmdm_b200 import GPMDM_PF, synthetic


def chunk_max(A, B):
    """A [N, d] training rows / l, B [P, d] particles / l (P multiple of 64, N multiple of 16) -> [P/64, N/16] max of K*."""
    out = []
    for s in range(0, B.shape[0], 4096):
        Bs = B[s:s + 4096]
        K = torch.exp(-torch.cdist(Bs, A) ** 2)
        out.append(K.view(Bs.shape[0] // 64, 64, A.shape[0] // 16, 16).amax(dim=(1, 3)))
    return torch.cat(out)


def work_share(live):
    """live [tiles, chunks] bool -> (live chunk fraction, live column-tile fraction, share of the triangular work left)."""
    tiles, nch = live.shape
    nct = nch // 16
    ct_live = live[:, :nct * 16].view(tiles, nct, 16).any(dim=2)                        # [tiles, nct]
    cum = torch.cumsum(live[:, :nct * 16].to(torch.float64), 1).view(tiles, nct, 16)[:, :, -1]   # live chunks k < 16 (ct + 1)
    done = (cum * ct_live).sum()
    total = tiles * sum(16 * (c + 1) for c in range(nct))
    return float(live.double().mean()), float(ct_live.double().mean()), float(done / total)


def morton_key(Z, bits=10):
    """Morton (Z-order) code of the first three coordinates of Z [n, d], quantised to `bits` bits each."""
    Z = Z[:, :3]
    lo, hi = Z.min(0).values, Z.max(0).values
    q = ((Z - lo) / (hi - lo + 1e-300) * (2 ** bits - 1)).long().clamp(0, 2 ** bits - 1)
    key = torch.zeros(Z.shape[0], dtype=torch.long, device=Z.device)
    for b in range(bits):
        for c in range(q.shape[1]):
            key |= ((q[:, c] >> b) & 1) << (3 * b + c)
    return key


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--classes", type=int, default=8)
    ap.add_argument("--seqs-per-class", type=int, default=10)
    ap.add_argument("--frames", type=int, default=250)
    ap.add_argument("--latent", type=int, default=3)
    ap.add_argument("--particles", type=int, default=37888)
    ap.add_argument("--steps", type=int, default=6)
    o = ap.parse_args()
    a = argparse.Namespace(classes=o.classes, seqs_per_class=o.seqs_per_class, frames=o.frames, latent=o.latent, obs_dim=62)
    wl, X0, hp = bench.synthetic_inputs(a)
    model = bench.build_product_model(a, wl, X0, hp)
    N = X0.shape[0] // 16 * 16
    ls = torch.exp(model.y_log_lengthscales.detach()).cuda()
    A = (torch.tensor(X0[:N]).cuda() / ls).contiguous()
    pf = GPMDM_PF(model, synthetic.markov_matrix(a.classes), o.particles, seed=1234, cdf_order="blocked")
    trial = wl.test_trials[0][1]
    rows = []
    for t in range(o.steps):
        pf.update(torch.tensor(trial[t]))
        x = pf.last_pre_resample_states.clone()
        c = pf.last_pre_resample_classes.clone()
        P = x.shape[0] // 64 * 64
        B = (x[:P] / ls).contiguous()
        key = c[:P].double() * 1e6 + B[:, 0]      # class-major, then along the first latent coordinate
        Bs = B[torch.argsort(key)]
        row = {"frame": t}
        # both sides in Z-order: training rows by the Morton code of their scaled latents, particles class-major then Morton
        Am = A[torch.argsort(morton_key(A))]
        lo, hi = A.min(0).values, A.max(0).values
        Bq = torch.max(torch.min(B, hi), lo)
        Bm = B[torch.argsort(c[:P] * (1 << 40) + morton_key(torch.cat([Bq, lo[None], hi[None]]))[:P])]
        for name, Ax, Bx in (("filter_order", A, B), ("sorted", A, Bs), ("both_morton_sorted", Am, Bm)):
            cm = chunk_max(Ax, Bx)
            row[name] = {str(thr): dict(zip(("chunks_live", "column_tiles_live", "work_share"), work_share(cm >= thr)))
                         for thr in (1e-30, 1e-20)}
        d2 = torch.cdist(B[:4096], A) ** 2
        row["entries_below_1e-30"] = float((torch.exp(-d2) < 1e-30).double().mean())
        rows.append(row)
    print(json.dumps({"workload": f"C={a.classes} N={X0.shape[0]} d={a.latent} P={o.particles}, observation GP, tiles of 64 "
                      "particles x chunks of 16 training rows", "frames": rows}))


if __name__ == "__main__":
    main()
