"""Trial driver: the evaluation loop of the reference's `notebooks/test_gpmdm_pf.ipynb` (cells 3-5) on seeded
synthetic trials -- per-frame / per-trial accuracy, macro F1 and frames per second, timed over the same region the
notebook times (`update` + `get_most_likely_class` + `class_probabilities`).

    python tools/run_trials.py [--particles 100] [--trials 20] [--frames 150] [--classes 2] [--train-steps 0]

Default = BASELINE config 1 ("README setup"): 2-class model, d = 3, D = 62, ~2k training frames, 100 particles,
150-frame test trials.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def f1_macro(y_true, y_pred, C):
    out = []
    for c in range(C):
        tp = np.sum((y_true == c) & (y_pred == c)); fp = np.sum((y_true != c) & (y_pred == c)); fn = np.sum((y_true == c) & (y_pred != c))
        out.append(2 * tp / max(2 * tp + fp + fn, 1))
    return float(np.mean(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=100)
    ap.add_argument("--trials", type=int, default=20)
    ap.add_argument("--frames", type=int, default=150)
    ap.add_argument("--classes", type=int, default=2)
    ap.add_argument("--seqs-per-class", type=int, default=10)
    ap.add_argument("--train-frames", type=int, default=100)
    ap.add_argument("--train-steps", type=int, default=0)
    ap.add_argument("--precision", default="fp64")
    ap.add_argument("--no-graph", action="store_true", help="launch the small-cloud step directly instead of replaying a CUDA graph")
    ap.add_argument("--staged", action="store_true", help="the stage-by-stage launch sequence (round-1 path)")
    ap.add_argument("--batched", action="store_true", help="one update_many() call per trial instead of the per-frame loop")
    a = ap.parse_args()
    from gpmdm_b200 import GPMDM, GPMDM_PF, synthetic

    C, d, D = a.classes, 3, 62
    wl = synthetic.make_sequences(C, D, a.seqs_per_class, a.train_frames, seed=0, n_test_trials=a.trials, test_frames=a.frames)
    model = GPMDM(D=D, d=d, n_classes=C, dyn_target="full", dyn_back_step=1, **synthetic.notebook_hyperparameters(D, d, 1e-1))
    for c in range(C):
        for s in wl.sequences[c]:
            model.add_data(s, c)
    model.init_X()
    if a.train_steps:
        model.train_adam(a.train_steps, 0, lr=0.01)
    T = synthetic.markov_matrix(C)
    pf = GPMDM_PF(model, T, a.particles, seed=0, precision=a.precision, cuda_graph=not a.no_graph,
                  native_step=not a.staged)
    for z in wl.test_trials[0][1][:4]:  # warm-up (kernel attributes, allocator, graph capture)
        pf.update(z)
    pf.reset()
    frame_true, frame_pred, trial_true, trial_pred, secs, frames = [], [], [], [], 0.0, 0
    for cls, trial in wl.test_trials:
        pf.reset()
        votes = np.zeros(C)
        if a.batched:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            probs, preds, _ = pf.update_many(trial)
            preds = preds.cpu().numpy()
            torch.cuda.synchronize()
            secs += time.perf_counter() - t0
            frames += len(trial)
            for pred in preds:
                frame_true.append(cls); frame_pred.append(int(pred)); votes[int(pred)] += 1
            trial_true.append(cls); trial_pred.append(int(np.argmax(votes)))
            continue
        for z in trial:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pf.update(z)
            pred = pf.get_most_likely_class()
            probs = pf.class_probabilities()
            torch.cuda.synchronize()
            secs += time.perf_counter() - t0
            frames += 1
            frame_true.append(cls); frame_pred.append(pred); votes[pred] += 1
        trial_true.append(cls); trial_pred.append(int(np.argmax(votes)))
    ft, fp_, tt, tp_ = map(np.array, (frame_true, frame_pred, trial_true, trial_pred))
    print(json.dumps({
        "workload": f"{C}-class GPMDM, d={d}, D={D}, N_train={C * a.seqs_per_class * a.train_frames}, P={a.particles}, "
                    f"{a.trials} synthetic trials x {a.frames} frames, precision={a.precision}",
        "frame_accuracy": float(np.mean(ft == fp_)), "frame_f1": f1_macro(ft, fp_, C),
        "trial_accuracy": float(np.mean(tt == tp_)), "trial_f1": f1_macro(tt, tp_, C),
        "driver": "update_many(trial): one call per trial" if a.batched else "per-frame loop: update, get_most_likely_class, class_probabilities",
        "step_path": ("small-cloud kernels, " + ("CUDA graph replay" if pf._use_graph else "direct launches")) if pf._small
                     else ("native launch sequence" if pf._native_step else "staged launches"),
        "launches_per_step": pf.launches_per_step,
        "seconds_per_frame": secs / frames, "fps": frames / secs, "particle_updates_per_sec": a.particles * frames / secs,
        "reference_published": {"fps": 12.78, "trial_accuracy": 0.974, "frame_accuracy": 0.921,
                                "note": "laptop CPU, CMU mocap (not in tree), P=100 -- test_gpmdm_pf.ipynb"},
    }))


if __name__ == "__main__":
    main()
