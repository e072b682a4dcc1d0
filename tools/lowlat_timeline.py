"""Diagnostic: per-item timeline of the low-latency (split) instance of gp_predict_kernel at the reference's operating
point (100 particles, N_train = 2 000 by default).

  python tools/lowlat_timeline.py --build     (here: compiles csrc/gp_predict.cu with -DGPMDM_TIMELINE into
                                               gpmdm_b200/lib/libgpmdm_sm100a_timeline.so; travels to the GPU box)
  python tools/lowlat_timeline.py             (GPU: runs the observation launch a few times, prints one JSON line)

Stamps (globaltimer, ns) per CTA and work item: loop top, item fetched, first chunk landed, k loop done, epilogue done."""
import argparse
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "gpmdm_b200", "lib", "libgpmdm_sm100a_timeline.so")


def build():
    from gpmdm_b200 import build as b

    b.build()
    tl_srcs = ("gp_predict.cu", "pf_small.cu")
    objs = []
    for name in tl_srcs:
        obj = os.path.join(b.LIB_DIR, name[:-3] + "_timeline.o")
        subprocess.check_call([b._nvcc(), *b.NVCC_FLAGS, "-DGPMDM_TIMELINE", "-I", b.INCLUDE, "-c", os.path.join(b.CSRC, name), "-o", obj])
        objs.append(obj)
    objs += [os.path.join(b.LIB_DIR, os.path.basename(s)[:-3] + ".o") for s in b.sources() if os.path.basename(s) not in tl_srcs]
    subprocess.check_call([b._nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    print(LIB)


def frame(o, lib, model, wl):
    """Node timeline of ONE graph-replayed frame of the small-cloud step: start / end of each of the six kernels."""
    import numpy as np
    import torch

    from gpmdm_b200 import GPMDM_PF, synthetic

    lib.gpmdm_debug_node_stamps.argtypes = [ctypes.c_void_p]
    lib.gpmdm_debug_small_stamps.argtypes = [ctypes.c_void_p]
    pf = GPMDM_PF(model, synthetic.markov_matrix(o.classes), o.particles, seed=0)
    trial = wl.test_trials[0][1]
    for z in trial[:20]:
        pf.update(z)
        pf.get_most_likely_class()
    node, small = np.zeros(16, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
    frames = []
    for z in trial[20:40]:
        lib.gpmdm_debug_node_stamps(node.ctypes.data)  # reset
        pf.update(z)
        pf.get_most_likely_class()
        lib.gpmdm_debug_node_stamps(node.ctypes.data)
        lib.gpmdm_debug_small_stamps(small.ctypes.data)
        n, s_ = node.astype(np.int64), small.astype(np.int64)
        t0 = int(s_[0])
        ev = {"pre": (0, int(s_[1]) - t0), "dyn_items": (int(n[2]) - t0, int(n[3]) - t0), "dyn_finalize": (int(n[6]) - t0, int(n[7]) - t0),
              "obs_items": (int(n[0]) - t0, int(n[1]) - t0), "obs_finalize": (int(n[4]) - t0, int(n[5]) - t0),
              "post": (int(s_[2]) - t0, int(s_[3]) - t0)}
        frames.append(ev)
    names = ["pre", "dyn_items", "dyn_finalize", "obs_items", "obs_finalize", "post"]
    med = {k: [float(np.median([f[k][i] for f in frames])) for i in (0, 1)] for k in names}
    out = {"workload": f"small-cloud frame, graph replay, P={o.particles}, N={model.X.shape[0]}: first start / last end of each kernel, ns after "
                       "the pre kernel's start (medians over 20 frames; %globaltimer)",
           "kernels": med,
           "durations_ns": {k: med[k][1] - med[k][0] for k in names},
           "gaps_ns": {f"{a}->{b}": med[b][0] - med[a][1] for a, b in zip(names[:-1], names[1:])},
           "frame_ns": med["post"][1]}
    print(json.dumps(out))


def fused(o, lib, pk, model, X0, wl):
    import numpy as np
    import torch

    from gpmdm_b200 import _cabi

    sms = torch.cuda.get_device_properties(0).multi_processor_count
    P = 64 * sms
    g = torch.Generator().manual_seed(1)
    idx = torch.randint(0, X0.shape[0], (P,), generator=g)
    xs = (torch.tensor(X0[idx.numpy()]) + 0.1 * torch.randn(P, 3, dtype=torch.float64, generator=g)).cuda().contiguous()
    n_pad = pk["obs_n_pad"]
    kws = torch.empty(sms * n_pad * 64, dtype=torch.float64, device="cuda")
    z = torch.tensor(wl.test_trials[0][1][0], dtype=torch.float64, device="cuda")
    ll = torch.empty(P, dtype=torch.float64, device="cuda")
    counter = torch.zeros(4, dtype=torch.int32, device="cuda")
    words = np.zeros(160 * 64 * 8, dtype=np.uint64)
    counts = np.zeros(160, dtype=np.int32)
    for it in range(3):
        _cabi.check(lib.gpmdm_pf_observe_cached_f64(ctypes.byref(pk["obs"]), xs.data_ptr(), P, z.data_ptr(), 0.0, ll.data_ptr(),
                                                    None, None, n_pad, counter.data_ptr(), kws.data_ptr(), kws.numel() * 8,
                                                    _cabi.stream()), "observe_cached")
        torch.cuda.synchronize()
        lib.gpmdm_debug_timeline(words.ctypes.data, counts.ctypes.data)
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
    w = words.reshape(160, 64, 8).astype(np.int64)
    N = X0.shape[0]
    nkc, nq = (N + 15) // 16, n_pad // 256
    chunks = sum(nkc - 16 * J for J in range(nq) if nkc > 16 * J)
    tile_ns = [int(w[b, 0, 4] - w[b, 0, 2]) for b in range(sms) if counts[b] > 0 and w[b, 0, 2] > 0]
    print(json.dumps({"workload": f"observation GP, fused cached launch, N={N} n_pad={n_pad} P={P} (one tile per SM)",
                      "tile_ns_median": float(np.median(tile_ns)), "chunks_of_L_per_tile": chunks,
                      "ns_per_chunk_incl_cache_fill_and_alpha_tile": float(np.median(tile_ns)) / chunks, "sm_clock_mhz_after": clk}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--classes", type=int, default=2)
    ap.add_argument("--seqs-per-class", type=int, default=10)
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--particles", type=int, default=100)
    ap.add_argument("--seg", type=int, default=-1, help="segment length in chunks (-1: the filter's own choice)")
    ap.add_argument("--fused", action="store_true", help="time the fused cached launch on 148 full tiles instead")
    ap.add_argument("--frame", action="store_true", help="node timeline of one graph-replayed small-cloud frame")
    ap.add_argument("--dynamics", action="store_true", help="the dynamics GP's low-latency launch instead of the observation GP's")
    o = ap.parse_args()
    if o.build:
        return build()
    os.environ["GPMDM_LIBRARY"] = LIB
    import numpy as np
    import torch

    import bench
    from gpmdm_b200 import _cabi

    a = argparse.Namespace(classes=o.classes, seqs_per_class=o.seqs_per_class, frames=o.frames, latent=3, obs_dim=62)
    wl, X0, hp = bench.synthetic_inputs(a)
    model = bench.build_product_model(a, wl, X0, hp)
    lib = _cabi.lib()
    lib.gpmdm_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    pk = model.packed_models()
    P = o.particles
    if o.fused:
        return fused(o, lib, pk, model, X0, wl)
    if o.frame:
        return frame(o, lib, model, wl)
    g = torch.Generator().manual_seed(1)
    idx = torch.randint(0, X0.shape[0], (P,), generator=g)
    xs = (torch.tensor(X0[idx.numpy()]) + 0.1 * torch.randn(P, 3, dtype=torch.float64, generator=g)).cuda().contiguous()
    n_pad = pk["obs_n_pad"]
    seg = o.seg if o.seg >= 0 else int(lib.gpmdm_predict_lowlat_pick_segment((P + 63) // 64, n_pad, 256, 1))
    ws = torch.empty(int(lib.gpmdm_predict_lowlat_workspace_bytes(P, n_pad, 62, seg, 1)) // 8 + 1, dtype=torch.float64, device="cuda")
    z = torch.tensor(wl.test_trials[0][1][0], dtype=torch.float64, device="cuda")
    ll = torch.empty(P, dtype=torch.float64, device="cuda")
    counter = torch.zeros(4, dtype=torch.int32, device="cuda")
    words = np.zeros(160 * 64 * 8, dtype=np.uint64)
    counts = np.zeros(160, dtype=np.int32)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    if o.dynamics:
        C = o.classes
        cls = torch.randint(0, C, (P,), generator=g).cuda()
        perm = torch.empty(P, dtype=torch.int32, device="cuda")
        tiles = torch.empty((P // 64 + C + 1) * 4, dtype=torch.int32, device="cuda")
        n_tiles = torch.empty(1, dtype=torch.int32, device="cuda")
        sws = torch.empty(int(lib.gpmdm_workspace_bytes(P, C)) // 8 + 1, dtype=torch.float64, device="cuda")
        _cabi.check(lib.gpmdm_pf_bucket_by_class(cls.data_ptr(), P, C, perm.data_ptr(), tiles.data_ptr(), n_tiles.data_ptr(),
                                                 sws.data_ptr(), _cabi.stream()), "bucket")
        n_pad = pk["dyn_max_n_pad"]
        seg = o.seg if o.seg >= 0 else int(lib.gpmdm_predict_lowlat_pick_segment(max((P + 63) // 64, min(C, P)), n_pad, 256, 1))
        ws = torch.empty(int(lib.gpmdm_predict_lowlat_workspace_bytes(P, n_pad, 3, seg, C)) // 8 + 1, dtype=torch.float64, device="cuda")
        eps = torch.randn(P, 3, dtype=torch.float64, generator=g).cuda()
        x_new = torch.empty(P, 3, dtype=torch.float64, device="cuda")
    for it in range(5):
        ev[0].record()
        if o.dynamics:
            _cabi.check(lib.gpmdm_pf_propagate_lowlat_f64(ctypes.byref(pk["dyn"]), xs.data_ptr(), perm.data_ptr(), tiles.data_ptr(),
                                                          n_tiles.data_ptr(), P, eps.data_ptr(), x_new.data_ptr(), None, None, n_pad,
                                                          seg, counter.data_ptr(), ws.data_ptr(), _cabi.stream()), "propagate_lowlat")
            ev[1].record()
            torch.cuda.synchronize()
            lib.gpmdm_debug_timeline(words.ctypes.data, counts.ctypes.data)
            continue
        _cabi.check(lib.gpmdm_pf_observe_lowlat_f64(ctypes.byref(pk["obs"]), xs.data_ptr(), P, z.data_ptr(), 0.0, None,
                                                    ll.data_ptr(), None, None, n_pad, seg, counter.data_ptr(), ws.data_ptr(),
                                                    _cabi.stream()), "observe_lowlat")
        ev[1].record()
        torch.cuda.synchronize()
        lib.gpmdm_debug_timeline(words.ctypes.data, counts.ctypes.data)
    w = words.reshape(160, 64, 8).astype(np.int64)
    ctas = [b for b in range(160) if counts[b] > 0]
    t_enter = min(int(w[b, 0, 7]) for b in ctas)
    rows = []
    for b in ctas:
        for i in range(min(int(counts[b]), 64)):
            r = w[b, i]
            if r[2] == 0:
                rows.append(dict(cta=b, item=int(r[6]), empty=True, top=int(r[0] - t_enter), fetched=int(r[1] - t_enter)))
            else:
                rows.append(dict(cta=b, item=int(r[6]), empty=False, top=int(r[0] - t_enter), fetched=int(r[1] - t_enter),
                                 first_chunk=int(r[2] - t_enter), loop_done=int(r[3] - t_enter), epi_done=int(r[4] - t_enter)))
    real = [r for r in rows if not r["empty"]]
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
    out = {
        "sm_clock_mhz_after": clk,
        "workload": f"{'dynamics' if o.dynamics else 'observation'} GP, low-latency items, N={X0.shape[0]} n_pad={n_pad} P={P} seg_chunks={seg}",
        "launch_ms_events": ev[0].elapsed_time(ev[1]), "ctas": len(ctas), "items_real": len(real),
        "items_empty": len(rows) - len(real),
        "cta_enter_spread_ns": max(int(w[b, 0, 7]) for b in ctas) - t_enter,
        "last_epilogue_done_ns": max(r["epi_done"] for r in real),
        "last_loop_top_ns": max(r["top"] for r in rows),
        "median": {k: float(np.median([r[k2] - r[k1] for r in real])) for k, k1, k2 in
                   (("fetch_ns", "top", "fetched"), ("prologue_ns", "fetched", "first_chunk"),
                    ("k_loop_ns", "first_chunk", "loop_done"), ("epilogue_ns", "loop_done", "epi_done"))},
        "max": {k: float(np.max([r[k2] - r[k1] for r in real])) for k, k1, k2 in
                (("fetch_ns", "top", "fetched"), ("prologue_ns", "fetched", "first_chunk"),
                 ("k_loop_ns", "first_chunk", "loop_done"), ("epilogue_ns", "loop_done", "epi_done"))},
        "empty_fetch_median_ns": float(np.median([r["fetched"] - r["top"] for r in rows if r["empty"]])) if len(rows) > len(real) else None,
        "per_cta_items": sorted(((b, int(counts[b])) for b in ctas), key=lambda t: -t[1])[:8],
        "slowest_ctas": sorted(({"cta": b, "done": max(r["epi_done"] for r in real if r["cta"] == b),
                                 "items": [(r["item"], r["epi_done"] - r["fetched"]) for r in real if r["cta"] == b]}
                                for b in set(r["cta"] for r in real)), key=lambda t: -t["done"])[:4],
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
