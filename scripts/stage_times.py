"""Stage times (us) of the frozen stress frame for the current ORCVIO_* tuning environment.

    python scripts/stage_times.py [--features 4096] [--repeat 30] [--flush]
    python scripts/stage_times.py --sweep        # re-runs itself over the tuning knobs, one process each

Prints one JSON line per run; used to pick launch bounds / split plans on the GPU box.
"""
import argparse
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KNOBS = ["ORCVIO_TRI_MINB", "ORCVIO_JAC_MINB", "ORCVIO_SYRK_WAVES", "ORCVIO_SYRK_GROUP", "ORCVIO_TRI_CAP"]


def one(n_feat, repeat, flush):
    import numpy as np
    import torch
    import bench
    from orcvio_b200 import api
    snap = bench.make_frame(0, n_feat)
    fr = api.Frame(bench.N_CLONES, 0, bench.NOISE_VAR, 0.95, -1.0, bench.TRI["cost_threshold"],
                   bench.TRI["init_final_dist_threshold"])
    out = fr.update(fr.prepare_inputs(snap))
    fr.load(snap)
    buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None
    for _ in range(3):
        fr.run(1)
    tot = 0.0
    for _ in range(repeat):
        if flush:
            buf.zero_()
            torch.cuda.synchronize()
        tot += fr.run(1)
    _, st = fr.run(10, stages=True)
    kt = fr.kernel_times()
    st = {k: round(v, 1) for k, v in st.items()}
    st.update(syrk=round(kt["syrk"], 1), chol_prior=round(kt["chol_prior"], 1))
    env = {k: os.environ[k] for k in KNOBS if k in os.environ}
    print(json.dumps(dict(features=n_feat, gated=int(((out["status"] & 2) != 0).sum()), us_per_frame=round(tot / repeat, 1),
                          stages=st, env=env, chk=float(np.abs(out["delta_x"]).sum()))), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--features", type=int, default=4096)
    ap.add_argument("--repeat", type=int, default=30)
    ap.add_argument("--flush", action="store_true")
    ap.add_argument("--sweep", nargs="*", default=None, help="KNOB=v1,v2 ... (cartesian product)")
    args = ap.parse_args()
    if args.sweep is None:
        one(args.features, args.repeat, args.flush)
        return
    axes = []
    for spec in args.sweep:
        k, vs = spec.split("=")
        axes.append([(k, v) for v in vs.split(",")])
    for combo in itertools.product(*axes):
        env = dict(os.environ)
        env.update(dict(combo))
        for nf in (2000, 4096):
            cmd = [sys.executable, __file__, "--features", str(nf), "--repeat", str(args.repeat)] + (["--flush"] if args.flush else [])
            subprocess.run(cmd, env=env, check=False)


if __name__ == "__main__":
    main()
