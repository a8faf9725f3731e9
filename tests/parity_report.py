"""GPU parity report (script, not a pytest module): prints and stores the error statistics of the CUDA path against
the oracle for the cases the tests assert on.  ``python tests/parity_report.py [out.json]``."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import product_bundle, product_model, rel_err  # noqa: E402

from cropnerf_b200 import synthetic  # noqa: E402
from oracle import cases  # noqa: E402


def stats(a, b, floor=1e-3):
    e = rel_err(a, b, floor)
    return {"max": float(e.max()), "mean": float(e.mean()), "p999": float(np.quantile(e, 0.999)), "frac_le_1e-4": float((e <= 1e-4).mean())}


def run(name, spec, dev, small=True, precision="fp32"):
    R = spec["num_rays"]
    cfg = cases.make_config(spec.get("cfg"), small=small)
    oracle, state = cases.build_oracle(cfg, 20, 0, spec["table_scale"])
    oracle.train(spec["training"])
    model = product_model(cfg, state, 20, dev, spec["training"], precision=precision)
    rays = synthetic.make_rays(R, seed=1, num_cameras=20)
    if spec["training"]:
        for m in (oracle, model):
            feed = synthetic.JitterFeed(synthetic.make_jitter(R, 3, seed=2))
            m.proposal_sampler.initial_sampler.rand_fn = feed
            m.proposal_sampler.pdf_sampler.rand_fn = feed
        oracle.set_anneal(500)
        model.get_training_callbacks()[0].func(500)
    model.proposal_sampler.pdf_sampler.keep_inds = True
    with torch.no_grad():
        ref = oracle(cases.oracle_bundle(rays))
        out = model(product_bundle(rays, dev))
    rep = {k: stats(out[k], ref[k]) for k in ("rgb", "depth", "accumulation", "prop_depth_0", "prop_depth_1")}
    rep["semantics"] = stats(out["semantics"], ref["semantics"], floor=1e-2)
    rep["label_agreement"] = float((out["semantics_colormap"].cpu() == ref["semantics_colormap"]).float().mean())
    rep["pdf_bins_equal"] = float((model.proposal_sampler.pdf_sampler.last_inds.cpu().long() == oracle.proposal_sampler.pdf_sampler.last_inds).float().mean())
    rep["median_index_equal"] = float((rel_err(out["depth"], ref["depth"]) <= 1e-6).mean())
    return rep


def main():
    dev = torch.device("cuda:0")
    report = {}
    for name, spec in cases.CASES.items():
        report[name] = run(name, spec, dev)
    report["full_size_eval_1024"] = run("full", dict(num_rays=1024, training=False, cfg=dict(), table_scale=0.5), dev, small=False)
    report["full_size_train_1024"] = run("full", dict(num_rays=1024, training=True, cfg=dict(), table_scale=0.5), dev, small=False)
    if os.environ.get("CNB_REPORT_MIXED"):
        report["full_size_eval_1024_mixed"] = run("full", dict(num_rays=1024, training=False, cfg=dict(), table_scale=0.5), dev, small=False, precision="mixed")
    txt = json.dumps(report, indent=1)
    print(txt)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(txt)


if __name__ == "__main__":
    main()
