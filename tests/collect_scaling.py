"""Measurement aid: gather the bench lines `bash tests/run_scaling.sh N` wrote under gpurun_out/ into profiles/r2_scaling.json
(one row per workload and N, plus the weak / strong scaling efficiency against the N = 1 row of the same table)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NOTE = {1: "round-2 final", 2: "round-2 final", 4: "round-2 final", 8: "round-2 final"}


def main(out_path: str) -> None:
    table = {"source": "bash tests/run_scaling.sh N under gpurun --gpus N, one B200 box; device-timed, max over ranks", "rows": []}
    base = {}
    for wl in ("train", "export", "projection"):
        for n in (1, 2, 4, 8):
            p = os.path.join(ROOT, "gpurun_out", f"r2_{wl}_{n}gpu.json")
            try:
                d = json.loads(open(p).read().strip().splitlines()[-1])
            except Exception:
                continue
            row = {"workload": wl, "n_gpus": n, "metric": d["metric"], "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                   "scaling": d["scaling"], "e2e": d["e2e"]["value"], "code_state": NOTE[n], "clocks": d.get("clocks"),
                   "data_parallel": d.get("detail", {}).get("data_parallel"), "replicas_identical": d.get("detail", {}).get("replicas_identical"),
                   "ddp_breakdown": d.get("detail", {}).get("ddp_breakdown")}
            if n == 1:
                base[wl] = d["value"]
            if wl in base:
                row["efficiency_vs_1gpu"] = d["value"] / (n * base[wl])
            table["rows"].append(row)
    with open(out_path, "w") as f:
        json.dump(table, f, indent=1)
    for r in table["rows"]:
        print(r["workload"], r["n_gpus"], round(r["value"]), round(r["ms_per_step"], 4), r.get("efficiency_vs_1gpu"), r["data_parallel"], r["replicas_identical"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_scaling.json"))
