"""tests/golden/calibration.json: the reference's OWN classifier_calibration/temperature_scaling.py run here on CPU
(its unconditional .cuda() calls patched to no-ops) on seeded logits, and the reference's ranking_score.py run on its
own methods/*.json.    python tools/make_golden_calibration.py"""
import contextlib
import importlib.util
import io
import json
import os
import subprocess
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tise_toolbox_b200 import synth  # noqa: E402

REF = "/root/reference"


def main():
    spec = importlib.util.spec_from_file_location("ref_ts", os.path.join(REF, "classifier_calibration", "temperature_scaling.py"))
    ts = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ts)
    keep_t, keep_m = torch.Tensor.cuda, nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    out = {"cases": []}
    try:
        for seed, n, c, sharp, noise, init in [(1, 4000, 50, 3.0, 0.25, 1.0), (2, 2500, 80, 5.0, 0.4, 1.5), (3, 3000, 1008, 4.0, 0.3, 1.0)]:
            logits, labels = synth.calibration_logits(seed, n, c, sharp, noise)
            m = ts.ModelWithTemperature(nn.Identity(), init_temp=init)
            loader = [(logits[i:i + 500], labels[i:i + 500]) for i in range(0, n, 500)]
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                m.set_temperature(loader)
            t = float(m.temperature.item())
            ece = ts._ECELoss()
            crit = nn.CrossEntropyLoss()
            case = {"seed": seed, "n": n, "C": c, "sharpness": sharp, "label_noise": noise, "init_temp": init, "temperature": t,
                    "nll_before": float(crit(logits, labels)), "ece_before": float(ece(logits, labels)),
                    "nll_after": float(crit(logits / t, labels)), "ece_after": float(ece(logits / t, labels)),
                    "printed": buf.getvalue().splitlines()}
            out["cases"].append(case)
            print(case)
    finally:
        torch.Tensor.cuda, nn.Module.cuda = keep_t, keep_m
    # ranking_score.py on the reference's own methods/*.json (run in a scratch copy: it writes results/)
    import shutil
    import tempfile
    tmp = tempfile.mkdtemp()
    shutil.copytree(os.path.join(REF, "ranking_scores", "methods"), os.path.join(tmp, "methods"))
    os.makedirs(os.path.join(tmp, "results"))
    subprocess.check_call([sys.executable, os.path.join(REF, "ranking_scores", "ranking_score.py")], cwd=tmp, stdout=subprocess.DEVNULL)
    table = open(os.path.join(tmp, "results", "coco_benchmark_results.txt")).read()
    methods = {}
    for f in sorted(os.listdir(os.path.join(tmp, "methods"))):
        if f.endswith(".json"):
            methods[f[:-5]] = json.load(open(os.path.join(tmp, "methods", f)))
    out["ranking"] = {"methods": methods, "listdir_order": [f.split(".")[0] for f in os.listdir(os.path.join(tmp, "methods")) if f.split(".")[1] == "json"],
                      "table": table}
    with open(os.path.join(ROOT, "tests", "golden", "calibration.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
