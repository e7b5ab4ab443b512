"""-m gpu: temperature calibration through tise_calibration_eval, against golden values produced by
the reference's own classifier_calibration/temperature_scaling.py (tools/make_golden_calibration.py) and against torch
autograd on the same logits."""
import io
import json
import os
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch
from torch import nn

from tise_toolbox_b200 import synth
from tise_toolbox_b200 import temperature_scaling as ts

pytestmark = pytest.mark.gpu

with open(os.path.join(os.path.dirname(__file__), "golden", "calibration.json")) as _f:
    CASES = json.load(_f)["cases"]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "C%d" % c["C"])
def test_set_temperature_matches_reference_golden(case, engine):
    logits, labels = synth.calibration_logits(case["seed"], case["n"], case["C"], case["sharpness"], case["label_noise"])
    model = ts.ModelWithTemperature(nn.Identity(), init_temp=case["init_temp"])
    loader = [(logits[i:i + 500], labels[i:i + 500]) for i in range(0, case["n"], 500)]
    buf = io.StringIO()
    with redirect_stdout(buf):
        assert model.set_temperature(loader) is model
    t = float(model.temperature.item())
    # tolerance: the reference evaluates loss and gradient in fp32 autograd, the kernel in closed form with fp64 sums;
    # LBFGS (the same torch.optim object) turns a 1e-7 gradient difference into ~1e-6 on T
    assert abs(t - case["temperature"]) <= 1e-4 * case["temperature"], (t, case["temperature"])
    rep = model.last_report
    assert abs(rep["before"]["nll"] - case["nll_before"]) <= 1e-5 * case["nll_before"]
    assert abs(rep["after"]["nll"] - case["nll_after"]) <= 1e-5 * case["nll_after"]
    assert abs(rep["before"]["ece"] - case["ece_before"]) <= 2e-4  # a confidence within 1 ulp of a bin edge may change bins
    assert abs(rep["after"]["ece"] - case["ece_after"]) <= 2e-4
    lines = buf.getvalue().splitlines()
    assert lines[0] == case["printed"][0] and lines[1].startswith("Optimal temperature: ") and lines[2].startswith("After temperature - NLL: ")
    # forward(): logits / T
    y = model(logits[:7].cuda())
    assert torch.allclose(y.cpu(), logits[:7] / t, rtol=1e-6)


@pytest.mark.parametrize("T", [0.5, 1.0, 2.1737587451934814])
def test_loss_gradient_and_ece_against_torch(T, engine):
    logits, labels = synth.calibration_logits(7, 1537, 51, 2.0, 0.3)
    r = ts.calibration_eval(logits.cuda(), labels.cuda(), T)
    t = torch.tensor([T], dtype=torch.float64, requires_grad=True)
    loss = nn.functional.cross_entropy(logits.double() / t, labels)
    loss.backward()
    assert abs(r["nll"] - float(loss.detach())) <= 1e-6 * float(loss.detach())
    assert abs(r["dnll_dT"] - float(t.grad)) <= 1e-5 * abs(float(t.grad)) + 1e-9
    p = torch.softmax(logits / T, 1)
    conf, pred = p.max(1)
    assert abs(r["accuracy"] - float((pred == labels).double().mean())) < 1e-12
    assert int(r["bin_count"].sum()) == 1537
    ece = ts._ECELoss()(logits.cuda() / T, labels.cuda())
    assert ece.shape == (1,) and abs(float(ece) - ts.calibration_eval(logits.cuda() / T, labels.cuda(), 1.0)["ece"]) < 1e-7


def test_bad_labels_raise(engine):
    logits, labels = synth.calibration_logits(1, 64, 10)
    labels[5] = 10
    with pytest.raises(IndexError):
        ts.calibration_eval(logits.cuda(), labels.cuda(), 1.0)
