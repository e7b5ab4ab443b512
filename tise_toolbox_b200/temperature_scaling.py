"""Drop-in for classifier_calibration/temperature_scaling.py on B200: ``ModelWithTemperature`` and ``_ECELoss``
with the reference's constructor arguments, method names and printed lines.

The reference optimises ONE scalar (the temperature) with ``optim.LBFGS(lr=0.01, max_iter=50)`` on
``CrossEntropyLoss(logits / T, labels)`` (:56-63).  Here the very same ``torch.optim.LBFGS`` object drives the
iteration on the host -- so the step sequence, tolerances and stopping rules are the reference's by construction --
while every closure evaluation (loss, dLoss/dT, and the ECE histogram) is one launch of libtise_b200's
``tise_calibration_eval`` over the validation logits resident in HBM.  The resulting constants are what
IS* / O-IS divide their logits by (inception_score_star_coco.py:107, inception_score_star_bird.py:192,
object_centric_inception_score.py:55).
"""
import ctypes as C

import numpy as np
import torch
from torch import nn, optim

from .engine import Engine, _stream, default_engine


def calibration_eval(logits, labels, temperature=1.0, n_bins=15, engine=None):
    """One pass over device logits [n, C] (float32) / labels [n] (int64) at ``temperature``:
    returns dict(nll, dnll_dT, accuracy, ece, bin_count, bin_conf, bin_acc) -- nll and its derivative are means."""
    e = engine or default_engine(logits.device)
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2
    labels = labels.to(logits.device, torch.int64).contiguous()
    logits = logits.contiguous()
    n, c = int(logits.shape[0]), int(logits.shape[1])
    assert labels.numel() == n
    bounds = torch.linspace(0, 1, n_bins + 1).numpy().astype(np.float32)  # _ECELoss.__init__ (:103-105)
    out = torch.zeros(4 + 3 * n_bins + 1, dtype=torch.float64, device=logits.device)
    with torch.cuda.device(logits.device):
        e._check(e.lib.tise_calibration_eval(e.ctx, logits.data_ptr(), labels.data_ptr(), n, c, float(temperature),
                                             bounds.ctypes.data, n_bins, out.data_ptr(), _stream()))
    h = out.cpu().numpy()
    if h[4 + 3 * n_bins:].view(np.int32)[0] != 0:
        raise IndexError("Target out of bounds: a label is outside [0, %d)" % c)  # what CrossEntropyLoss raises
    rows = max(h[2], 1.0)
    cnt, conf, acc = h[4:4 + n_bins], h[4 + n_bins:4 + 2 * n_bins], h[4 + 2 * n_bins:4 + 3 * n_bins]
    ece = 0.0
    for b in range(n_bins):  # :110-117
        if cnt[b] > 0:
            ece += abs(conf[b] / cnt[b] - acc[b] / cnt[b]) * (cnt[b] / rows)
    return {"nll": h[0] / rows, "dnll_dT": h[1] / rows, "accuracy": h[3] / rows, "ece": ece, "bin_count": cnt,
            "bin_conf": conf, "bin_acc": acc}


class ModelWithTemperature(nn.Module):
    """temperature_scaling.py:7-77.  ``model`` maps an input batch to classification logits on the GPU."""

    def __init__(self, model, init_temp=1.0):
        super(ModelWithTemperature, self).__init__()
        self.model = model
        self.temperature = nn.Parameter(torch.ones(1) * init_temp)

    def forward(self, input):
        logits = self.model(input)
        return self.temperature_scale(logits)

    def temperature_scale(self, logits):
        temperature = self.temperature.to(logits.device).unsqueeze(1).expand(logits.size(0), logits.size(1))
        return logits / temperature

    def set_temperature(self, valid_loader):
        """Tune the temperature on the validation set (NLL), printing the reference's three lines."""
        logits_list, labels_list = [], []
        with torch.no_grad():
            for input, label in valid_loader:
                logits_list.append(self.model(input.cuda() if torch.is_tensor(input) else input).float())
                labels_list.append(label)
        logits = torch.cat(logits_list).cuda()
        labels = torch.cat(labels_list).cuda()
        self.set_temperature_from_logits(logits, labels)
        return self

    def set_temperature_from_logits(self, logits, labels, verbose=True):
        before = calibration_eval(logits, labels, 1.0)
        if verbose:
            print("Before temperature - NLL: %.3f, ECE: %.3f" % (before["nll"], before["ece"]))
        optimizer = optim.LBFGS([self.temperature], lr=0.01, max_iter=50)  # :56

        def eval():
            r = calibration_eval(logits, labels, float(self.temperature.item()))
            g = torch.full_like(self.temperature, r["dnll_dT"])
            # the reference's closure (:58-61) calls loss.backward() WITHOUT zeroing the gradient first, so autograd
            # accumulates over the closure calls of the LBFGS step; kept, because it decides where T ends up
            self.temperature.grad = g if self.temperature.grad is None else self.temperature.grad + g
            return torch.tensor(r["nll"], dtype=torch.float32)

        optimizer.step(eval)
        after = calibration_eval(logits, labels, float(self.temperature.item()))
        if verbose:
            print("Optimal temperature: {}".format(self.temperature.item()))
            print("After temperature - NLL: %.10f, ECE: %.10f" % (after["nll"], after["ece"]))
        self.last_report = {"before": before, "after": after, "temperature": float(self.temperature.item())}
        return self


class _ECELoss(nn.Module):
    """temperature_scaling.py:80-119: expected calibration error of logits (NOT softmax scores), 15 bins."""

    def __init__(self, n_bins=15):
        super(_ECELoss, self).__init__()
        bin_boundaries = torch.linspace(0, 1, n_bins + 1)
        self.bin_lowers = bin_boundaries[:-1]
        self.bin_uppers = bin_boundaries[1:]
        self.n_bins = n_bins

    def forward(self, logits, labels):
        r = calibration_eval(logits, labels, 1.0, self.n_bins)
        return torch.tensor([r["ece"]], dtype=torch.float32, device=logits.device)


class EngineClassifier:
    """``model`` for ModelWithTemperature: uint8 image batches -> logits of the loaded Inception-v3 head."""

    def __init__(self, engine: Engine, use_bias=True, class_offset=0):
        self.e, self.use_bias, self.off = engine, use_bias, class_offset

    def __call__(self, images):
        return self.e.logits(self.e.features(images), use_bias=self.use_bias)[:, self.off:]
