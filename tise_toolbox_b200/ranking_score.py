"""Drop-in for ranking_scores/ranking_score.py: the reference script ranks every method in
``methods/<name>.json`` on nine metrics and sums six aspect scores into RS.  Only IS*, FID, O-IS and O-FID are
produced by this package (the other five metrics come from detectors / CLIP that are out of scope), so the two
pieces offered are: ``update_method_scores`` -- write / merge this run's numbers into ``methods/<name>.json`` in the
reference's schema -- and ``ranking_table`` -- the reference's computation as a function (host-side tabulation,
no device work; the reference runs it as a script with CWD = ranking_scores/)."""
import json
import os
from collections import OrderedDict

import numpy as np

METRICS = ["IS*", "FID", "RP", "SOA-C", "SOA-I", "O-IS", "O-FID", "CA", "PA"]  # ranking_score.py:11


def update_method_scores(methods_dir, method, **scores):
    """Merges e.g. ``**{"IS*": 33.1, "FID": 12.4, "O-IS": 4.2, "O-FID": 20.1}`` into methods/<method>.json."""
    unknown = [k for k in scores if k not in METRICS]
    if unknown:
        raise KeyError("unknown metric(s) %s; the reference ranks %s" % (unknown, METRICS))
    os.makedirs(methods_dir, exist_ok=True)
    path = os.path.join(methods_dir, method + ".json")
    cur = OrderedDict()
    if os.path.exists(path):
        with open(path) as f:
            cur = json.load(f, object_pairs_hook=OrderedDict)
    for k, v in scores.items():
        cur[k] = float(v)
    with open(path, "w") as f:
        json.dump(cur, f, indent=4)
    return path


def ranking_scores(method_scores):
    """``method_scores``: ordered {method: {metric: value}} -> ordered {method: [9 metric values..., RS]}
    (ranking_score.py:16-66: FID, O-FID and CA are lower-is-better; rank 1 = worst; RS = sum of six aspect scores)."""
    methods = list(method_scores)
    scores = OrderedDict((m, [float(method_scores[m][k]) for k in METRICS]) for m in methods)
    a = np.array([scores[m] for m in methods])
    a[:, 1] = -a[:, 1]
    a[:, 6] = -a[:, 6]
    a[:, 7] = -a[:, 7]
    order = np.argsort(a, 0)
    rank = np.zeros(a.shape)
    for mi in range(len(methods)):
        for ki in range(len(METRICS)):
            rank[mi, ki] = int(np.where(order[:, ki] == mi)[0][0]) + 1
    for mi, m in enumerate(methods):
        r = rank[mi]
        aspects = [np.mean(r[0:2]), r[2], np.mean(r[3:5]), np.mean(r[5:7]), r[7], r[8]]
        scores[m].append(float(np.sum(aspects)))
    return scores


def ranking_table(methods_dir=None, method_scores=None):
    """The psql-format table the reference writes to results/coco_benchmark_results.txt."""
    import pandas as pd
    from tabulate import tabulate

    if method_scores is None:
        method_scores = OrderedDict()
        for f in os.listdir(methods_dir):  # ranking_score.py:14, filesystem order
            if f.split(".")[1] == "json":
                with open(os.path.join(methods_dir, f)) as fh:
                    method_scores[f.split(".")[0]] = json.load(fh)
    scores = ranking_scores(method_scores)
    df = pd.DataFrame(np.array([scores[m] for m in scores]), columns=METRICS + ["RS"])
    df.insert(loc=0, column="Method", value=list(scores))
    return tabulate(df, headers="keys", tablefmt="psql", showindex=False)
