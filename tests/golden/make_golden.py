"""Generates tests/golden/loss_golden.json: values and gradients of every loss on the path at a few
(target, output) pairs, computed INDEPENDENTLY of oracle/ and of the CUDA code with Python's math module straight
from the formulas of the reference headers (include/nano/loss/*.h, src/loss/pinball.cpp) — plus the reference's own
pinball golden vectors (test/test_loss.cpp:271-307) verbatim.

The reference is C++ and cannot be built here (Eigen3 missing), so these fixtures are formula-level known answers,
not outputs of the reference binary.   python tests/golden/make_golden.py
"""
import json
import math
import os


def sign(a):
    return float((a > 0) - (a < 0))


def elementwise(loss, t, o, alpha):
    to = t * o
    if loss == "mse":
        return (o - t) ** 2, o - t
    if loss == "mae":
        return abs(o - t), sign(o - t)
    if loss == "cauchy":
        return math.log((t - o) ** 2 + 1.0), (o - t) / (1.0 + (o - t) ** 2)
    if loss.endswith("squared-hinge"):
        m = max(1.0 - to, 0.0)
        return m * m, -t * m * 2.0
    if loss.endswith("hinge"):
        return max(1.0 - to, 0.0), -t * (sign(1.0 - to) + 1.0) * 0.5
    if loss.endswith("savage"):
        return 1.0 / (1.0 + math.exp(to)) ** 2, -2.0 * t / ((1.0 + math.exp(to)) * (2.0 + math.exp(to) + math.exp(-to)))
    if loss.endswith("tangent"):
        a = 2.0 * math.atan(to) - 1.0
        return a * a, 4.0 * t * a / (1.0 + to * to)
    if loss.endswith("logistic"):
        x = -to
        value = math.log1p(math.exp(x)) if x < 1.0 else x + math.log1p(math.exp(-x))
        g = math.exp(x) / (1.0 + math.exp(x)) if x < 1.0 else 1.0 / (1.0 + math.exp(-x))
        return value, -t * g
    if loss.endswith("exponential"):
        return math.exp(-to), -t * math.exp(-to)
    if loss == "pinball":
        return alpha * max(t - o, 0.0) + (1.0 - alpha) * max(o - t, 0.0), -alpha + 0.5 * (1.0 - sign(t - o))
    raise ValueError(loss)


def sample(loss, targets, outputs, alpha):
    if loss == "s-classnll":
        omax = max(outputs)
        exps = [math.exp(o - omax) for o in outputs]
        total = sum(exps)
        value = math.log(total) - 0.5 * sum((1.0 + t) * o for t, o in zip(targets, outputs)) + omax
        return value, [e / total - 0.5 * (1.0 + t) for e, t in zip(exps, targets)]
    pairs = [elementwise(loss, t, o, alpha) for t, o in zip(targets, outputs)]
    value = sum(p[0] for p in pairs)
    if loss in ("mse", "cauchy"):
        value *= 0.5
    return value, [p[1] for p in pairs]


CLASS_T = [[1.0, -1.0, -1.0], [-1.0, 1.0, -1.0], [-1.0, -1.0, 1.0], [1.0, -1.0, -1.0]]
CLASS_O = [[0.5, -0.25, 2.0], [-1.5, 1.0, 0.0], [3.0, -2.0, 0.125], [1.0, 1.0, -1.0]]
REG_T = [[0.0, 0.5, -1.0], [2.0, -0.75, 0.25], [1.0, 1.0, 1.0], [-3.0, 0.0, 0.5]]
REG_O = [[0.25, 0.5, 1.0], [-1.0, -0.5, 0.25], [1.0, 2.0, 0.0], [-2.0, 0.125, 4.0]]

cases = []
for loss in ("mse", "mae", "cauchy", "pinball"):
    for alpha in ((0.5, 0.2) if loss == "pinball" else (0.5,)):
        rows = [sample(loss, t, o, alpha) for t, o in zip(REG_T, REG_O)]
        cases.append({"loss": loss, "alpha": alpha, "targets": REG_T, "outputs": REG_O,
                      "values": [r[0] for r in rows], "vgrads": [r[1] for r in rows]})
for loss in ("s-hinge", "m-hinge", "s-squared-hinge", "s-classnll", "s-savage", "s-tangent", "s-logistic",
             "m-logistic", "s-exponential"):
    rows = [sample(loss, t, o, 0.5) for t, o in zip(CLASS_T, CLASS_O)]
    cases.append({"loss": loss, "targets": CLASS_T, "outputs": CLASS_O, "values": [r[0] for r in rows],
                  "vgrads": [r[1] for r in rows]})

# the reference's own golden vectors (test/test_loss.cpp:271-307)
PIN_T = [[0.0, 0.1], [0.2, 0.3], [0.4, 0.5], [0.6, 0.7], [0.8, 0.9]]
PIN_O = [[0.1, 0.0], [0.2, 0.4], [0.2, 0.3], [0.7, 0.9], [0.9, 0.6]]
for alpha, expected in ((0.5, [0.10, 0.05, 0.20, 0.15, 0.20]), (0.2, [0.10, 0.08, 0.08, 0.24, 0.14]),
                        (0.8, [0.10, 0.02, 0.32, 0.06, 0.26])):
    rows = [sample("pinball", t, o, alpha) for t, o in zip(PIN_T, PIN_O)]
    assert all(abs(r[0] - e) < 1e-12 for r, e in zip(rows, expected))
    cases.append({"loss": "pinball", "alpha": alpha, "targets": PIN_T, "outputs": PIN_O, "values": expected,
                  "vgrads": [r[1] for r in rows], "source": "test/test_loss.cpp:271-307"})

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_golden.json")
with open(path, "w") as f:
    json.dump({"generator": "tests/golden/make_golden.py", "cases": cases}, f, indent=1)
print(path, len(cases), "cases")
