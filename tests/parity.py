"""Shared parity bookkeeping of the GPU tests: BASELINE.json's north_star bars and a collector that records every
measured number (so one GPU run yields the whole table of DESIGN.md §5) before asserting them all at once.

Bars (north_star): per-boundary activation max-rel-error <= 1e-2, gradient cosine >= 0.999, losses within 2 %."""
import contextlib
import io
import json
import os

ACT_BAR, COS_BAR, LOSS_BAR = 1e-2, 0.999, 0.02

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLE = os.path.join(ROOT, "gpurun_out", "parity_table.jsonl")


def quiet(fn):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn()


def relerr(a, b):
    """max |a - b| / max |b| over the tensor (SURVEY.md §8c's activation metric)."""
    return ((a.detach().float().cpu() - b.float()).abs().max() / (b.float().abs().max() + 1e-30)).item()


def global_cos(named_params, ref, skip=()):
    """Global cosine between the .grad of the named parameters and the reference gradients `ref` (dict name -> tensor);
    names in `skip` (analytically-zero gradients: biases feeding a normalisation) are left out."""
    params = dict(named_params)
    num = da = db = 0.0
    for k, r in ref.items():
        if k in skip:
            continue
        g = params[k].grad.detach().float().cpu().double()
        r = r.double()
        num += (g * r).sum().item()
        da += (g * g).sum().item()
        db += (r * r).sum().item()
    return num / (da ** 0.5 * db ** 0.5 + 1e-30)


def prebn_biases(net):
    """Conv biases that feed a BatchNorm: analytically-zero gradient, rounding noise in the reference.
    DCGAN-family naming: `blocks.i.0.bias` followed by `blocks.i.1.weight`."""
    names = dict(net.named_parameters())
    return [k for k in names if k.endswith(".0.bias") and k.replace(".0.bias", ".1.weight") in names]


def prebn_biases_blur_g(net):
    """dcgan_blur.Generator blocks are [Upsample, Conv, BlurPool, BatchNorm, LeakyReLU]: `blocks.i.1.bias` feeds the
    BatchNorm `blocks.i.3` through the (linear, bias-preserving) blur."""
    names = dict(net.named_parameters())
    return [k for k in names if k.endswith(".1.bias") and k.replace(".1.bias", ".3.weight") in names]


def resnet_g_zero_biases(net):
    """ResNetGenerator: every block conv (c1 -> b2; c2 + c_sc -> the next block's b1 / b6) feeds a BatchNorm."""
    return [k for k, _ in net.named_parameters() if k.startswith("block") and k.endswith(("c1.bias", "c2.bias", "c_sc.bias"))]


class Bars:
    """Collects (kind, name, value) rows for one configuration, prints them, appends them to gpurun_out/parity_table.jsonl
    and asserts the north_star bars on all of them at the end (so a failing row does not hide the others)."""

    def __init__(self, config, act_bar=ACT_BAR, cos_bar=COS_BAR, loss_bar=LOSS_BAR, loss_abs=0.0):
        """loss_abs: absolute scale added to |ref| in the loss rows — for losses that sit near zero (the hinge generator
        loss -mean D(G(z)) of a fresh network), where a relative error has no meaning; 0 for everything else."""
        self.config, self.rows, self.loss_abs = config, [], loss_abs
        self.bars = {"act": act_bar, "cos": cos_bar, "loss": loss_bar}

    def act(self, name, got, ref):
        self.rows.append(("act", name, relerr(got, ref)))

    def cos(self, name, value):
        self.rows.append(("cos", name, float(value)))

    def loss(self, name, got, ref):
        got, ref = float(got), float(ref)
        self.rows.append(("loss", name, abs(got - ref) / (abs(ref) + self.loss_abs + 1e-12)))

    def ok(self, kind, v):
        return v >= self.bars[kind] if kind == "cos" else v <= self.bars[kind]

    def finish(self):
        line = " | ".join("%s %s %s%s" % (k, n, ("%.6f" % v) if k == "cos" else ("%.2e" % v), "" if self.ok(k, v) else " (!)")
                          for k, n, v in self.rows)
        print("\n[%s] %s" % (self.config, line))
        try:
            os.makedirs(os.path.dirname(TABLE), exist_ok=True)
            with open(TABLE, "a") as f:
                f.write(json.dumps({"config": self.config, "rows": self.rows, "bars": self.bars}) + "\n")
        except OSError:
            pass
        bad = [(k, n, v) for k, n, v in self.rows if not self.ok(k, v)]
        assert not bad, "[%s] below the north_star bars %s: %s" % (self.config, self.bars, bad)
